// HBM-bound kernels of the chunk pipeline AROUND the generator:
//   * s3_output_transform : u / v -> windspeed / winddirection on the rotated grid + physical limits
//                           (sup3r/writers/base.py:233-346, preprocessing/derivers/utilities.py:204-255,
//                           utilities/utilities.py:155-220), in place on the cropped chunk
//   * s3_coarsen          : batch production, hi-res sample -> low-res input: (s x s) block mean then
//                           temporal subsample / average / total / max / min
//                           (preprocessing/batch_queues/base.py:32-87, utilities/utilities.py:345-523)
//   * s3_gauss_smooth2d   : scipy.ndimage.gaussian_filter(sigma, mode='nearest', truncate=4) on the
//                           (s1, s2) planes of selected features (batch_queues/utilities.py:57-104)
//   * s3_gather_samples   : random (s1, s2, t) crops of a device-resident hi-res dataset
// One thread per voxel / output element, channel-fastest coalesced accesses, grids sized to the
// SM count.  All memory bound: bytes touched once.
#include "common.cuh"

namespace s3 {

static inline unsigned pgrid(size_t n, int threads = 256) {
  size_t blocks = (n + threads - 1) / threads;
  size_t cap = (size_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

constexpr int kMaxF = 32;
constexpr int kMaxPairs = 8;

struct OutXform {
  int f, n_pairs, clip;
  int iu[kMaxPairs], iv[kMaxPairs];
  float lo[kMaxF], hi[kMaxF];
};

// data (S1*S2, T, F) in place.  cs: (S1*S2, 2) = (cos theta, sin theta) of the grid rotation.
// counts[2 f] / counts[2 f + 1]: voxels of feature f below lo / above hi BEFORE clipping (the
// reference warns with these fractions and, for nn_fill, replaces exactly those points).
__global__ void output_transform_kernel(float* __restrict__ data, const float* __restrict__ cs,
                                        size_t n_sp, int T, OutXform p,
                                        unsigned long long* __restrict__ counts) {
  __shared__ unsigned int sc[2 * kMaxF];
  for (int i = threadIdx.x; i < 2 * p.f; i += blockDim.x) sc[i] = 0u;
  __syncthreads();
  const size_t total = n_sp * (size_t)T;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    float* row = data + idx * p.f;
    const size_t sp = idx / T;
    float v[kMaxF];
    for (int i = 0; i < p.f; ++i) v[i] = row[i];
    if (p.n_pairs) {
      const float c = cs[2 * sp], s = cs[2 * sp + 1];
      for (int k = 0; k < p.n_pairs; ++k) {
        const float u = v[p.iu[k]], w = v[p.iv[k]];
        const float ur = c * u - s * w, vr = s * u + c * w;
        v[p.iu[k]] = hypotf(ur, vr);
        float wd = atan2f(ur, vr) * 57.29577951308232f + 360.f;
        v[p.iv[k]] = fmodf(wd, 360.f);
      }
    }
    for (int i = 0; i < p.f; ++i) {
      const float x = v[i];
      if (x < p.lo[i]) atomicAdd(&sc[2 * i], 1u);
      if (x > p.hi[i]) atomicAdd(&sc[2 * i + 1], 1u);
      row[i] = p.clip ? fminf(fmaxf(x, p.lo[i]), p.hi[i]) : x;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * p.f; i += blockDim.x)
    if (sc[i]) atomicAdd(&counts[i], (unsigned long long)sc[i]);
}

// method: 0 subsample, 1 average (nansum / t), 2 total (nansum), 3 max, 4 min
__global__ void coarsen_kernel(const float* __restrict__ hr, float* __restrict__ lr, int B, int S1,
                               int S2, int T, int F, int s, int t, int method, size_t total) {
  const int L1 = S1 / s, L2 = S2 / s, LT = T / t;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int f = (int)(r % F); r /= F;
    const int lt = (int)(r % LT); r /= LT;
    const int l2 = (int)(r % L2); r /= L2;
    const int l1 = (int)(r % L1); r /= L1;
    const int b = (int)r;
    const int nt = method == 0 ? 1 : t;
    float acc = method == 3 ? -INFINITY : (method == 4 ? INFINITY : 0.f);
    bool any_nan = false;
    for (int dt = 0; dt < nt; ++dt) {
      const int tt = lt * t + dt;
      float sum = 0.f;   // the (s x s) block mean of this hi-res time step
      for (int a = 0; a < s; ++a)
        for (int c = 0; c < s; ++c)
          sum += hr[((((size_t)b * S1 + l1 * s + a) * S2 + l2 * s + c) * T + tt) * F + f];
      const float m = sum / (float)(s * s);
      if (method == 0) acc = m;
      else if (method == 1 || method == 2) { if (m == m) acc += m; }     // nansum
      else if (method == 3) { any_nan |= m != m; acc = fmaxf(acc, m); }
      else { any_nan |= m != m; acc = fminf(acc, m); }
    }
    if (method == 1) acc /= (float)t;
    if ((method == 3 || method == 4) && any_nan) acc = NAN;             // np.max / np.min
    lr[idx] = acc;
  }
}

// one pass of the separable filter along axis `ax` (0: s1, 1: s2) of x (B, S1, S2, T, F), features
// with fmask bit set only (others are copied); nearest-edge extension; w: 2 r + 1 weights
__global__ void gauss1d_kernel(const float* __restrict__ x, float* __restrict__ y, int S1, int S2,
                               int TF, int F, unsigned fmask, int ax, const float* __restrict__ w,
                               int r, size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t q = idx;
    const int tf = (int)(q % TF); q /= TF;
    const int j = (int)(q % S2); q /= S2;
    const int i = (int)(q % S1); q /= S1;
    const size_t b = q;
    const int f = tf % F;
    if (!((fmask >> f) & 1u)) { y[idx] = x[idx]; continue; }
    const int n = ax == 0 ? S1 : S2, pos = ax == 0 ? i : j;
    const size_t stride = ax == 0 ? (size_t)S2 * TF : (size_t)TF;
    const size_t base = ((b * S1 + (ax == 0 ? 0 : i)) * S2 + (ax == 0 ? j : 0)) * TF + tf;
    double acc = 0.0;   // scipy's correlate1d accumulates in double
    for (int k = -r; k <= r; ++k) {
      int p = pos + k;
      p = p < 0 ? 0 : (p >= n ? n - 1 : p);
      acc += (double)w[k + r] * (double)x[base + (size_t)p * stride];
    }
    y[idx] = (float)acc;
  }
}

// data (S1, S2, T, F) -> out (B, s1, s2, t, F); origins (B, 3) int32 = (i0, j0, t0)
__global__ void gather_samples_kernel(const float* __restrict__ data, float* __restrict__ out,
                                      const int* __restrict__ org, int S2, int T, int F, int s1,
                                      int s2, int t, size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t q = idx;
    const int f = (int)(q % F); q /= F;
    const int k = (int)(q % t); q /= t;
    const int j = (int)(q % s2); q /= s2;
    const int i = (int)(q % s1); q /= s1;
    const int b = (int)q;
    const int i0 = org[3 * b], j0 = org[3 * b + 1], t0 = org[3 * b + 2];
    out[idx] = data[(((size_t)(i0 + i) * S2 + (j0 + j)) * T + (t0 + k)) * F + f];
  }
}

// ---------------------------------------------------------------- quantile delta mapping
// Empirical QDM (Cannon et al. 2015, eq. 3-6) as the reference applies it to a low-res chunk
// (sup3r/bias/bias_transforms.py:490-619 -> rex.utilities.bc_utils.QuantileDeltaMapping):
//   tau = F_mf(x);  x_oh = F_oh^-1(tau);  x_mh = F_mh^-1(tau);
//   relative: x_oh * (x / x_mh)      absolute: x_oh + (x - x_mh)
// with every CDF a table of N quantile values per site and time window, evaluated by linear
// interpolation.  interp() restates numpy's np.interp in double precision without fused
// multiply-adds (last index with xp[j] <= x; exact hit returns fp[j]; its NaN fall-backs).
__device__ __forceinline__ double qdm_interp(double x, const float* __restrict__ xp_f,
                                             const double* __restrict__ xp_d,
                                             const float* __restrict__ fp_f,
                                             const double* __restrict__ fp_d, int n) {
  // exactly one of (xp_f, xp_d) and one of (fp_f, fp_d) is non-null
  auto XP = [&](int i) { return xp_f ? (double)xp_f[i] : xp_d[i]; };
  auto FP = [&](int i) { return fp_f ? (double)fp_f[i] : fp_d[i]; };
  if (isnan(x)) return x;
  if (x > XP(n - 1)) return FP(n - 1);
  if (x < XP(0)) return FP(0);
  int lo = 0, hi = n;                 // first index with xp > x
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (x >= XP(mid)) lo = mid + 1; else hi = mid;
  }
  const int j = lo - 1;
  if (j < 0) return FP(0);
  if (j >= n - 1) return FP(n - 1);
  const double xj = XP(j), fj = FP(j);
  if (xj == x) return fj;
  const double xj1 = XP(j + 1), fj1 = FP(j + 1);
  const double slope = __ddiv_rn(__dsub_rn(fj1, fj), __dsub_rn(xj1, xj));
  double r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xj)), fj);
  if (isnan(r)) {
    r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xj1)), fj1);
    if (isnan(r) && fj == fj1) r = fj;
  }
  return r;
}

struct QdmParams {
  int n_sites, n_times, n_win, n_q, relative, has_zero, has_min, has_range, has_out_range;
  double denom_zero, denom_min, delta_lo, delta_hi, out_lo, out_hi;
};

// data / out: (site, time) float32; win: window index of every time step; tables (site, window,
// quantile) float32; q: the N quantile levels (double).  bad: count of non-finite results.
__global__ void qdm_kernel(const float* __restrict__ data, const int* __restrict__ win,
                           const float* __restrict__ oh, const float* __restrict__ mh,
                           const float* __restrict__ mf, const double* __restrict__ q,
                           const float* __restrict__ tau_fut, const double* __restrict__ k_factor,
                           QdmParams p, float* __restrict__ out,
                           unsigned long long* __restrict__ bad) {
  const size_t total = (size_t)p.n_sites * p.n_times;
  unsigned local_bad = 0, local_nan = 0;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(idx % p.n_times);
    const size_t site = idx / p.n_times;
    const int w = win[t];
    const size_t row = (site * p.n_win + w) * p.n_q;
    const double x = (double)data[idx];
    const double tau = qdm_interp(x, mf + row, nullptr, nullptr, q, p.n_q);
    const double x_oh = qdm_interp(tau, nullptr, q, oh + row, nullptr, p.n_q);
    double x_mh = qdm_interp(tau, nullptr, q, mh + row, nullptr, p.n_q);
    double r;
    if (p.relative) {
      if (p.has_zero && x_mh == 0.0) x_mh = p.denom_zero;
      if (p.has_min && !isnan(x_mh)) x_mh = fmax(x_mh, p.denom_min);   // np.maximum keeps NaN
      double delta = __ddiv_rn(x, x_mh);
      if (p.has_range && !isnan(delta)) delta = fmin(fmax(delta, p.delta_lo), p.delta_hi);
      r = __dmul_rn(x_oh, delta);
    } else {
      double delta = __dsub_rn(x, x_mh);
      if (p.has_range && !isnan(delta)) delta = fmin(fmax(delta, p.delta_lo), p.delta_hi);
      r = __dadd_rn(x_oh, delta);
    }
    // PresRat (bias_transforms.py:1117-1120): dry days below the future zero-rate threshold,
    // wet days scaled by the window's K factor
    if (tau_fut) r = r < (double)tau_fut[site] ? 0.0 : __dmul_rn(r, k_factor[site * p.n_win + w]);
    float rf = (float)r;
    if (p.has_out_range && !isnan(rf)) rf = fminf(fmaxf(rf, (float)p.out_lo), (float)p.out_hi);
    if (!isfinite(rf)) ++local_bad;
    if (isnan(rf)) ++local_nan;
    out[idx] = rf;
  }
  if (local_bad) atomicAdd(bad, (unsigned long long)local_bad);
  if (local_nan) atomicAdd(bad + 1, (unsigned long long)local_nan);
}

}  // namespace s3

using namespace s3;

extern "C" int s3_output_transform(float* data, size_t n_spatial, int n_t, int n_f,
                                   const float* cos_sin, const int* pair_u, const int* pair_v,
                                   int n_pairs, const float* lo, const float* hi, int clip,
                                   unsigned long long* counts, s3_stream stream) {
  S3_REQUIRE(data && lo && hi && counts, "s3_output_transform: null argument");
  S3_REQUIRE(n_f >= 1 && n_f <= kMaxF, "s3_output_transform: 1 <= features <= %d, got %d", kMaxF, n_f);
  S3_REQUIRE(n_pairs >= 0 && n_pairs <= kMaxPairs, "s3_output_transform: at most %d u/v pairs", kMaxPairs);
  S3_REQUIRE(n_pairs == 0 || (cos_sin && pair_u && pair_v), "s3_output_transform: u/v pairs need cos_sin");
  OutXform p;
  p.f = n_f; p.n_pairs = n_pairs; p.clip = clip;
  for (int k = 0; k < n_pairs; ++k) {
    S3_REQUIRE(pair_u[k] >= 0 && pair_u[k] < n_f && pair_v[k] >= 0 && pair_v[k] < n_f &&
                   pair_u[k] != pair_v[k], "s3_output_transform: bad u/v pair %d", k);
    p.iu[k] = pair_u[k]; p.iv[k] = pair_v[k];
  }
  for (int i = 0; i < n_f; ++i) { p.lo[i] = lo[i]; p.hi[i] = hi[i]; }
  S3_CUDA(cudaMemsetAsync(counts, 0, 2 * n_f * sizeof(unsigned long long), as_stream(stream)));
  const size_t total = n_spatial * (size_t)n_t;
  if (total)
    output_transform_kernel<<<pgrid(total), 256, 0, as_stream(stream)>>>(data, cos_sin, n_spatial,
                                                                          n_t, p, counts);
  S3_LAUNCH_CHECK("output_transform");
  return S3_OK;
}

extern "C" int s3_coarsen(const float* hr, float* lr, int n, int s1, int s2, int t, int f,
                          int s_enhance, int t_enhance, int method, s3_stream stream) {
  S3_REQUIRE(hr && lr, "s3_coarsen: null argument");
  S3_REQUIRE(s_enhance >= 1 && t_enhance >= 1 && s1 % s_enhance == 0 && s2 % s_enhance == 0 &&
                 t % t_enhance == 0,
             "s3_coarsen: extents (%d, %d, %d) must be multiples of the enhancements (%d, %d)", s1,
             s2, t, s_enhance, t_enhance);
  S3_REQUIRE(method >= 0 && method <= 4, "s3_coarsen: method must be 0..4");
  const size_t total = (size_t)n * (s1 / s_enhance) * (s2 / s_enhance) * (t / t_enhance) * f;
  if (total)
    coarsen_kernel<<<pgrid(total), 256, 0, as_stream(stream)>>>(hr, lr, n, s1, s2, t, f, s_enhance,
                                                                 t_enhance, method, total);
  S3_LAUNCH_CHECK("coarsen");
  return S3_OK;
}

extern "C" int s3_gauss_smooth2d(const float* x, float* tmp, float* y, int n, int s1, int s2,
                                 int tf, int f, unsigned feature_mask, const float* weights,
                                 int radius, s3_stream stream) {
  S3_REQUIRE(x && tmp && y && weights && radius >= 0, "s3_gauss_smooth2d: bad arguments");
  S3_REQUIRE(f >= 1 && f <= 32 && tf % f == 0, "s3_gauss_smooth2d: 1 <= features <= 32");
  const size_t total = (size_t)n * s1 * s2 * tf;
  if (total) {
    // scipy filters axis 0 first, storing the intermediate in the output dtype (float32)
    gauss1d_kernel<<<pgrid(total), 256, 0, as_stream(stream)>>>(x, tmp, s1, s2, tf, f, feature_mask,
                                                                 0, weights, radius, total);
    gauss1d_kernel<<<pgrid(total), 256, 0, as_stream(stream)>>>(tmp, y, s1, s2, tf, f, feature_mask,
                                                                 1, weights, radius, total);
  }
  S3_LAUNCH_CHECK("gauss_smooth2d");
  return S3_OK;
}

extern "C" int s3_gather_samples(const float* data, int S1, int S2, int T, int F, const int* origins,
                                 int n, int s1, int s2, int t, float* out, s3_stream stream) {
  S3_REQUIRE(data && origins && out, "s3_gather_samples: null argument");
  S3_REQUIRE(s1 <= S1 && s2 <= S2 && t <= T, "s3_gather_samples: sample larger than the data");
  const size_t total = (size_t)n * s1 * s2 * t * F;
  if (total)
    gather_samples_kernel<<<pgrid(total), 256, 0, as_stream(stream)>>>(data, out, origins, S2, T, F,
                                                                        s1, s2, t, total);
  S3_LAUNCH_CHECK("gather_samples");
  return S3_OK;
}

// ---------------------------------------------------------------------------------------------
// Gradient all-reduce over NVLink peer memory (sup3r/models/abstract.py:785-805 _sum_parallel_grad:
// the SUM of the per-GPU gradients).  Every rank's flat fp32 gradient arena lives in symmetric
// memory (peer-mapped through NVSwitch); after a cross-GPU barrier each rank reads ALL arenas with
// plain P2P loads and adds them IN RANK ORDER -- the same order on every rank, so the result is
// bit-identical everywhere and equals the reference's sequential sum over the shards.
namespace s3 {

struct PeerPtrs {
  const float4* p[16];
};

__global__ void peer_sum_kernel(PeerPtrs pp, int world, size_t n4, float4* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    float4 acc = __ldcg(pp.p[0] + i);
    for (int r = 1; r < world; ++r) {
      const float4 v = __ldcg(pp.p[r] + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[i] = acc;
  }
}

// The optimiser step fused into the exchange: one launch reads every rank's gradient arena over
// NVLink (rank-order sum) and applies keras Adam to all weight tensors of the network -- the summed
// gradient never goes back to HBM and the ~100 per-tensor Adam launches disappear.
struct AdamSeg {
  float* p; float* m; float* v;
  unsigned long long off, n;     // position of this tensor's gradient in the arenas
};

struct PeerPtrs1 {
  const float* p[16];
};

__global__ void peer_sum_adam_kernel(PeerPtrs1 pp, int world, const AdamSeg* __restrict__ segs,
                                     float lr_t, float b1, float b2, float eps) {
  const AdamSeg sg = segs[blockIdx.x];     // (x = table entry: up to 2^31 - 1 of them)
  for (size_t i = blockIdx.y * (size_t)blockDim.x + threadIdx.x; i < sg.n;
       i += (size_t)gridDim.y * blockDim.x) {
    float g = __ldcg(pp.p[0] + sg.off + i);
    for (int r = 1; r < world; ++r) g += __ldcg(pp.p[r] + sg.off + i);
    float pi = sg.p[i], mi = sg.m[i], vi = sg.v[i];
    adam_update(pi, mi, vi, g, lr_t, b1, b2, eps);
    sg.m[i] = mi;
    sg.v[i] = vi;
    sg.p[i] = pi;
  }
}

}  // namespace s3

extern "C" int s3_qdm_bc(const float* data, const int* window, const float* params_oh,
                         const float* params_mh, const float* params_mf, const double* quantiles,
                         const float* tau_fut, const double* k_factor, int n_sites, int n_times,
                         int n_windows, int n_quantiles, int relative,
                         const double* delta_denom_zero, const double* delta_denom_min,
                         const double* delta_range, const double* out_range, float* out,
                         unsigned long long* n_bad, s3_stream stream) {
  S3_REQUIRE(data && window && params_oh && params_mh && params_mf && quantiles && out && n_bad,
             "s3_qdm_bc: null pointer");
  S3_REQUIRE(n_sites > 0 && n_times > 0 && n_windows > 0 && n_quantiles >= 2,
             "s3_qdm_bc: needs sites, times, windows > 0 and >= 2 quantiles");
  S3_REQUIRE((tau_fut == nullptr) == (k_factor == nullptr),
             "s3_qdm_bc: tau_fut and k_factor go together (PresRat) or are both NULL (QDM)");
  s3::QdmParams p{};
  p.n_sites = n_sites; p.n_times = n_times; p.n_win = n_windows; p.n_q = n_quantiles;
  p.relative = relative ? 1 : 0;
  if (delta_denom_zero) { p.has_zero = 1; p.denom_zero = *delta_denom_zero; }
  if (delta_denom_min) { p.has_min = 1; p.denom_min = *delta_denom_min; }
  if (delta_range) {
    p.has_range = 1;
    p.delta_lo = delta_range[0] < delta_range[1] ? delta_range[0] : delta_range[1];
    p.delta_hi = delta_range[0] < delta_range[1] ? delta_range[1] : delta_range[0];
  }
  if (out_range) {
    p.has_out_range = 1;
    p.out_lo = out_range[0] < out_range[1] ? out_range[0] : out_range[1];
    p.out_hi = out_range[0] < out_range[1] ? out_range[1] : out_range[0];
  }
  cudaStream_t st = as_stream(stream);
  S3_CUDA(cudaMemsetAsync(n_bad, 0, 2 * sizeof(unsigned long long), st));
  const size_t total = (size_t)n_sites * n_times;
  s3::qdm_kernel<<<s3::pgrid(total), 256, 0, st>>>(data, window, params_oh, params_mh, params_mf,
                                                   quantiles, tau_fut, k_factor, p, out, n_bad);
  S3_LAUNCH_CHECK("qdm_bc");
  return S3_OK;
}

extern "C" int s3_peer_sum_adam(const void* const* peer_ptrs, int world, const void* segs_dev,
                                int n_seg, unsigned long long max_n, float lr, float beta1,
                                float beta2, float eps, long long step, s3_stream stream) {
  S3_REQUIRE(peer_ptrs && segs_dev && world >= 1 && world <= 16 && n_seg >= 1 && step >= 1,
             "s3_peer_sum_adam: bad arguments");
  s3::PeerPtrs1 pp;
  for (int r = 0; r < world; ++r) {
    S3_REQUIRE(peer_ptrs[r] != nullptr, "s3_peer_sum_adam: peer pointer %d is null", r);
    pp.p[r] = static_cast<const float*>(peer_ptrs[r]);
  }
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) /
                      (1.0 - pow((double)beta1, (double)step));
  unsigned bx = (unsigned)((max_n + 255) / 256);
  unsigned cap = (unsigned)(sm_count() * 16 / (n_seg < 1 ? 1 : n_seg)) + 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  if (bx > 65535u) bx = 65535u;
  dim3 grid((unsigned)n_seg, bx);
  s3::peer_sum_adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(
      pp, world, static_cast<const s3::AdamSeg*>(segs_dev), (float)lr_t, beta1, beta2, eps);
  S3_LAUNCH_CHECK("peer_sum_adam");
  return S3_OK;
}

extern "C" int s3_peer_sum_f32(const void* const* peer_ptrs, int world, size_t n, float* out,
                               s3_stream stream) {
  S3_REQUIRE(peer_ptrs && out && world >= 1 && world <= 16, "s3_peer_sum_f32: 1 <= world <= 16");
  S3_REQUIRE(n % 4 == 0, "s3_peer_sum_f32: n must be a multiple of 4 floats (pad the arena)");
  s3::PeerPtrs pp;
  for (int r = 0; r < world; ++r) {
    S3_REQUIRE(peer_ptrs[r] && (reinterpret_cast<uintptr_t>(peer_ptrs[r]) & 15) == 0,
               "s3_peer_sum_f32: peer pointer %d is null or not 16-byte aligned", r);
    pp.p[r] = static_cast<const float4*>(peer_ptrs[r]);
  }
  const size_t n4 = n / 4;
  if (n4)
    s3::peer_sum_kernel<<<s3::pgrid(n4), 256, 0, as_stream(stream)>>>(
        pp, world, n4, reinterpret_cast<float4*>(out));
  S3_LAUNCH_CHECK("peer_sum");
  return S3_OK;
}
