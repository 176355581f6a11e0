// Generic fp32 convolution on CUDA cores: implicit GEMM with register tiling.
// Handles every geometry the sup3r configs use (2-D / 3-D, stride 1 / 2, valid / same-zero /
// reflect / symmetric implicit padding, any Cin / Cout) with the fused epilogue of common.cuh.
// It is the exact-fp32 device path (first / last generator layers, discriminator, training)
// and the on-device parity anchor for the tcgen05 kernel.
//
//   forward : out[v, co]  = sum_{tap, ci} x[in(v, tap), ci] * w[tap, ci, co]
//   dgrad   : dx[u, ci]   = sum_{tap, co} dy[out(u, tap), co] * w[tap, ci, co]
//   wgrad   : dw[tap, ci, co] = sum_v x[in(v, tap), ci] * dy[v, co]
//
// Replaces keras Conv2D/Conv3D/Conv2DTranspose (+ tape.gradient through them) as called from
// sup3r/models/abstract.py:1081-1092, 1157-1165, 1230-1238.
#include "common.cuh"

namespace s3 {

__device__ __forceinline__ int fold_coord(int q, int n, int mode, bool* ok) {
  if (q >= 0 && q < n) return q;
  if (mode == S3_PAD_REFLECT) {
    if (q < 0) q = -q;
    if (q >= n) q = 2 * n - 2 - q;
  } else if (mode == S3_PAD_SYMMETRIC) {
    if (q < 0) q = -q - 1;
    if (q >= n) q = 2 * n - 1 - q;
  } else {
    *ok = false;
    return 0;
  }
  if (q < 0 || q >= n) *ok = false;
  return q;
}

enum { MODE_FWD = 0, MODE_DGRAD = 1 };

template <int BM, int BN, int BK, int TM, int TN, int MODE>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_direct_kernel(const ConvGeom g, const float* __restrict__ src, const float* __restrict__ w,
                   const Epilogue ep, float* __restrict__ dx) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int AS = BM + 4;
  __shared__ __align__(16) float As[BK][AS];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ long long soff[BM];
  __shared__ int svox[BM][4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  // FWD: rows are conv-output voxels; DGRAD: rows are input voxels.
  const int R0 = MODE == MODE_FWD ? g.od[0] : g.in[0];
  const int R1 = MODE == MODE_FWD ? g.od[1] : g.in[1];
  const int R2 = MODE == MODE_FWD ? g.od[2] : g.in[2];
  const long long M = (long long)g.n * R0 * R1 * R2;
  const int KC = MODE == MODE_FWD ? g.cin : g.cout;   // contraction channels
  const int NC = MODE == MODE_FWD ? g.cout : g.cin;   // produced channels
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  for (int i = tid; i < BM; i += NT) {
    long long mv = m0 + i;
    if (mv < M) {
      int x = (int)(mv % R2);
      long long t = mv / R2;
      int y = (int)(t % R1);
      t /= R1;
      int z = (int)(t % R0);
      int b = (int)(t / R0);
      svox[i][0] = b; svox[i][1] = z; svox[i][2] = y; svox[i][3] = x;
    } else {
      svox[i][0] = -1;
    }
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int ntaps = g.k[0] * g.k[1] * g.k[2];
  for (int tap = 0; tap < ntaps; ++tap) {
    const int kx = tap % g.k[2];
    const int ky = (tap / g.k[2]) % g.k[1];
    const int kz = tap / (g.k[2] * g.k[1]);
    __syncthreads();
    for (int i = tid; i < BM; i += NT) {
      long long off = -1;
      const int b = svox[i][0];
      if (b >= 0) {
        bool ok = true;
        int cz, cy, cx;
        if (MODE == MODE_FWD) {
          cz = fold_coord(svox[i][1] * g.st[0] + kz - g.pl[0], g.in[0], g.pad_mode, &ok);
          cy = fold_coord(svox[i][2] * g.st[1] + ky - g.pl[1], g.in[1], g.pad_mode, &ok);
          cx = fold_coord(svox[i][3] * g.st[2] + kx - g.pl[2], g.in[2], g.pad_mode, &ok);
          if (ok) off = ((((long long)b * g.in[0] + cz) * g.in[1] + cy) * g.in[2] + cx) * g.cin;
        } else {
          int qz = svox[i][1] + g.pl[0] - kz, qy = svox[i][2] + g.pl[1] - ky,
              qx = svox[i][3] + g.pl[2] - kx;
          ok = qz >= 0 && qy >= 0 && qx >= 0 && qz % g.st[0] == 0 && qy % g.st[1] == 0 &&
               qx % g.st[2] == 0;
          cz = qz / g.st[0]; cy = qy / g.st[1]; cx = qx / g.st[2];
          ok = ok && cz < g.od[0] && cy < g.od[1] && cx < g.od[2];
          if (ok) off = ((((long long)b * g.od[0] + cz) * g.od[1] + cy) * g.od[2] + cx) * g.cout;
        }
      }
      soff[i] = off;
    }
    __syncthreads();
    for (int c0 = 0; c0 < KC; c0 += BK) {
      for (int e = tid; e < BM * BK; e += NT) {
        const int m = e / BK, kk = e % BK;
        const long long off = soff[m];
        float v = 0.f;
        if (off >= 0 && c0 + kk < KC) v = __ldg(src + off + c0 + kk);
        As[kk][m] = v;
      }
      for (int e = tid; e < BK * BN; e += NT) {
        const int kk = e / BN, nn = e % BN;
        float v = 0.f;
        if (c0 + kk < KC && n0 + nn < NC) {
          if (MODE == MODE_FWD)
            v = __ldg(w + ((size_t)tap * g.cin + c0 + kk) * g.cout + n0 + nn);
          else
            v = __ldg(w + ((size_t)tap * g.cin + n0 + nn) * g.cout + c0 + kk);
        }
        Bs[kk][nn] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], bq[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) bq[j] = Bs[kk][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bq[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ------------------------------------------------------------------ epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int ml = ty * TM + i;
    const int b = svox[ml][0];
    if (b < 0) continue;
    const long long mv = m0 + ml;
    const int cbase = n0 + tx * TN;
    if (cbase >= NC) continue;
    if (MODE == MODE_DGRAD) {
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (cbase + j < NC) dx[(size_t)mv * g.cin + cbase + j] = acc[i][j];
      continue;
    }
    float v[TN];
    int len = 0;
#pragma unroll
    for (int j = 0; j < TN; ++j)
      if (cbase + j < NC) {
        v[j] = finish(g, ep, acc[i][j], cbase + j, (size_t)mv);
        len = j + 1;
      } else {
        v[j] = 0.f;
      }
    if (g.r == 1 && g.m == 1) {
      Dest d = map_dest(g, svox[ml][1], svox[ml][2], svox[ml][3], cbase);
      store_run<TN>(g, ep, b, d, v, len);
    } else {
      for (int j = 0; j < len; ++j) {
        Dest d = map_dest(g, svox[ml][1], svox[ml][2], svox[ml][3], cbase + j);
        float one[1] = {v[j]};
        store_run<1>(g, ep, b, d, one, 1);
      }
    }
  }
}

// wgrad: per tap GEMM  dw[tap] (cin x cout) += X_tap^T (cin x V) * dY (V x cout), split over V.
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_wgrad_kernel(const ConvGeom g, const float* __restrict__ x, const float* __restrict__ dy,
                  float* __restrict__ dw, int vox_per_split, size_t dw_split_stride) {
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ long long soff[BK];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int tiles_m = (g.cin + BM - 1) / BM;
  const int tap = blockIdx.x / tiles_m;
  const int m0 = (blockIdx.x % tiles_m) * BM;
  const int n0 = blockIdx.y * BN;
  const long long V = (long long)g.n * g.od[0] * g.od[1] * g.od[2];
  const long long v_begin = (long long)blockIdx.z * vox_per_split;
  long long v_end = v_begin + vox_per_split;
  if (v_end > V) v_end = V;
  const int kx = tap % g.k[2];
  const int ky = (tap / g.k[2]) % g.k[1];
  const int kz = tap / (g.k[2] * g.k[1]);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (long long v0 = v_begin; v0 < v_end; v0 += BK) {
    __syncthreads();
    if (tid < BK) {
      long long mv = v0 + tid;
      long long off = -1;
      if (mv < v_end) {
        int xo = (int)(mv % g.od[2]);
        long long t = mv / g.od[2];
        int yo = (int)(t % g.od[1]);
        t /= g.od[1];
        int zo = (int)(t % g.od[0]);
        int b = (int)(t / g.od[0]);
        bool ok = true;
        int cz = fold_coord(zo * g.st[0] + kz - g.pl[0], g.in[0], g.pad_mode, &ok);
        int cy = fold_coord(yo * g.st[1] + ky - g.pl[1], g.in[1], g.pad_mode, &ok);
        int cx = fold_coord(xo * g.st[2] + kx - g.pl[2], g.in[2], g.pad_mode, &ok);
        if (ok) off = ((((long long)b * g.in[0] + cz) * g.in[1] + cy) * g.in[2] + cx) * g.cin;
      }
      soff[tid] = off;
    }
    __syncthreads();
    for (int e = tid; e < BK * BM; e += NT) {
      const int kk = e / BM, mm = e % BM;
      const long long off = soff[kk];
      float v = 0.f;
      if (off >= 0 && m0 + mm < g.cin) v = __ldg(x + off + m0 + mm);
      As[kk][mm] = v;
    }
    for (int e = tid; e < BK * BN; e += NT) {
      const int kk = e / BN, nn = e % BN;
      float v = 0.f;
      if (v0 + kk < v_end && n0 + nn < g.cout) v = __ldg(dy + (size_t)(v0 + kk) * g.cout + n0 + nn);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], bq[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bq[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bq[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int ci = m0 + ty * TM + i, co = n0 + tx * TN + j;
      if (ci < g.cin && co < g.cout)
        dw[(size_t)blockIdx.z * dw_split_stride + ((size_t)tap * g.cin + ci) * g.cout + co] =
            acc[i][j];
    }
}

// dw[i] = sum over the voxel splits, fixed order (deterministic second stage of the split-K wgrad)
__global__ void wgrad_split_reduce_kernel(const float* __restrict__ part, int splits, size_t n,
                                          float* __restrict__ dw) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(size_t)k * n + i];
    dw[i] = s;
  }
}

int launch_colsum(const float* x, long long rows, int cols, float* out, cudaStream_t st);
int try_conv_small(const ConvGeom& g, const float* x, const float* w, const Epilogue& ep,
                   cudaStream_t st);

int try_conv_small_mma(const ConvGeom& g, const float* x, const void* x16, const float* w,
                       const Epilogue& ep, cudaStream_t st);

}  // namespace s3

using namespace s3;

extern "C" int s3_conv_fwd_small_bf16(const s3_conv_desc* d, const float* x, const void* x_bf16,
                                      const float* w, const float* bias, const float* residual,
                                      const float* post_scale, const float* post_shift, float* y,
                                      s3_stream stream) {
  ConvGeom g;
  int rc = make_geom(d, &g);
  if (rc) return rc;
  S3_REQUIRE((x != nullptr) != (x_bf16 != nullptr) && w && y,
             "s3_conv_fwd_small_bf16: give x (f32) OR x_bf16, and weight / destination");
  S3_REQUIRE(!x_bf16 || g.cin == 8, "s3_conv_fwd_small_bf16: bf16 input needs exactly 8 channels");
  Epilogue ep{bias, residual, post_scale, post_shift, y, nullptr, nullptr, 0};
  rc = try_conv_small_mma(g, x, x_bf16, w, ep, as_stream(stream));
  if (rc < 0) return rc;
  S3_REQUIRE(rc == 1, "s3_conv_fwd_small_bf16: needs a 3-D 3x3x3 stride-1 pad-1 convolution with "
             "cin <= 8 and cout <= 8 and a plain output map");
  return S3_OK;
}

extern "C" int s3_conv_fwd_small_fp16(const s3_conv_desc* d, const float* x, const void* x_fp16,
                                      const float* w, const float* bias, const float* residual,
                                      const float* post_scale, const float* post_shift, float* y,
                                      s3_stream stream) {
  ConvGeom g;
  int rc = make_geom(d, &g);
  if (rc) return rc;
  S3_REQUIRE((x != nullptr) != (x_fp16 != nullptr) && w && y,
             "s3_conv_fwd_small_fp16: give x (f32) OR x_fp16, and weight / destination");
  S3_REQUIRE(!x_fp16 || g.cin == 8, "s3_conv_fwd_small_fp16: fp16 input needs exactly 8 channels");
  Epilogue ep{bias, residual, post_scale, post_shift, y, nullptr, nullptr, kFmtFp16};
  rc = try_conv_small_mma(g, x, x_fp16, w, ep, as_stream(stream));
  if (rc < 0) return rc;
  S3_REQUIRE(rc == 1, "s3_conv_fwd_small_fp16: needs a 3-D 3x3x3 stride-1 pad-1 convolution with "
             "cin <= 8 and cout <= 8 and a plain output map");
  return S3_OK;
}

extern "C" int s3_conv_fwd_f32(const s3_conv_desc* d, const float* x, const float* w,
                               const float* bias, const float* residual, const float* post_scale,
                               const float* post_shift, float* y, void* y_hi, void* y_lo,
                               s3_stream stream) {
  ConvGeom g;
  int rc = make_geom(d, &g);
  if (rc) return rc;
  S3_REQUIRE(x && w, "s3_conv_fwd_f32: null input/weight");
  S3_REQUIRE(y || y_hi, "s3_conv_fwd_f32: no destination");
  S3_REQUIRE(!(y_lo && !y_hi), "s3_conv_fwd_f32: y_lo without y_hi");
  Epilogue ep{bias, residual, post_scale, post_shift, y, y_hi, y_lo, 0};
  const long long M = (long long)g.n * g.od[0] * g.od[1] * g.od[2];
  cudaStream_t st = as_stream(stream);
  const int small = try_conv_small(g, x, w, ep, st);
  if (small < 0) return small;
  if (small == 1) return S3_OK;
  if (g.cout > 16) {
    dim3 grid((unsigned)((M + 127) / 128), (g.cout + 63) / 64);
    conv_direct_kernel<128, 64, 8, 8, 4, MODE_FWD><<<grid, 256, 0, st>>>(g, x, w, ep, nullptr);
  } else {
    dim3 grid((unsigned)((M + 255) / 256), (g.cout + 15) / 16);
    conv_direct_kernel<256, 16, 8, 4, 4, MODE_FWD><<<grid, 256, 0, st>>>(g, x, w, ep, nullptr);
  }
  S3_LAUNCH_CHECK("conv_direct_kernel");
  return S3_OK;
}

extern "C" int s3_conv_dgrad_f32(const s3_conv_desc* d, const float* dy, const float* w,
                                 float* dx, s3_stream stream) {
  ConvGeom g;
  int rc = make_geom(d, &g);
  if (rc) return rc;
  S3_REQUIRE(dy && w && dx, "s3_conv_dgrad_f32: null pointer");
  if (g.pad_mode != S3_PAD_ZERO) {
    set_error("s3_conv_dgrad_f32: only zero/valid padding (run reflect pads as s3_pad_bwd)");
    return S3_ERR_UNSUPPORTED;
  }
  Epilogue ep{};
  const long long M = (long long)g.n * g.in[0] * g.in[1] * g.in[2];
  cudaStream_t st = as_stream(stream);
  if (g.cin > 16) {
    dim3 grid((unsigned)((M + 127) / 128), (g.cin + 63) / 64);
    conv_direct_kernel<128, 64, 8, 8, 4, MODE_DGRAD><<<grid, 256, 0, st>>>(g, dy, w, ep, dx);
  } else {
    dim3 grid((unsigned)((M + 255) / 256), (g.cin + 15) / 16);
    conv_direct_kernel<256, 16, 8, 4, 4, MODE_DGRAD><<<grid, 256, 0, st>>>(g, dy, w, ep, dx);
  }
  S3_LAUNCH_CHECK("conv_dgrad_kernel");
  return S3_OK;
}

static long long wgrad_splits(const ConvGeom& g, int* vps_out) {
  const long long V = (long long)g.n * g.od[0] * g.od[1] * g.od[2];
  const int ntaps = g.k[0] * g.k[1] * g.k[2];
  // enough splits to fill the machine, at least 256 voxels each
  const int tiles = ntaps * ((g.cin + 63) / 64) * ((g.cout + 63) / 64);
  long long want = (4LL * sm_count() + tiles - 1) / tiles;
  long long max_splits = (V + 255) / 256;
  long long splits = want < 1 ? 1 : (want > max_splits ? max_splits : want);
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  int vps = (int)((V + splits - 1) / splits);
  vps = (vps + 15) / 16 * 16;
  if (vps < 16) vps = 16;
  splits = (V + vps - 1) / vps;
  if (splits < 1) splits = 1;
  *vps_out = vps;
  return splits;
}

/* Scratch of the deterministic split-K weight gradient: one partial dw per voxel split. */
extern "C" size_t s3_conv_wgrad_scratch_bytes(const s3_conv_desc* d) {
  ConvGeom g;
  if (make_geom(d, &g)) return 0;
  int vps;
  const long long splits = wgrad_splits(g, &vps);
  const size_t n = (size_t)g.k[0] * g.k[1] * g.k[2] * g.cin * g.cout;
  return splits > 1 ? (size_t)splits * n * sizeof(float) : 0;
}

extern "C" int s3_conv_wgrad_f32(const s3_conv_desc* d, const float* x, const float* dy,
                                 float* dw, float* dbias, void* scratch, s3_stream stream) {
  ConvGeom g;
  int rc = make_geom(d, &g);
  if (rc) return rc;
  S3_REQUIRE(dy && (dw == nullptr || x), "s3_conv_wgrad_f32: null pointer");
  cudaStream_t st = as_stream(stream);
  const long long V = (long long)g.n * g.od[0] * g.od[1] * g.od[2];
  const int ntaps = g.k[0] * g.k[1] * g.k[2];
  if (dw) {
    int vps;
    const long long splits = wgrad_splits(g, &vps);
    const size_t n = (size_t)ntaps * g.cin * g.cout;
    S3_REQUIRE(splits == 1 || scratch, "s3_conv_wgrad_f32: this geometry splits the voxels %lld "
               "ways and needs s3_conv_wgrad_scratch_bytes() of scratch", splits);
    dim3 grid(ntaps * ((g.cin + 63) / 64), (g.cout + 63) / 64, (unsigned)splits);
    float* dst = splits == 1 ? dw : static_cast<float*>(scratch);
    conv_wgrad_kernel<64, 64, 16, 4, 4><<<grid, 256, 0, st>>>(g, x, dy, dst, vps,
                                                                splits == 1 ? 0 : n);
    S3_LAUNCH_CHECK("conv_wgrad_kernel");
    if (splits > 1) {
      unsigned blocks = (unsigned)((n + 255) / 256);
      if (blocks > 4096u) blocks = 4096u;
      wgrad_split_reduce_kernel<<<blocks, 256, 0, st>>>(dst, (int)splits, n, dw);
      S3_LAUNCH_CHECK("wgrad_split_reduce_kernel");
    }
  }
  if (dbias) {
    rc = launch_colsum(dy, V, g.cout, dbias, st);
    if (rc) return rc;
  }
  return S3_OK;
}
