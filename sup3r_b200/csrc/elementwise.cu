// HBM-bound layer kernels of the eager path and the training step: padding, cropping,
// activations, skip add, depth_to_space / depth_to_time / nearest expansion (+ adjoints),
// channel concat, per-channel affine, 16-bit padded-activation and weight packing for the
// tcgen05 kernel, content / adversarial losses, Adam, tensor statistics.
// All are grid-stride, coalesced on the channel-fastest axis, grids sized to the SM count.
#include "common.cuh"

namespace s3 {

static inline unsigned grid_for(size_t n, int threads = 256) {
  size_t blocks = (n + threads - 1) / threads;
  size_t cap = (size_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

struct Shape5 {
  int d[5], lo[5], hi[5];
};

__device__ __forceinline__ int fold_pad(int q, int n, int mode, bool* ok) {
  if (q >= 0 && q < n) return q;
  if (mode == S3_PAD_REFLECT) {
    q = q < 0 ? -q : 2 * n - 2 - q;
  } else if (mode == S3_PAD_SYMMETRIC) {
    q = q < 0 ? -q - 1 : 2 * n - 1 - q;
  } else {
    *ok = false;
    return 0;
  }
  if (q < 0 || q >= n) *ok = false;
  return q;
}

__global__ void pad_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, Shape5 s,
                               int mode, size_t total) {
  int od[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) od[i] = s.d[i] + s.lo[i] + s.hi[i];
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    int c[5];
#pragma unroll
    for (int i = 4; i >= 0; --i) {
      c[i] = (int)(t % od[i]);
      t /= od[i];
    }
    bool ok = true;
    size_t src = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      int q = fold_pad(c[i] - s.lo[i], s.d[i], mode, &ok);
      src = src * s.d[i] + q;
    }
    y[idx] = ok ? x[src] : 0.f;
  }
}

// adjoint of pad: gather every padded position that reads input element idx
__global__ void pad_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, Shape5 s,
                               int mode, size_t total) {
  int od[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) od[i] = s.d[i] + s.lo[i] + s.hi[i];
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    int c[5];
#pragma unroll
    for (int i = 4; i >= 0; --i) {
      c[i] = (int)(t % s.d[i]);
      t /= s.d[i];
    }
    int pos[5][3], cnt[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      int n = s.d[i], k = 0;
      pos[i][k++] = c[i] + s.lo[i];
      if (mode == S3_PAD_REFLECT) {
        if (c[i] >= 1 && c[i] <= s.lo[i]) pos[i][k++] = s.lo[i] - c[i];
        if (c[i] <= n - 2 && c[i] >= n - 1 - s.hi[i]) pos[i][k++] = s.lo[i] + 2 * n - 2 - c[i];
      } else if (mode == S3_PAD_SYMMETRIC) {
        if (c[i] <= s.lo[i] - 1) pos[i][k++] = s.lo[i] - 1 - c[i];
        if (c[i] >= n - s.hi[i]) pos[i][k++] = s.lo[i] + 2 * n - 1 - c[i];
      }
      cnt[i] = k;
    }
    float acc = 0.f;
    for (int a = 0; a < cnt[0]; ++a)
      for (int b = 0; b < cnt[1]; ++b)
        for (int cc = 0; cc < cnt[2]; ++cc)
          for (int e = 0; e < cnt[3]; ++e)
            for (int f = 0; f < cnt[4]; ++f) {
              size_t o = pos[0][a];
              o = o * od[1] + pos[1][b];
              o = o * od[2] + pos[2][cc];
              o = o * od[3] + pos[3][e];
              o = o * od[4] + pos[4][f];
              acc += dy[o];
            }
    dx[idx] = acc;
  }
}

// Same adjoint for an unpadded channel axis with 4 | channels: one float4 per thread, the halo
// positions of the four outer axes enumerated in the same order (bit-identical sums).
__global__ void pad_bwd_vec4_kernel(const float4* __restrict__ dy, float4* __restrict__ dx,
                                    Shape5 s, int mode, size_t total4) {
  int od[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) od[i] = s.d[i] + s.lo[i] + s.hi[i];
  const int c4 = s.d[4] >> 2;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total4;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    const int ch = (int)(t % c4); t /= c4;
    int c[4];
#pragma unroll
    for (int i = 3; i >= 0; --i) {
      c[i] = (int)(t % s.d[i]);
      t /= s.d[i];
    }
    int pos[4][3], cnt[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = s.d[i];
      int k = 0;
      pos[i][k++] = c[i] + s.lo[i];
      if (mode == S3_PAD_REFLECT) {
        if (c[i] >= 1 && c[i] <= s.lo[i]) pos[i][k++] = s.lo[i] - c[i];
        if (c[i] <= n - 2 && c[i] >= n - 1 - s.hi[i]) pos[i][k++] = s.lo[i] + 2 * n - 2 - c[i];
      } else if (mode == S3_PAD_SYMMETRIC) {
        if (c[i] <= s.lo[i] - 1) pos[i][k++] = s.lo[i] - 1 - c[i];
        if (c[i] >= n - s.hi[i]) pos[i][k++] = s.lo[i] + 2 * n - 1 - c[i];
      }
      cnt[i] = k;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (a >= cnt[0]) break;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        if (b >= cnt[1]) break;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          if (cc >= cnt[2]) break;
#pragma unroll
          for (int e = 0; e < 3; ++e) {
            if (e >= cnt[3]) break;
            size_t o = pos[0][a];
            o = o * od[1] + pos[1][b];
            o = o * od[2] + pos[2][cc];
            o = o * od[3] + pos[3][e];
            const float4 v = dy[o * c4 + ch];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
        }
      }
    }
    dx[idx] = acc;
  }
}

__global__ void crop_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, Shape5 s,
                                size_t total) {
  int od[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) od[i] = s.d[i] - s.lo[i] - s.hi[i];
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    int c[5];
#pragma unroll
    for (int i = 4; i >= 0; --i) {
      c[i] = (int)(t % od[i]);
      t /= od[i];
    }
    size_t src = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) src = src * s.d[i] + c[i] + s.lo[i];
    y[idx] = x[src];
  }
}

__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n,
                               int act, float alpha) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] = apply_act(x[i], act, alpha);
}

__global__ void act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                               float* __restrict__ dx, size_t n, int act, float alpha) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float o = y[i], g = dy[i];
    float d;
    switch (act) {
      case S3_ACT_RELU: d = o > 0.f ? g : 0.f; break;
      case S3_ACT_LEAKY: d = o >= 0.f ? g : alpha * g; break;
      case S3_ACT_SIGMOID: d = g * o * (1.f - o); break;
      case S3_ACT_TANH: d = g * (1.f - o * o); break;
      default: d = g;
    }
    dx[i] = d;
  }
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                           float* __restrict__ y, size_t n, size_t nb) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] = a[i] + b[nb == n ? i : i % nb];
}

struct ExpandGeom {
  int ndim, n, d[3], c, r, m, method, roll;
  int od[3], oc;
};

// element of the expansion INPUT that lands on output element (b, oz, oy, ox, c0)
__device__ __forceinline__ size_t expand_src(const ExpandGeom& g, int b, int oz, int oy, int ox,
                                             int c0) {
  int z, y, x, i, j, tt = 0;
  if (g.ndim == 3) {
    z = oz / g.r; i = oz % g.r;
    y = oy / g.r; j = oy % g.r;
    if (g.m > 1 && g.method == 1) {
      int T = g.d[2] * g.m;
      int xt = (ox - g.roll) % T;
      if (xt < 0) xt += T;
      x = xt / g.m;
      tt = xt % g.m;
    } else {
      x = ox / g.m;
    }
  } else {
    z = 0; i = oy % g.r; y = oy / g.r; j = ox % g.r; x = ox / g.r;
  }
  const int cq = (g.m > 1 && g.method == 1) ? g.c / g.m : g.c;
  const int c = tt * cq + (i * g.r + j) * g.oc + c0;
  return ((((size_t)b * g.d[0] + z) * g.d[1] + y) * g.d[2] + x) * g.c + c;
}

__global__ void expand_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                  ExpandGeom g, size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    int c0 = (int)(t % g.oc); t /= g.oc;
    int ox = (int)(t % g.od[2]); t /= g.od[2];
    int oy = (int)(t % g.od[1]); t /= g.od[1];
    int oz = (int)(t % g.od[0]);
    int b = (int)(t / g.od[0]);
    y[idx] = x[expand_src(g, b, oz, oy, ox, c0)];
  }
}

// adjoint: thread per INPUT element; gathers its (one or m) output images
__global__ void expand_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                  ExpandGeom g, size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    int c = (int)(t % g.c); t /= g.c;
    int x = (int)(t % g.d[2]); t /= g.d[2];
    int y = (int)(t % g.d[1]); t /= g.d[1];
    int z = (int)(t % g.d[0]);
    int b = (int)(t / g.d[0]);
    const bool d2t = g.m > 1 && g.method == 1;
    const int cq = d2t ? g.c / g.m : g.c;
    int tt = c / cq;
    int cr = c - tt * cq;
    int ij = cr / g.oc;
    int c0 = cr - ij * g.oc;
    int i = ij / g.r, j = ij % g.r;
    float acc = 0.f;
    if (g.ndim == 3) {
      int oz = z * g.r + i, oy = y * g.r + j;
      if (d2t) {
        int T = g.d[2] * g.m;
        int ox = (x * g.m + tt + g.roll) % T;
        if (ox < 0) ox += T;
        acc = dy[((((size_t)b * g.od[0] + oz) * g.od[1] + oy) * g.od[2] + ox) * g.oc + c0];
      } else {
        for (int k = 0; k < g.m; ++k)
          acc += dy[((((size_t)b * g.od[0] + oz) * g.od[1] + oy) * g.od[2] + x * g.m + k) * g.oc +
                    c0];
      }
    } else {
      int oy = y * g.r + i, ox = x * g.r + j;
      acc = dy[(((size_t)b * g.od[1] + oy) * g.od[2] + ox) * g.oc + c0];
    }
    dx[idx] = acc;
  }
}

__global__ void concat_fwd_kernel(const float* __restrict__ a, int ca, const float* __restrict__ b,
                                  int cb, float* __restrict__ y, size_t total) {
  const int ct = ca + cb;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t v = idx / ct;
    int c = (int)(idx - v * ct);
    y[idx] = c < ca ? a[v * ca + c] : b[v * cb + (c - ca)];
  }
}

__global__ void concat_bwd_kernel(const float* __restrict__ dy, float* __restrict__ da, int ca,
                                  float* __restrict__ db, int cb, size_t total) {
  const int ct = ca + cb;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t v = idx / ct;
    int c = (int)(idx - v * ct);
    if (c < ca) {
      if (da) da[v * ca + c] = dy[idx];
    } else if (db) {
      db[v * cb + (c - ca)] = dy[idx];
    }
  }
}

__global__ void channel_affine_kernel(const float* __restrict__ x, float* __restrict__ y,
                                      size_t total, int c, const float* __restrict__ scale,
                                      const float* __restrict__ shift) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int ch = (int)(idx % c);
    y[idx] = x[idx] * (scale ? scale[ch] : 1.f) + (shift ? shift[ch] : 0.f);
  }
}

// ------------------------------------------------- 16-bit padded activations / weights
__global__ void pack_act_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi,
                                uint16_t* __restrict__ lo, int pz, int n, int Z, int Y, int X,
                                int c, int fmt, size_t total, int halo_mode, int hw) {
  const int PZ = Z + 2 * hw * pz, PY = Y + 2 * hw, PX = X + 2 * hw;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    int ch = (int)(t % c); t /= c;
    int px = (int)(t % PX); t /= PX;
    int py = (int)(t % PY); t /= PY;
    int pzc = (int)(t % PZ);
    int b = (int)(t / PZ);
    bool ok = true;
    int z = pz ? fold_pad(pzc - hw, Z, halo_mode, &ok) : pzc;
    int y = fold_pad(py - hw, Y, halo_mode, &ok);
    int xx = fold_pad(px - hw, X, halo_mode, &ok);
    float v = ok ? x[((((size_t)b * Z + z) * Y + y) * X + xx) * c + ch] : 0.f;
    uint16_t h = to16(v, fmt);
    hi[idx] = h;
    if (lo) {
      if (fmt == kFmtFp16c) {   // corr row (c == 64): e4m3 residue and e4m3 copy
        uint8_t* row = reinterpret_cast<uint8_t*>(lo) + (idx - ch) * 2;
        row[corr_byte(0, ch)] = (uint8_t)e4m3x2((v - from16(h, fmt)) * kCorrScale, 0.f);
        row[corr_byte(1, ch)] = (uint8_t)e4m3x2(v, 0.f);
      } else {
        lo[idx] = to16(v - from16(h, fmt), fmt);
      }
    }
  }
}

// Same packing, 8 channels per thread (c % 8 == 0): two float4 loads, one 16-byte store of the
// 16-bit values, 8-byte stores of the two e4m3 runs of the correction row (8 aligned channels are
// contiguous inside each run).  Element-wise arithmetic identical to pack_act_kernel.
__global__ void pack_act_vec8_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi,
                                     uint16_t* __restrict__ lo, int pz, int n, int Z, int Y,
                                     int X, int c, int fmt, size_t total8, int halo_mode,
                                     int hw) {
  const int PZ = Z + 2 * hw * pz, PY = Y + 2 * hw, PX = X + 2 * hw, c8 = c >> 3;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total8;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    const int ch = (int)(t % c8) * 8; t /= c8;
    const size_t vox = t;                      // padded voxel index
    int px = (int)(t % PX); t /= PX;
    int py = (int)(t % PY); t /= PY;
    int pzc = (int)(t % PZ);
    int b = (int)(t / PZ);
    bool ok = true;
    int z = pz ? fold_pad(pzc - hw, Z, halo_mode, &ok) : pzc;
    int y = fold_pad(py - hw, Y, halo_mode, &ok);
    int xx = fold_pad(px - hw, X, halo_mode, &ok);
    float v[8];
    if (ok) {
      const float4* src = reinterpret_cast<const float4*>(
          x + ((((size_t)b * Z + z) * Y + y) * X + xx) * c + ch);
      const float4 a = __ldg(src), d = __ldg(src + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      v[4] = d.x; v[5] = d.y; v[6] = d.z; v[7] = d.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    uint16_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = to16(v[i], fmt);
    uint4 hv;
    hv.x = h[0] | ((uint32_t)h[1] << 16); hv.y = h[2] | ((uint32_t)h[3] << 16);
    hv.z = h[4] | ((uint32_t)h[5] << 16); hv.w = h[6] | ((uint32_t)h[7] << 16);
    *reinterpret_cast<uint4*>(hi + vox * c + ch) = hv;
    if (lo) {
      if (fmt == kFmtFp16c) {   // (c == 64)
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = (v[i] - from16(h[i], fmt)) * kCorrScale;
        uint8_t* row = reinterpret_cast<uint8_t*>(lo) + vox * 128;
        *reinterpret_cast<uint2*>(row + corr_byte(0, ch)) =
            make_uint2(e4m3x4(r[0], r[1], r[2], r[3]), e4m3x4(r[4], r[5], r[6], r[7]));
        *reinterpret_cast<uint2*>(row + corr_byte(1, ch)) =
            make_uint2(e4m3x4(v[0], v[1], v[2], v[3]), e4m3x4(v[4], v[5], v[6], v[7]));
      } else {
        uint16_t l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) l[i] = to16(v[i] - from16(h[i], fmt), fmt);
        uint4 lv;
        lv.x = l[0] | ((uint32_t)l[1] << 16); lv.y = l[2] | ((uint32_t)l[3] << 16);
        lv.z = l[4] | ((uint32_t)l[5] << 16); lv.w = l[6] | ((uint32_t)l[7] << 16);
        *reinterpret_cast<uint4*>(lo + vox * c + ch) = lv;
      }
    }
  }
}

__global__ void unpack_act_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo,
                                  float* __restrict__ x, int pz, int n, int Z, int Y, int X, int c,
                                  int fmt, size_t total) {
  const int PZ = Z + 2 * pz, PY = Y + 2, PX = X + 2;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    int ch = (int)(t % c); t /= c;
    int xx = (int)(t % X); t /= X;
    int y = (int)(t % Y); t /= Y;
    int z = (int)(t % Z);
    int b = (int)(t / Z);
    size_t p = ((((size_t)b * PZ + z + pz) * PY + y + 1) * PX + xx + 1) * c + ch;
    float v = from16(hi[p], fmt);
    if (lo) {
      if (fmt == kFmtFp16c)
        v += e4m3_to_f32(reinterpret_cast<const uint8_t*>(lo)[(p - ch) * 2 + corr_byte(0, ch)]) *
             kCorrInv;
      else
        v += from16(lo[p], fmt);
    }
    x[idx] = v;
  }
}

// A view of a keras kernel (taps, src_cin, src_cout): rows ci0.. / columns co0.. (zero outside the
// tensor), or -- adjoint -- the spatially flipped kernel with input and output channels swapped
// (the operand of the input-gradient convolution), without materialising either.
struct WView {
  int src_cin, src_cout, ci0, co0, adjoint;
};

__global__ void pack_w_kernel(const float* __restrict__ w, uint16_t* __restrict__ hi,
                              uint16_t* __restrict__ lo, int taps, int cin, int cout, int npad,
                              int fmt, int layout, size_t total, float scale, WView v) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    int ci = (int)(t % cin); t /= cin;
    int co = (int)(t % npad);
    int tap = (int)(t / npad);
    if (layout == 1) {  // destination order (dy*3+dx, dz) -> source tap (dz*9 + dy*3 + dx)
      const int dz = tap % 3, dydx = tap / 3;
      tap = dz * 9 + dydx;
    }
    float val = 0.f;
    if (co < cout) {
      // source element: plain (tap, ci0 + ci, co0 + co); adjoint (taps-1-tap, co0 + co, ci0 + ci)
      const int st = v.adjoint ? taps - 1 - tap : tap;
      const int sci = v.adjoint ? v.co0 + co : v.ci0 + ci;
      const int sco = v.adjoint ? v.ci0 + ci : v.co0 + co;
      if (sci < v.src_cin && sco < v.src_cout)
        val = w[((size_t)st * v.src_cin + sci) * v.src_cout + sco];
    }
    val *= scale;
    uint16_t h = to16(val, fmt);
    hi[idx] = h;
    if (lo) {
      if (fmt == kFmtFp16c) {   // corr row (cin == 64): [e4m3(w S 2^-11) | e4m3(w S - hi)] per half
        uint8_t* row = reinterpret_cast<uint8_t*>(lo) + (idx - ci) * 2;
        row[corr_byte(0, ci)] = (uint8_t)e4m3x2(val * kCorrInv, 0.f);
        row[corr_byte(1, ci)] = (uint8_t)e4m3x2(val - from16(h, fmt), 0.f);
      } else {
        lo[idx] = to16(val - from16(h, fmt), fmt);
      }
    }
  }
}

// ------------------------------------------------------------------------------ losses
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sh[32];
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
  if (threadIdx.x < 32)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;  // valid in thread 0
}

__global__ void content_loss_kernel(const float* __restrict__ gen, const float* __restrict__ tru,
                                    size_t total, int c, int c_use, int kind, float weight,
                                    float inv_count, float* __restrict__ loss,
                                    float* __restrict__ dgen, float* __restrict__ scratch) {
  float acc = 0.f;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int ch = (int)(idx % c);
    float g = 0.f;
    if (ch < c_use) {
      float d = gen[idx] - tru[idx];
      if (kind == 0) {
        acc += d * d;
        g = 2.f * d * inv_count * weight;
      } else {
        acc += fabsf(d);
        g = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv_count * weight;
      }
    }
    if (dgen) dgen[idx] = g;
  }
  // DETERMINISTIC sum: every block stores its partial, the last block to finish (ticket in
  // scratch[0]) adds them in block order
  acc = block_sum(acc);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    scratch[1 + blockIdx.x] = acc;
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned*>(scratch), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float s = 0.f;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) s += __ldcg(scratch + 1 + b);
  s = block_sum(s);
  if (threadIdx.x == 0) *loss = s * inv_count;
}

__device__ __forceinline__ float sce_logits(float x, float z) {
  return fmaxf(x, 0.f) - x * z + log1pf(__expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// single block; b <= a few thousand logits
__global__ void loss_disc_kernel(const float* __restrict__ t, const float* __restrict__ f, int b,
                                 float weight, float* __restrict__ loss, float* __restrict__ dt,
                                 float* __restrict__ df) {
  __shared__ float sm[4];
  float st = 0.f, sf = 0.f;
  for (int i = threadIdx.x; i < b; i += blockDim.x) {
    st += t[i];
    sf += f[i];
  }
  st = block_sum(st);
  if (threadIdx.x == 0) sm[0] = st / b;
  sf = block_sum(sf);
  if (threadIdx.x == 0) sm[1] = sf / b;
  __syncthreads();
  const float tbar = sm[0], fbar = sm[1];
  float l = 0.f, sa = 0.f, sb = 0.f;
  for (int i = threadIdx.x; i < b; i += blockDim.x) {
    float xt = t[i] - fbar, xf = f[i] - tbar;
    l += sce_logits(xt, 1.f) + sce_logits(xf, 0.f);
    sa += sigmoidf_(xt) - 1.f;
    sb += sigmoidf_(xf);
  }
  l = block_sum(l);
  if (threadIdx.x == 0) loss[0] = l / (2.f * b);
  sa = block_sum(sa);
  if (threadIdx.x == 0) sm[2] = sa;
  sb = block_sum(sb);
  if (threadIdx.x == 0) sm[3] = sb;
  __syncthreads();
  const float inv = weight / (2.f * b);
  for (int i = threadIdx.x; i < b; i += blockDim.x) {
    float a_i = sigmoidf_(t[i] - fbar) - 1.f, b_i = sigmoidf_(f[i] - tbar);
    if (dt) dt[i] = inv * (a_i - sm[3] / b);
    if (df) df[i] = inv * (b_i - sm[2] / b);
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, size_t n, float lr_t,
                            float b1, float b2, float eps) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    adam_update(pi, mi, vi, g[i], lr_t, b1, b2, eps);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
  }
}

__global__ void stats_kernel(const float* __restrict__ x, size_t n, float* __restrict__ out) {
  float s = 0.f, sa = 0.f, bad = 0.f, mn = INFINITY, mx = -INFINITY;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float v = x[i];
    if (isfinite(v)) {
      s += v;
      sa += fabsf(v);
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    } else {
      bad += 1.f;
    }
  }
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(out + 0, s);
  sa = block_sum(sa);
  if (threadIdx.x == 0) atomicAdd(out + 1, sa);
  bad = block_sum(bad);
  if (threadIdx.x == 0) atomicAdd(out + 2, bad);
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    // float atomics via int ordering tricks: values compared as ordered ints
    int imn = __float_as_int(mn), imx = __float_as_int(mx);
    if (mn >= 0.f) atomicMin(reinterpret_cast<int*>(out + 3), imn);
    else atomicMax(reinterpret_cast<unsigned*>(out + 3), (unsigned)imn);
    if (mx >= 0.f) atomicMax(reinterpret_cast<int*>(out + 4), imx);
    else atomicMin(reinterpret_cast<unsigned*>(out + 4), (unsigned)imx);
  }
}

__global__ void stats_init_kernel(float* out) {
  out[0] = out[1] = out[2] = 0.f;
  out[3] = INFINITY;
  out[4] = -INFINITY;
}

// per-channel (min, max, nan count); one block per channel chunk, loops voxels
__global__ void channel_check_kernel(const float* __restrict__ x, size_t nvox, int c,
                                     float* __restrict__ out) {
  const int ch = blockIdx.x;
  float mn = INFINITY, mx = -INFINITY, bad = 0.f;
  for (size_t v = blockIdx.y * (size_t)blockDim.x + threadIdx.x; v < nvox;
       v += (size_t)gridDim.y * blockDim.x) {
    float val = x[v * c + ch];
    if (isnan(val)) bad += 1.f;
    else {
      mn = fminf(mn, val);
      mx = fmaxf(mx, val);
    }
  }
  __shared__ float smn[32], smx[32];
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    smn[threadIdx.x >> 5] = mn;
    smx[threadIdx.x >> 5] = mx;
  }
  bad = block_sum(bad);
  if (threadIdx.x == 0) {
    for (int i = 1; i < (blockDim.x >> 5); ++i) {
      mn = fminf(mn, smn[i]);
      mx = fmaxf(mx, smx[i]);
    }
    // out rows are initialised to (+inf, -inf, 0) by the caller-side init kernel
    int imn = __float_as_int(mn), imx = __float_as_int(mx);
    if (mn >= 0.f) atomicMin(reinterpret_cast<int*>(out + ch * 3 + 0), imn);
    else atomicMax(reinterpret_cast<unsigned*>(out + ch * 3 + 0), (unsigned)imn);
    if (mx >= 0.f) atomicMax(reinterpret_cast<int*>(out + ch * 3 + 1), imx);
    else atomicMin(reinterpret_cast<unsigned*>(out + ch * 3 + 1), (unsigned)imx);
    atomicAdd(out + ch * 3 + 2, bad);
  }
}

__global__ void channel_check_init_kernel(float* out, int c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) {
    out[i * 3 + 0] = INFINITY;
    out[i * 3 + 1] = -INFINITY;
    out[i * 3 + 2] = 0.f;
  }
}

static int fill_shape(Shape5* s, const int32_t dims[5], const int32_t lo[5], const int32_t hi[5],
                      bool crop) {
  for (int i = 0; i < 5; ++i) {
    s->d[i] = dims[i];
    s->lo[i] = lo[i];
    s->hi[i] = hi[i];
    S3_REQUIRE(dims[i] > 0 && lo[i] >= 0 && hi[i] >= 0, "bad pad/crop shape on dim %d", i);
    if (crop) S3_REQUIRE(dims[i] - lo[i] - hi[i] > 0, "cropping removes dim %d entirely", i);
  }
  return S3_OK;
}

static int fill_expand(ExpandGeom* g, int ndim, int n, const int32_t dims[3], int c, int r, int m,
                       int method, int roll) {
  S3_REQUIRE(ndim == 2 || ndim == 3, "expand: ndim must be 2 or 3");
  S3_REQUIRE(r >= 1 && m >= 1 && n > 0 && c > 0, "expand: bad multipliers");
  S3_REQUIRE(ndim == 3 || m == 1, "expand: temporal_mult needs a 5-D tensor");
  g->ndim = ndim; g->n = n; g->c = c; g->r = r; g->m = m; g->method = method; g->roll = roll;
  for (int i = 0; i < 3; ++i) g->d[i] = dims[i];
  int cq = c;
  if (m > 1 && method == 1) {
    S3_REQUIRE(c % m == 0, "depth_to_time: channels %d not divisible by temporal_mult %d", c, m);
    cq = c / m;
  }
  S3_REQUIRE(cq % (r * r) == 0, "depth_to_space: channels %d not divisible by spatial_mult^2 %d",
             cq, r * r);
  g->oc = cq / (r * r);
  if (ndim == 3) {
    g->od[0] = dims[0] * r; g->od[1] = dims[1] * r; g->od[2] = dims[2] * m;
  } else {
    S3_REQUIRE(dims[0] == 1, "expand: 2-D tensors must have z extent 1");
    g->od[0] = 1; g->od[1] = dims[1] * r; g->od[2] = dims[2] * r;
  }
  return S3_OK;
}

}  // namespace s3

using namespace s3;

extern "C" int s3_pad_fwd(const float* x, float* y, const int32_t dims[5], const int32_t lo[5],
                          const int32_t hi[5], int mode, s3_stream stream) {
  Shape5 s;
  int rc = fill_shape(&s, dims, lo, hi, false);
  if (rc) return rc;
  S3_REQUIRE(x && y, "s3_pad_fwd: null pointer");
  size_t total = 1;
  for (int i = 0; i < 5; ++i) {
    if (mode == S3_PAD_REFLECT)
      S3_REQUIRE(lo[i] < dims[i] && hi[i] < dims[i], "REFLECT pad %d/%d >= extent %d", lo[i],
                 hi[i], dims[i]);
    if (mode == S3_PAD_SYMMETRIC)
      S3_REQUIRE(lo[i] <= dims[i] && hi[i] <= dims[i], "SYMMETRIC pad larger than extent");
    total *= (size_t)(dims[i] + lo[i] + hi[i]);
  }
  pad_fwd_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, y, s, mode, total);
  S3_LAUNCH_CHECK("pad_fwd");
  return S3_OK;
}

extern "C" int s3_pad_bwd(const float* dy, float* dx, const int32_t dims[5], const int32_t lo[5],
                          const int32_t hi[5], int mode, s3_stream stream) {
  Shape5 s;
  int rc = fill_shape(&s, dims, lo, hi, false);
  if (rc) return rc;
  S3_REQUIRE(dy && dx, "s3_pad_bwd: null pointer");
  size_t total = 1;
  for (int i = 0; i < 5; ++i) total *= (size_t)dims[i];
  const bool vec4 = s.lo[4] == 0 && s.hi[4] == 0 && s.d[4] % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
  if (vec4)
    pad_bwd_vec4_kernel<<<grid_for(total / 4), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx), s, mode, total / 4);
  else
    pad_bwd_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(dy, dx, s, mode, total);
  S3_LAUNCH_CHECK("pad_bwd");
  return S3_OK;
}

extern "C" int s3_crop_fwd(const float* x, float* y, const int32_t dims[5], const int32_t lo[5],
                           const int32_t hi[5], s3_stream stream) {
  Shape5 s;
  int rc = fill_shape(&s, dims, lo, hi, true);
  if (rc) return rc;
  S3_REQUIRE(x && y, "s3_crop_fwd: null pointer");
  size_t total = 1;
  for (int i = 0; i < 5; ++i) total *= (size_t)(dims[i] - lo[i] - hi[i]);
  crop_fwd_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, y, s, total);
  S3_LAUNCH_CHECK("crop_fwd");
  return S3_OK;
}

extern "C" int s3_crop_bwd(const float* dy, float* dx, const int32_t dims[5], const int32_t lo[5],
                           const int32_t hi[5], s3_stream stream) {
  int32_t cd[5];
  for (int i = 0; i < 5; ++i) cd[i] = dims[i] - lo[i] - hi[i];
  return s3_pad_fwd(dy, dx, cd, lo, hi, S3_PAD_ZERO, stream);
}

extern "C" int s3_act_fwd(const float* x, float* y, size_t n, int act, float alpha,
                          s3_stream stream) {
  S3_REQUIRE(x && y, "s3_act_fwd: null pointer");
  if (n == 0) return S3_OK;
  act_fwd_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(x, y, n, act, alpha);
  S3_LAUNCH_CHECK("act_fwd");
  return S3_OK;
}

extern "C" int s3_act_bwd(const float* y, const float* dy, float* dx, size_t n, int act,
                          float alpha, s3_stream stream) {
  S3_REQUIRE(y && dy && dx, "s3_act_bwd: null pointer");
  if (n == 0) return S3_OK;
  act_bwd_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(y, dy, dx, n, act, alpha);
  S3_LAUNCH_CHECK("act_bwd");
  return S3_OK;
}

extern "C" int s3_add(const float* a, const float* b, float* y, size_t n, size_t nb,
                      s3_stream stream) {
  S3_REQUIRE(a && b && y && nb > 0 && n % nb == 0, "s3_add: bad arguments");
  if (n == 0) return S3_OK;
  add_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(a, b, y, n, nb);
  S3_LAUNCH_CHECK("add");
  return S3_OK;
}

extern "C" int s3_expand_fwd(const float* x, float* y, int ndim, int n, const int32_t dims[3],
                             int c, int spatial_mult, int temporal_mult, int method, int t_roll,
                             s3_stream stream) {
  ExpandGeom g;
  int rc = fill_expand(&g, ndim, n, dims, c, spatial_mult, temporal_mult, method, t_roll);
  if (rc) return rc;
  S3_REQUIRE(x && y, "s3_expand_fwd: null pointer");
  size_t total = (size_t)n * g.od[0] * g.od[1] * g.od[2] * g.oc;
  expand_fwd_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, y, g, total);
  S3_LAUNCH_CHECK("expand_fwd");
  return S3_OK;
}

extern "C" int s3_expand_bwd(const float* dy, float* dx, int ndim, int n, const int32_t dims[3],
                             int c, int spatial_mult, int temporal_mult, int method, int t_roll,
                             s3_stream stream) {
  ExpandGeom g;
  int rc = fill_expand(&g, ndim, n, dims, c, spatial_mult, temporal_mult, method, t_roll);
  if (rc) return rc;
  S3_REQUIRE(dy && dx, "s3_expand_bwd: null pointer");
  size_t total = (size_t)n * dims[0] * dims[1] * dims[2] * c;
  expand_bwd_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(dy, dx, g, total);
  S3_LAUNCH_CHECK("expand_bwd");
  return S3_OK;
}

extern "C" int s3_concat_fwd(const float* a, int ca, const float* b, int cb, float* y,
                             size_t nvox, s3_stream stream) {
  S3_REQUIRE(a && b && y && ca > 0 && cb > 0, "s3_concat_fwd: bad arguments");
  size_t total = nvox * (size_t)(ca + cb);
  if (total == 0) return S3_OK;
  concat_fwd_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(a, ca, b, cb, y, total);
  S3_LAUNCH_CHECK("concat_fwd");
  return S3_OK;
}

extern "C" int s3_concat_bwd(const float* dy, float* da, int ca, float* db, int cb, size_t nvox,
                             s3_stream stream) {
  S3_REQUIRE(dy && ca > 0 && cb > 0, "s3_concat_bwd: bad arguments");
  size_t total = nvox * (size_t)(ca + cb);
  if (total == 0) return S3_OK;
  concat_bwd_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(dy, da, ca, db, cb, total);
  S3_LAUNCH_CHECK("concat_bwd");
  return S3_OK;
}

extern "C" int s3_channel_affine(const float* x, float* y, size_t nvox, int c, const float* scale,
                                 const float* shift, s3_stream stream) {
  S3_REQUIRE(x && y && c > 0, "s3_channel_affine: bad arguments");
  size_t total = nvox * (size_t)c;
  if (total == 0) return S3_OK;
  channel_affine_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, y, total, c, scale,
                                                                          shift);
  S3_LAUNCH_CHECK("channel_affine");
  return S3_OK;
}

extern "C" int s3_pack_act_pad16(const float* x, int ndim, int n, const int32_t dims[3], int c,
                                 void* hi, void* lo, int fmt, s3_stream stream) {
  return s3_pack_act_pad16_ex(x, ndim, n, dims, c, hi, lo, fmt, S3_PAD_REFLECT, stream);
}

extern "C" int s3_pack_act_pad16_ex(const float* x, int ndim, int n, const int32_t dims[3], int c,
                                    void* hi, void* lo, int fmt, int halo_mode,
                                    s3_stream stream) {
  return s3_pack_act_pad16_hw(x, ndim, n, dims, c, hi, lo, fmt, halo_mode, 1, stream);
}

extern "C" int s3_pack_act_pad16_hw(const float* x, int ndim, int n, const int32_t dims[3], int c,
                                    void* hi, void* lo, int fmt, int halo_mode, int halo_width,
                                    s3_stream stream) {
  S3_REQUIRE(x && hi && (ndim == 2 || ndim == 3), "s3_pack_act_pad16: bad arguments");
  S3_REQUIRE(halo_mode == S3_PAD_REFLECT || halo_mode == S3_PAD_ZERO,
             "s3_pack_act_pad16: halo_mode must be S3_PAD_REFLECT or S3_PAD_ZERO");
  S3_REQUIRE(halo_width == 1 || (halo_width == 2 && halo_mode == S3_PAD_ZERO),
             "s3_pack_act_pad16: halo_width 1, or 2 with a zero halo");
  S3_REQUIRE(fmt != kFmtFp16c || !lo || c == 64, "s3_pack_act_pad16: fp16c rows need c == 64");
  const int pz = ndim == 3 ? 1 : 0, hw = halo_width;
  S3_REQUIRE(halo_mode == S3_PAD_ZERO || (dims[1] >= 2 && dims[2] >= 2 && (!pz || dims[0] >= 2)),
             "s3_pack_act_pad16: reflect halo needs extents >= 2");
  size_t total = (size_t)n * (dims[0] + 2 * hw * pz) * (dims[1] + 2 * hw) * (dims[2] + 2 * hw) * c;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) |
                         reinterpret_cast<uintptr_t>(lo)) & 15) == 0;
  if (c % 8 == 0 && aligned)
    pack_act_vec8_kernel<<<grid_for(total / 8), 256, 0, as_stream(stream)>>>(
        x, (uint16_t*)hi, (uint16_t*)lo, pz, n, dims[0], dims[1], dims[2], c, fmt, total / 8,
        halo_mode, hw);
  else
    pack_act_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(
        x, (uint16_t*)hi, (uint16_t*)lo, pz, n, dims[0], dims[1], dims[2], c, fmt, total,
        halo_mode, hw);
  S3_LAUNCH_CHECK("pack_act");
  return S3_OK;
}

extern "C" int s3_unpack_act_pad16(const void* hi, const void* lo, int ndim, int n,
                                   const int32_t dims[3], int c, float* x, int fmt,
                                   s3_stream stream) {
  S3_REQUIRE(x && hi && (ndim == 2 || ndim == 3), "s3_unpack_act_pad16: bad arguments");
  S3_REQUIRE(fmt != kFmtFp16c || !lo || c == 64, "s3_unpack_act_pad16: fp16c rows need c == 64");
  const int pz = ndim == 3 ? 1 : 0;
  size_t total = (size_t)n * dims[0] * dims[1] * dims[2] * c;
  unpack_act_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(
      (const uint16_t*)hi, (const uint16_t*)lo, x, pz, n, dims[0], dims[1], dims[2], c, fmt,
      total);
  S3_LAUNCH_CHECK("unpack_act");
  return S3_OK;
}

__global__ void cast_f16_kernel(const float4* __restrict__ x, uint2* __restrict__ y, size_t n4,
                                const float* __restrict__ xs, uint16_t* __restrict__ ys,
                                size_t tail0, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    y[i] = make_uint2(f16x2_sat(v.x, v.y), f16x2_sat(v.z, v.w));
  }
  if (blockIdx.x == 0)
    for (size_t i = tail0 + threadIdx.x; i < n; i += blockDim.x)
      ys[i] = (uint16_t)(f16x2_sat(xs[i], 0.f) & 0xffffu);
}

extern "C" int s3_cast_f16(const float* x, void* y, size_t n, s3_stream stream) {
  S3_REQUIRE(x && y, "s3_cast_f16: null pointer");
  if (n == 0) return S3_OK;
  const size_t n4 = n / 4;
  cast_f16_kernel<<<grid_for(n4 ? n4 : 1), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(y), n4, x,
      reinterpret_cast<uint16_t*>(y), n4 * 4, n);
  S3_LAUNCH_CHECK("cast_f16");
  return S3_OK;
}

extern "C" int s3_umma_npad(int cout) { return (cout + 15) / 16 * 16; }

namespace s3 { struct WView; }
static int pack_weights_umma(const float* w, int taps, int cin, int cout, void* w_hi, void* w_lo,
                             int fmt, int layout, float scale, s3_stream stream,
                             const WView* view);

extern "C" int s3_pack_weights_umma(const float* w, int taps, int cin, int cout, void* w_hi,
                                    void* w_lo, int fmt, int layout, s3_stream stream) {
  S3_REQUIRE(fmt != kFmtFp16c, "s3_pack_weights_umma: fp16c weights carry a scale, use "
             "s3_pack_weights_umma_c");
  return pack_weights_umma(w, taps, cin, cout, w_hi, w_lo, fmt, layout, 1.f, stream, nullptr);
}

extern "C" int s3_pack_weights_umma_c(const float* w, int taps, int cin, int cout, void* w_hi,
                                      void* w_corr, float scale, int layout, s3_stream stream) {
  S3_REQUIRE(w_corr && scale > 0.f, "s3_pack_weights_umma_c: needs w_corr and a positive scale");
  return pack_weights_umma(w, taps, cin, cout, w_hi, w_corr, kFmtFp16c, layout, scale, stream,
                           nullptr);
}

static int pack_weights_umma(const float* w, int taps, int cin, int cout, void* w_hi, void* w_lo,
                             int fmt, int layout, float scale, s3_stream stream,
                             const WView* view) {
  S3_REQUIRE(layout == 0 || (layout == 1 && taps == 27),
             "s3_pack_weights_umma: layout 1 (zcat) needs 27 taps");
  S3_REQUIRE(w && w_hi && taps > 0 && cin == 64 && cout > 0 && cout <= 256,
             "s3_pack_weights_umma: needs cin == 64 and cout <= 256 (got %d, %d)", cin, cout);
  const int npad = s3_umma_npad(cout);
  size_t total = (size_t)taps * npad * cin;
  WView v{cin, cout, 0, 0, 0};
  if (view) v = *view;
  pack_w_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(
      w, (uint16_t*)w_hi, (uint16_t*)w_lo, taps, cin, cout, npad, fmt, layout, total, scale, v);
  S3_LAUNCH_CHECK("pack_w");
  return S3_OK;
}

extern "C" int s3_pack_weights_umma_view(const float* w, int taps, int src_cin, int src_cout,
                                         int ci0, int co0, int adjoint, int cout, void* w_hi,
                                         void* w_corr, float scale, int layout,
                                         s3_stream stream) {
  S3_REQUIRE(w_corr && scale > 0.f, "s3_pack_weights_umma_view: needs w_corr and a scale > 0");
  S3_REQUIRE(src_cin > 0 && src_cout > 0 && ci0 >= 0 && co0 >= 0,
             "s3_pack_weights_umma_view: bad view");
  WView v{src_cin, src_cout, ci0, co0, adjoint ? 1 : 0};
  return pack_weights_umma(w, taps, 64, cout, w_hi, w_corr, kFmtFp16c, layout, scale, stream, &v);
}

extern "C" int s3_content_loss(const float* gen, const float* truth, size_t nvox, int c, int c_use,
                               int kind, float weight, float* loss, float* dgen,
                               float* scratch, s3_stream stream) {
  S3_REQUIRE(gen && truth && loss && scratch && c > 0 && c_use > 0 && c_use <= c &&
                 (kind == 0 || kind == 1),
             "s3_content_loss: bad arguments");
  size_t total = nvox * (size_t)c;
  if (total == 0) {
    S3_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), as_stream(stream)));
    return S3_OK;
  }
  S3_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float), as_stream(stream)));   // the ticket
  float inv = 1.f / ((float)nvox * (float)c_use);
  unsigned blocks = grid_for(total);
  if (blocks > S3_LOSS_SCRATCH_FLOATS - 1) blocks = S3_LOSS_SCRATCH_FLOATS - 1;
  content_loss_kernel<<<blocks, 256, 0, as_stream(stream)>>>(gen, truth, total, c, c_use, kind,
                                                             weight, inv, loss, dgen, scratch);
  S3_LAUNCH_CHECK("content_loss");
  return S3_OK;
}

extern "C" int s3_loss_disc(const float* out_real, const float* out_fake, int b, float weight,
                            float* loss, float* d_real, float* d_fake, s3_stream stream) {
  S3_REQUIRE(out_real && out_fake && loss && b > 0, "s3_loss_disc: bad arguments");
  loss_disc_kernel<<<1, 256, 0, as_stream(stream)>>>(out_real, out_fake, b, weight, loss, d_real,
                                                     d_fake);
  S3_LAUNCH_CHECK("loss_disc");
  return S3_OK;
}

extern "C" int s3_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr,
                            float beta1, float beta2, float eps, int64_t step, s3_stream stream) {
  S3_REQUIRE(p && g && m && v && step >= 1, "s3_adam_step: bad arguments");
  if (n == 0) return S3_OK;
  // keras Adam: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  p -= lr_t * m / (sqrt(v) + eps)
  double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) /
                (1.0 - pow((double)beta1, (double)step));
  adam_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(p, g, m, v, n, (float)lr_t, beta1, beta2,
                                                          eps);
  S3_LAUNCH_CHECK("adam");
  return S3_OK;
}

extern "C" int s3_stats(const float* x, size_t n, float* out5, s3_stream stream) {
  S3_REQUIRE(x && out5, "s3_stats: null pointer");
  stats_init_kernel<<<1, 1, 0, as_stream(stream)>>>(out5);
  if (n) stats_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(x, n, out5);
  S3_LAUNCH_CHECK("stats");
  return S3_OK;
}

extern "C" int s3_channel_check(const float* x, size_t nvox, int c, float* out, s3_stream stream) {
  S3_REQUIRE(x && out && c > 0, "s3_channel_check: bad arguments");
  channel_check_init_kernel<<<(c + 255) / 256, 256, 0, as_stream(stream)>>>(out, c);
  if (nvox) {
    unsigned gy = (unsigned)((nvox + 255) / 256);
    unsigned cap = (unsigned)(sm_count() * 8 / (c < 1 ? 1 : c)) + 1;
    if (gy > cap) gy = cap;
    dim3 grid(c, gy);
    channel_check_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, nvox, c, out);
  }
  S3_LAUNCH_CHECK("channel_check");
  return S3_OK;
}
