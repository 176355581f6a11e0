// Shared host/device helpers for libsup3r_b200: error convention, the conv epilogue
// (bias / activation / residual / affine) and the output scatter (depth_to_space,
// depth_to_time + roll, nearest repeat, concat-by-stride, padded+mirrored 16-bit copies).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include "../../include/sup3r_b200.h"

namespace s3 {

// ----------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define S3_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      ::s3::set_error(__VA_ARGS__);  \
      return S3_ERR_INVALID;         \
    }                                \
  } while (0)

#define S3_CUDA(expr)                                                  \
  do {                                                                 \
    cudaError_t e__ = (expr);                                          \
    if (e__ != cudaSuccess) return ::s3::cuda_fail(e__, #expr);        \
  } while (0)

#define S3_LAUNCH_CHECK(name) S3_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(s3_stream s) { return reinterpret_cast<cudaStream_t>(s); }
int sm_count();

// ------------------------------------------------------------------ conv geometry (POD)
struct ConvGeom {
  int ndim, n;
  int in[3], cin, cout, k[3], st[3], pl[3], ph[3], pad_mode;
  int od[3];       // conv output extents
  int act;
  float alpha;
  // scatter
  int r, m, roll, rep[3];
  int fd[3];       // final (mapped) extents
  int cmap;        // mapped channels per voxel (cout / (r*r*m))
  int cstride, coff;
  int ctotal, cbase;   // channel slice of a wider convolution (scatter map of the full one)
  int res_pre;         // f32 residual is added before the activation
};

int make_geom(const s3_conv_desc* d, ConvGeom* g);  // validates; returns S3_OK or error

struct Epilogue {
  const float* bias;
  const float* residual;
  const float* post_scale;
  const float* post_shift;
  float* y;          // f32 mapped destination (nullable)
  void* y_hi;        // 16-bit padded mirrored destination (nullable)
  void* y_lo;        // residual half of the split (nullable)
  int fmt;           // 0 bf16, 1 fp16, 2 fp16 + e4m3 correction rows ("fp16c", see below)
  const void* res_hi;  // residual as a 16-bit padded pair (hi + lo), same layout as y_hi
  const void* res_lo;
  float acc_scale;   // accumulator -> value factor (fp16c weights carry a power-of-two scale)
};

// ---- "fp16c" operand format (fmt 2) --------------------------------------------------------
// An activation tensor is a PAIR of padded tensors with 128 bytes per 64-channel voxel:
//   hi   : fp16(x)                                                     (64 x 2 B)
//   corr : per 32-channel half h = 0, 1 (64 bytes each):
//          [ lo8[32] = e4m3((x - hi) * 2^11) | a8[32] = e4m3(x) ]      (e4m3, 1 B each)
// A packed weight tensor is the matching pair (rows = output channels, scaled by the per-layer
// power of two S):  hi = fp16(w S);  corr = [ e4m3(w S 2^-11) | e4m3(w S - hi) ] per half.
// One kind::f16 pass over (hi, hi) plus one kind::f8f6f4 pass over (corr, corr) -- K = 128
// e4m3 elements per voxel at twice the fp16 rate, i.e. the time of one fp16 pass -- accumulates
//   S (hi_x hi_w + lo_x w + x lo_w)  ~  S x w   to ~2^-15 relative,
// and hi + lo8 2^-11 (15 significant bits) is the value a SkipConnection adds.
constexpr int kFmtBf16 = 0, kFmtFp16 = 1, kFmtFp16c = 2;
constexpr float kCorrScale = 2048.f, kCorrInv = 1.f / 2048.f;

// ------------------------------------------------------------------------ device side
__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  switch (act) {
    case S3_ACT_RELU: return v > 0.f ? v : 0.f;
    case S3_ACT_LEAKY: return v >= 0.f ? v : alpha * v;
    case S3_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case S3_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// out-of-line copy for hot epilogues: keeps tanh / sigmoid code out of their unrolled loops
// (the inlined switch per element bloated the epilogues past the instruction cache: ncu showed
// `no_inst` stalls on the LeakyReLU compares of the depth_to_space head)
static __device__ __noinline__ float apply_act_slow(float v, int act, float alpha) {
  return apply_act(v, act, alpha);
}

// inline fast cases + out-of-line rest (for unrolled epilogue loops)
__device__ __forceinline__ float apply_act_lean(float v, int act, float alpha) {
  if (act == S3_ACT_NONE) return v;
  if (act == S3_ACT_LEAKY) return v >= 0.f ? v : alpha * v;
  if (act == S3_ACT_RELU) return fmaxf(v, 0.f);
  return apply_act_slow(v, act, alpha);
}

__device__ __forceinline__ uint16_t to16(float v, int fmt) {
  if (fmt == 0) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  if (fmt == kFmtFp16c) v = fminf(fmaxf(v, -65504.f), 65504.f);   // finite: lo stays meaningful
  return __half_as_ushort(__float2half_rn(v));
}
// two floats -> packed e4m3 pair (a in the low byte), saturating
__device__ __forceinline__ uint32_t e4m3x2(float a, float b) {
  return (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ uint32_t e4m3x4(float a, float b, float c, float d) {
  return e4m3x2(a, b) | (e4m3x2(c, d) << 16);
}
__device__ __forceinline__ float2 e4m3x2_to_f32(uint32_t u) {
  __half2_raw h = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(u & 0xffffu), __NV_E4M3);
  return __half22float2(*reinterpret_cast<__half2*>(&h));
}
__device__ __forceinline__ float e4m3_to_f32(uint8_t u) { return e4m3x2_to_f32(u).x; }
// saturating fp16 pair (a in the low half)
__device__ __forceinline__ uint32_t f16x2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// byte offset inside a 128-byte corr row: kind 0 = lo8 / w8, kind 1 = a8 / w_lo8
__host__ __device__ __forceinline__ int corr_byte(int kind, int ch) {
  return 64 * (ch >> 5) + 32 * kind + (ch & 31);
}
__device__ __forceinline__ float from16(uint16_t h, int fmt) {
  if (fmt == 0) return __bfloat162float(__ushort_as_bfloat16(h));
  return __half2float(__ushort_as_half(h));
}

// Destination coordinates of one conv-output element after the scatter.
struct Dest {
  int z, y, x, c;  // mapped voxel (before out_repeat) and channel
};

__device__ __forceinline__ Dest map_dest(const ConvGeom& g, int z, int y, int x, int c) {
  Dest d;
  int c0 = c + g.cbase, tt = 0, i = 0, j = 0;
  if (g.m > 1) {
    int cq = g.ctotal / g.m;
    tt = c0 / cq;
    c0 -= tt * cq;
  }
  if (g.r > 1) {
    int ij = c0 / g.cmap;
    c0 -= ij * g.cmap;
    i = ij / g.r;
    j = ij - i * g.r;
  }
  if (g.ndim == 3) {
    d.z = z * g.r + i;
    d.y = y * g.r + j;
    int xt = x * g.m + tt;
    if (g.m > 1 && g.roll != 0) {
      int T = g.od[2] * g.m;
      xt = (xt + g.roll) % T;
      if (xt < 0) xt += T;
    }
    d.x = xt;
  } else {
    d.z = 0;
    d.y = y * g.r + i;
    d.x = x * g.r + j;
  }
  d.c = c0;
  return d;
}

// Padded positions (interior + mirrored halo copies) of coordinate o along a dim of extent F.
// Returns count (1..3); positions are in padded coordinates (0..F+1).
__device__ __forceinline__ int mirror_positions(int o, int F, int (&pos)[3]) {
  int cnt = 0;
  pos[cnt++] = o + 1;
  if (o == 1) pos[cnt++] = 0;
  if (o == F - 2) pos[cnt++] = F + 1;
  return cnt;
}

// Store a run of `len` consecutive destination channels (same mapped voxel) starting at mapped
// channel d.c.  `v` holds the final values.  Handles out_repeat and the 16-bit mirrored copies.
template <int MAXLEN>
__device__ __forceinline__ void store_run(const ConvGeom& g, const Epilogue& ep, int b,
                                          const Dest& d, const float* v, int len) {
  const int FZ = g.fd[0], FY = g.fd[1], FX = g.fd[2];
  for (int rz = 0; rz < g.rep[0]; ++rz)
    for (int ry = 0; ry < g.rep[1]; ++ry)
      for (int rx = 0; rx < g.rep[2]; ++rx) {
        const int oz = d.z * g.rep[0] + rz, oy = d.y * g.rep[1] + ry, ox = d.x * g.rep[2] + rx;
        if (ep.y) {
          size_t off = ((((size_t)b * FZ + oz) * FY + oy) * FX + ox) * g.cstride + g.coff + d.c;
          float* dst = ep.y + off;
          if ((len & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < MAXLEN / 4; ++q)
              if (q * 4 < len)
                reinterpret_cast<float4*>(dst)[q] =
                    make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          } else {
#pragma unroll
            for (int q = 0; q < MAXLEN; ++q)
              if (q < len) dst[q] = v[q];
          }
        }
        if (ep.y_hi) {
          const int pz = (g.ndim == 3) ? 1 : 0;
          const int PZ = FZ + 2 * pz, PY = FY + 2, PX = FX + 2;
          int zs[3], ys[3], xs[3];
          int nz = 1;
          zs[0] = oz;
          if (pz) nz = mirror_positions(oz, FZ, zs);
          const int ny = mirror_positions(oy, FY, ys);
          const int nx = mirror_positions(ox, FX, xs);
          uint16_t hi[MAXLEN], lo[MAXLEN];
#pragma unroll
          for (int q = 0; q < MAXLEN; ++q) {
            float f = q < len ? v[q] : 0.f;
            hi[q] = to16(f, ep.fmt);
            lo[q] = to16(f - from16(hi[q], ep.fmt), ep.fmt);
          }
          for (int a = 0; a < nz; ++a)
            for (int bq = 0; bq < ny; ++bq)
              for (int cq = 0; cq < nx; ++cq) {
                size_t off =
                    ((((size_t)b * PZ + zs[a]) * PY + ys[bq]) * PX + xs[cq]) * g.cstride +
                    g.coff + d.c;
                uint16_t* dh = reinterpret_cast<uint16_t*>(ep.y_hi) + off;
                uint16_t* dl = ep.y_lo ? reinterpret_cast<uint16_t*>(ep.y_lo) + off : nullptr;
                if ((len & 7) == 0 && ((reinterpret_cast<uintptr_t>(dh) & 15) == 0)) {
#pragma unroll
                  for (int q = 0; q < MAXLEN / 8; ++q)
                    if (q * 8 < len) {
                      uint4 u;
                      u.x = hi[q * 8] | (uint32_t(hi[q * 8 + 1]) << 16);
                      u.y = hi[q * 8 + 2] | (uint32_t(hi[q * 8 + 3]) << 16);
                      u.z = hi[q * 8 + 4] | (uint32_t(hi[q * 8 + 5]) << 16);
                      u.w = hi[q * 8 + 6] | (uint32_t(hi[q * 8 + 7]) << 16);
                      reinterpret_cast<uint4*>(dh)[q] = u;
                      if (dl) {
                        u.x = lo[q * 8] | (uint32_t(lo[q * 8 + 1]) << 16);
                        u.y = lo[q * 8 + 2] | (uint32_t(lo[q * 8 + 3]) << 16);
                        u.z = lo[q * 8 + 4] | (uint32_t(lo[q * 8 + 5]) << 16);
                        u.w = lo[q * 8 + 6] | (uint32_t(lo[q * 8 + 7]) << 16);
                        reinterpret_cast<uint4*>(dl)[q] = u;
                      }
                    }
                } else {
#pragma unroll
                  for (int q = 0; q < MAXLEN; ++q)
                    if (q < len) {
                      dh[q] = hi[q];
                      if (dl) dl[q] = lo[q];
                    }
                }
              }
        }
      }
}

// Finish one conv-output element: bias, activation, residual, affine.
__device__ __forceinline__ float finish(const ConvGeom& g, const Epilogue& ep, float acc, int c,
                                        size_t conv_vox) {
  float v = acc;
  if (ep.bias) v += ep.bias[c];
  if (ep.residual && g.res_pre) v += ep.residual[conv_vox * g.cout + c];
  v = apply_act_lean(v, g.act, g.alpha);
  if (ep.residual && !g.res_pre) v += ep.residual[conv_vox * g.cout + c];
  if (ep.post_scale) v = v * ep.post_scale[c] + (ep.post_shift ? ep.post_shift[c] : 0.f);
  return v;
}

// keras Adam update of one element (one definition for the plain and the fused peer-sum kernels so
// that both round identically): lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) is computed on the host
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, float lr_t,
                                            float b1, float b2, float eps) {
  const float mi = m + (g - m) * (1.f - b1);
  const float vi = v + (g * g - v) * (1.f - b2);
  m = mi;
  v = vi;
  p -= lr_t * mi / (sqrtf(vi) + eps);
}

}  // namespace s3
