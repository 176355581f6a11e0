// Shared pieces of the tcgen05 convolution kernels: parameters, work-item decoding, shared
// memory carve-up, CTA set-up and the per-tile epilogue driver.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "umma_epilogue.cuh"

namespace s3 {

struct UmmaParams {
  ConvGeom g;
  Epilogue ep;
  int kz, ntaps, npad, split, fmt;
  int npass;          // zring: operand passes per item (2 = fp16 pass + e4m3 corr pass, fp16c)
  int R, TS, XB, YB, ZB, WS, AS, acc_bufs;
  int flat;           // 0: plane mode (16-row y blocks), 1: flat mode (full padded height)
  int nxb, nyb;       // x blocks of 8, y blocks of 16 (plane mode)
  int groups_per_b;   // plane groups (plane mode) / flat items (flat mode) per batch entry
  int nb;             // batch entries looped as separate tensors (3-D: n; 2-D: 1)
  int planes;         // output planes per batch entry (3-D: Z; 2-D: N)
  int plane_pitch;    // padded planes per batch entry (3-D: Z+2; 2-D: 0)
  int n_items;
  uint32_t box_bytes, box_stride, w_bytes;  // per operand half (stride = 1 KiB aligned)
  uint32_t idesc;
  DebugRec* dbg;
  long long* trace;   // optional device buffer: per-role clock64 accumulators of CTA 0
  int dbg_flags;      // experiment flags (s3_umma_tuning.box_y / ring_slots, see include/sup3r_b200.h)
  int epi_v4;         // zring 16-bit epilogue: 0 thread-per-row (8 warps), 1 TMA tile I/O (16 warps)
  int epi_row_tma;    // V4: y-halo rows stored by TMA (needs X % 8 == 0)
  int tile_fast;      // tile kernel: straight-line MMA role (3-D plane mode, R = 2, XB = 10, WS = 4)
  int seq2;           // tile kernel: fp16c as two sequential passes per item through the same
                      // activation / weight stages (instead of both operand halves resident)
  int ring_fast;      // zring: every item is the hot shape (R = 4, npad = 64, XB = 10, P = 7, WS = 2)
};

struct ItemCoord {
  int xb, y0, b, pl0, row0;
};

__device__ __forceinline__ ItemCoord decode_item(const UmmaParams& p, int item) {
  ItemCoord c;
  c.xb = item % p.nxb;
  int rest = item / p.nxb;
  if (!p.flat) {
    int yb = rest % p.nyb;
    int pg = rest / p.nyb;
    c.b = pg / p.groups_per_b;
    c.pl0 = (pg % p.groups_per_b) * p.R;
    c.y0 = yb * 16;
    c.row0 = 0;
  } else {
    c.b = rest / p.groups_per_b;
    int f0 = (rest % p.groups_per_b) * p.R * 16;
    c.pl0 = f0 / p.YB;
    c.row0 = f0 % p.YB;
    c.y0 = 0;
  }
  return c;
}

constexpr int kThreads = 192;
constexpr int kTileThreads = 320;   // tile kernel: TMA + MMA + 8 epilogue warps
constexpr int kMaxWS = 8;
constexpr int kMaxPlanes = 10;

// barrier slot indices (8 B each) inside the barrier block
constexpr int B_AFULL = 0;                        // [2 stages][kMaxPlanes]
constexpr int B_AEMPTY = B_AFULL + 2 * kMaxPlanes;
constexpr int B_WFULL = B_AEMPTY + 2 * kMaxPlanes;
constexpr int B_WEMPTY = B_WFULL + kMaxWS;
constexpr int B_ACCFULL = B_WEMPTY + kMaxWS;
constexpr int B_ACCEMPTY = B_ACCFULL + 2;
constexpr int B_TMEMPTR = B_ACCEMPTY + 2;
constexpr int B_COUNT = B_TMEMPTR + 1;
static_assert(B_COUNT * 8 <= 1024, "barrier block overflows its 1 KiB");

struct SmemMap {
  uint32_t a_base, w_base, bar_base, a_stage_bytes, w_slab, w_stage_bytes;
  float* sbias;  // [npad] bias staged in shared memory (1 KiB after the barrier block)
};

__device__ __forceinline__ SmemMap carve(const UmmaParams& p, const uint8_t* smem_raw) {
  SmemMap m;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int halves = (p.split && !p.seq2) ? 2 : 1;
  m.a_stage_bytes = p.box_stride * halves;
  m.w_slab = (p.w_bytes + 1023u) & ~1023u;
  m.w_stage_bytes = m.w_slab * halves;
  m.a_base = base;
  m.w_base = m.a_base + m.a_stage_bytes * p.AS;
  m.bar_base = m.w_base + m.w_stage_bytes * p.WS;
  m.sbias = reinterpret_cast<float*>(const_cast<uint8_t*>(smem_raw) +
                                     (m.bar_base + 1024u - smem_u32(smem_raw)));
  return m;
}

__device__ __forceinline__ uint32_t setup_cta(const UmmaParams& p, const SmemMap& m,
                                              const CUtensorMap* a_hi, const CUtensorMap* a_lo,
                                              const CUtensorMap* w_hi, const CUtensorMap* w_lo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * kMaxPlanes; ++i) {
      mbar_init(m.bar_base + 8u * (B_AFULL + i), 1);
      mbar_init(m.bar_base + 8u * (B_AEMPTY + i), 1);
    }
    for (int i = 0; i < kMaxWS; ++i) {
      mbar_init(m.bar_base + 8u * (B_WFULL + i), 1);
      mbar_init(m.bar_base + 8u * (B_WEMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(m.bar_base + 8u * (B_ACCFULL + i), 1);
      mbar_init(m.bar_base + 8u * (B_ACCEMPTY + i), (blockDim.x >> 5) - 2);   // epilogue warps
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.npad; i += blockDim.x)
    m.sbias[i] = (p.ep.bias && i < p.g.cout) ? p.ep.bias[i] : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(a_hi);
    tma_prefetch_desc(w_hi);
    if (p.split) {
      tma_prefetch_desc(a_lo);
      tma_prefetch_desc(w_lo);
    }
  }
  if (warp == 1) {
    tmem_alloc(m.bar_base + 8u * B_TMEMPTR, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(m.bar_base + 8u * B_TMEMPTR));
  return tmem_base;
}

// One accumulator tile (128 voxels of one plane / flat row range) through the lean epilogue.
template <int EPI>
__device__ __forceinline__ void epilogue_tile(const UmmaParams& p, const SmemMap& sm,
                                              const ItemCoord& c, int fr0, uint32_t t_addr,
                                              int warp, int lane) {
  const ConvGeom& g = p.g;
  const int q = warp & 3;
  const int mrow = q * 32 + lane;
  const int grp = mrow >> 3, xl = mrow & 7;
  const int fr = fr0 + grp;
  const int zq = fr / p.YB, yq = fr - zq * p.YB;
  const int plane = c.pl0 + zq;
  RowPlan rp;
  rp.y = c.y0 + yq;
  rp.x = c.xb * 8 + xl;
  rp.valid = yq <= p.YB - 3 && rp.y < g.in[1] && rp.x < g.in[2] && plane < p.planes;
  if (g.ndim == 3) { rp.b = c.b; rp.z = plane; } else { rp.b = plane; rp.z = 0; }
  rp.conv_vox = (((size_t)rp.b * g.in[0] + rp.z) * g.in[1] + rp.y) * g.in[2] + rp.x;
  if (EPI != EPI_D2S) plan_plain(g, p.ep, rp);
  epilogue_row<EPI>(g, p.ep, sm.sbias, t_addr + ((uint32_t)(q * 32) << 16), rp);
}


// depth_to_space head writing an unpadded 16-bit tensor (3-D, spatial factor r, 8 mapped channels
// per voxel, no residual / affine): one thread = one LR voxel, its cout = r*r*8 channels go to
// r*r HR voxels as 16-byte runs; 8 consecutive x voxels of a warp form full 128-byte lines.
// Two 16-column TMEM loads in flight per step, no divisions (run counters advance (i, j)).
__device__ __forceinline__ void epilogue_tile_d2s16(const UmmaParams& p, const SmemMap& sm,
                                                    const ItemCoord& c, int fr0, uint32_t t_addr,
                                                    int warp, int lane) {
  const ConvGeom& g = p.g;
  const Epilogue& ep = p.ep;
  const int q = warp & 3;
  const int mrow = q * 32 + lane;
  const int grp = mrow >> 3, xl = mrow & 7;
  const int fr = fr0 + grp;
  const int zq = fr / p.YB, yq = fr - zq * p.YB;
  const int plane = c.pl0 + zq;
  const int y = c.y0 + yq, x = c.xb * 8 + xl;
  const bool valid = yq <= p.YB - 3 && y < g.in[1] && x < g.in[2] && plane < p.planes;
  const uint32_t ta = t_addr + ((uint32_t)(q * 32) << 16);
  const int r = g.r;
  // destination element offset of run (i = 0, j = 0): HR voxel (z r, y r, x), 8 channels
  const long long sy = (long long)g.fd[2] * g.cstride, sz = (long long)g.fd[1] * sy;
  uint16_t* base = reinterpret_cast<uint16_t*>(ep.y_hi) +
                   ((((long long)c.b * g.fd[0] + (long long)plane * r) * g.fd[1] + (long long)y * r) *
                        g.fd[2] + x) * g.cstride + g.coff;
  int ri = 0, rj = 0;   // run counters: channel run number = ri * r + rj
  const int cout = g.cout, act = g.act;
  const float alpha = g.alpha, sc = ep.acc_scale;
#pragma unroll 1
  for (int c0 = 0; c0 < cout; c0 += 32) {
    uint32_t raw[32];
    tmem_ld16(ta + c0, *reinterpret_cast<uint32_t(*)[16]>(&raw[0]));
    if (c0 + 16 < cout) tmem_ld16(ta + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[16]));
    tmem_ld_wait();
    if (!valid) continue;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (c0 + 8 * s < cout) {
        float v[8];
        const float4 b0 = *reinterpret_cast<const float4*>(sm.sbias + c0 + 8 * s);
        const float4 b1 = *reinterpret_cast<const float4*>(sm.sbias + c0 + 8 * s + 4);
        v[0] = fmaf(__uint_as_float(raw[8 * s]), sc, b0.x);     v[1] = fmaf(__uint_as_float(raw[8 * s + 1]), sc, b0.y);
        v[2] = fmaf(__uint_as_float(raw[8 * s + 2]), sc, b0.z); v[3] = fmaf(__uint_as_float(raw[8 * s + 3]), sc, b0.w);
        v[4] = fmaf(__uint_as_float(raw[8 * s + 4]), sc, b1.x); v[5] = fmaf(__uint_as_float(raw[8 * s + 5]), sc, b1.y);
        v[6] = fmaf(__uint_as_float(raw[8 * s + 6]), sc, b1.z); v[7] = fmaf(__uint_as_float(raw[8 * s + 7]), sc, b1.w);
        if (act == S3_ACT_LEAKY) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = v[k] >= 0.f ? v[k] : alpha * v[k];
        } else if (act == S3_ACT_RELU) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
        } else if (act != S3_ACT_NONE) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = apply_act_slow(v[k], act, alpha);
        }
        uint4 u;
        u.x = pack2(v[0], v[1], ep.fmt);
        u.y = pack2(v[2], v[3], ep.fmt);
        u.z = pack2(v[4], v[5], ep.fmt);
        u.w = pack2(v[6], v[7], ep.fmt);
        *reinterpret_cast<uint4*>(base + ri * sz + rj * sy) = u;
        if (++rj == r) { rj = 0; ++ri; }
      }
    }
  }
}

// launchers implemented in conv_umma_zring.cu / conv_umma_tile.cu
int launch_umma_zring(const UmmaParams& p, const CUtensorMap* maps /* a, w, a2, w2 */,
                      const CUtensorMap* epi_maps /* res_hi, res_lo, y_hi, y_lo, row_hi, row_lo or NULL */,
                      int epi, int ctas, uint32_t smem, cudaStream_t st);
int launch_umma_tile(const UmmaParams& p, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                     const CUtensorMap& w_hi, const CUtensorMap& w_lo, int epi, int ctas,
                     uint32_t smem, cudaStream_t st);

}  // namespace s3
