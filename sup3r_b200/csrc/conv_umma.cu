// tcgen05 implicit-GEMM convolution for the 64-channel body of the sup3r generators:
// 3x3x3 (5-D tensors) or 3x3 (4-D tensors), stride 1, reflect-pad-1 semantics.  This is the
// fused form of the reference's  FlexiblePadding(3, REFLECT) -> Conv(valid) -> Cropping(2)
// [-> LeakyReLU] [-> SpatioTemporalExpansion] [-> SkipConnection add]  layer runs
// (sup3r/configs/spatiotemporal/gen_*.json, executed by sup3r/models/abstract.py:1081-1092).
//
// Data layout in HBM
//   activations : 16-bit (bf16 | fp16), channels-last, padded by one voxel on every convolved
//                 dim with the REFLECT halo already materialised by the producing kernel:
//                 [planes][Y+2][X+2][64], planes = N*(Z+2) (3-D) or N (2-D).  One voxel = 128 B
//                 = one SWIZZLE_128B row, so any (dz,dy,dx)-shifted window of a smem-resident
//                 box is a legal K-major UMMA operand: 8 consecutive x voxels form a core
//                 group, consecutive y rows are SBO = box_x*128 B apart.
//   weights     : [taps][Npad][64] 16-bit (Cout rows, Cin contiguous), Npad = Cout up to x16.
//   split mode  : activations and weights also carry a "lo" tensor (x - bf16(x)); the kernel
//                 accumulates hi*hi + lo*hi + hi*lo in fp32 (~16 mantissa bits).
//
// Kernel (persistent, one CTA per SM, 192 threads)
//   warp 0   : TMA producer  - one 4-D box load of the activation halo box per work item,
//              one 3-D load of the [Npad][64] weight slab per tap through a WS-deep ring
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=Npad, K=16);
//              R output tiles (128 voxels each) share every weight slab -> weight traffic / R
//   warps 2-5: epilogue - tcgen05.ld accumulators, bias / activation / residual / affine,
//              scatter (depth_to_space, nearest repeat, ...) to fp32 and/or the next layer's
//              16-bit padded+mirrored tensor.  Accumulators are double buffered in TMEM so
//              the epilogue of item i overlaps the MMAs of item i+1.
#include <cuda.h>

#include <cstring>

#include "common.cuh"
#include "ptx.cuh"
#include "umma_epilogue.cuh"

namespace s3 {

struct UmmaParams {
  ConvGeom g;
  Epilogue ep;
  int kz, ntaps, npad, split, fmt;
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
  const int halves = p.split ? 2 : 1;
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
      mbar_init(m.bar_base + 8u * (B_ACCEMPTY + i), 4);
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
  plan_plain(g, p.ep, rp);
  if (g.cout <= 64)
    epilogue_row<true>(g, p.ep, sm.sbias, t_addr + ((uint32_t)(q * 32) << 16), rp);
  else
    epilogue_row<false>(g, p.ep, sm.sbias, t_addr + ((uint32_t)(q * 32) << 16), rp);
}

// ============================================================================ kernel "tile"
// Every output tile accumulates all taps itself (N = npad).  Used for 2-D convolutions and
// for wide outputs (npad > 80) where one MMA already has N >= 128.
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_tile_kernel(const __grid_constant__ CUtensorMap tm_a_hi,
                      const __grid_constant__ CUtensorMap tm_a_lo,
                      const __grid_constant__ CUtensorMap tm_w_hi,
                      const __grid_constant__ CUtensorMap tm_w_lo, const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemMap sm = carve(p, smem_raw);
  auto bar = [&](int i) { return sm.bar_base + 8u * i; };
  const int halves = p.split ? 2 : 1;
  const uint32_t a_tx_bytes = p.box_bytes * halves;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_base = setup_cta(p, sm, &tm_a_hi, &tm_a_lo, &tm_w_hi, &tm_w_lo);

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (warp-uniform)
    int as = 0, aph = 0, ws = 0, wph = 0, it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const ItemCoord c = decode_item(p, item);
      mbar_wait(bar(B_AEMPTY + as), aph ^ 1, p.dbg, 1, as, it);
      if (elect_one()) {
        mbar_expect_tx(bar(B_AFULL + as), a_tx_bytes);
        const int plane = c.b * p.plane_pitch + c.pl0;
        tma_load_4d(sm.a_base + as * sm.a_stage_bytes, &tm_a_hi, bar(B_AFULL + as), 0, c.xb * 8,
                    c.y0, plane);
        if (p.split)
          tma_load_4d(sm.a_base + as * sm.a_stage_bytes + p.box_stride, &tm_a_lo,
                      bar(B_AFULL + as), 0, c.xb * 8, c.y0, plane);
      }
      __syncwarp();
      if (++as == p.AS) { as = 0; aph ^= 1; }
      for (int tap = 0; tap < p.ntaps; ++tap) {
        mbar_wait(bar(B_WEMPTY + ws), wph ^ 1, p.dbg, 2, ws, it * 100 + tap);
        if (elect_one()) {
          mbar_expect_tx(bar(B_WFULL + ws), p.w_bytes * halves);
          tma_load_3d(sm.w_base + ws * sm.w_stage_bytes, &tm_w_hi, bar(B_WFULL + ws), 0, 0, tap);
          if (p.split)
            tma_load_3d(sm.w_base + ws * sm.w_stage_bytes + sm.w_slab, &tm_w_lo,
                        bar(B_WFULL + ws), 0, 0, tap);
        }
        __syncwarp();
        if (++ws == p.WS) { ws = 0; wph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (warp-uniform, elected issue)
    int as = 0, aph = 0, ws = 0, wph = 0, ab = 0, abph = 0, it = 0;
    const uint32_t hi_a = sdesc_hi_sw128((uint32_t)p.XB * 128u);
    const uint32_t hi_b = sdesc_hi_sw128(1024u);
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const ItemCoord c = decode_item(p, item);
      mbar_wait(bar(B_ACCEMPTY + ab), abph ^ 1, p.dbg, 3, ab, it);
      mbar_wait(bar(B_AFULL + as), aph, p.dbg, 4, as, it);
      tc_fence_after();
      const uint32_t a_hi = sm.a_base + as * sm.a_stage_bytes;
      const uint32_t d_base = tmem_base + (uint32_t)(ab * p.R * p.npad);
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const int dx = tap % 3, dy = (tap / 3) % 3, dz = tap / 9;
        mbar_wait(bar(B_WFULL + ws), wph, p.dbg, 5, ws, it * 100 + tap);
        tc_fence_after();
        const uint32_t wl = sdesc_lo(sm.w_base + ws * sm.w_stage_bytes);
        const uint32_t wl_lo = sdesc_lo(sm.w_base + ws * sm.w_stage_bytes + sm.w_slab);
        if (elect_one()) {
          for (int r = 0; r < p.R; ++r) {
            const uint32_t row = (uint32_t)(c.row0 + r * p.TS + dz * p.YB + dy);
            const uint32_t al = sdesc_lo(a_hi + (row * p.XB + dx) * 128u);
            const uint32_t al_lo = sdesc_lo(a_hi + p.box_stride + (row * p.XB + dx) * 128u);
            const uint32_t d_addr = d_base + (uint32_t)(r * p.npad);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t da = mk_desc(al + 2u * kk, hi_a);
              const uint64_t db = mk_desc(wl + 2u * kk, hi_b);
              if (tap == 0 && kk == 0) umma_f16_new(d_addr, da, db, p.idesc);
              else umma_f16_acc(d_addr, da, db, p.idesc);
              if (p.split) {
                umma_f16_acc(d_addr, mk_desc(al_lo + 2u * kk, hi_a), db, p.idesc);
                umma_f16_acc(d_addr, da, mk_desc(wl_lo + 2u * kk, hi_b), p.idesc);
              }
            }
          }
          umma_commit(bar(B_WEMPTY + ws));
          if (tap == p.ntaps - 1) {
            umma_commit(bar(B_AEMPTY + as));
            umma_commit(bar(B_ACCFULL + ab));
          }
        }
        __syncwarp();
        if (++ws == p.WS) { ws = 0; wph ^= 1; }
      }
      if (++as == p.AS) { as = 0; aph ^= 1; }
      if (++ab == p.acc_bufs) { ab = 0; abph ^= 1; }
    }
  } else {
    int ab = 0, abph = 0, it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const ItemCoord c = decode_item(p, item);
      mbar_wait(bar(B_ACCFULL + ab), abph, p.dbg, 6, ab, it);
      tc_fence_after();
      for (int r = 0; r < p.R; ++r)
        epilogue_tile(p, sm, c, c.row0 + r * p.TS, tmem_base + (uint32_t)((ab * p.R + r) * p.npad),
                      warp, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_ACCEMPTY + ab));
      if (++ab == p.acc_bufs) { ab = 0; abph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ============================================================================ kernel "zcat"
// 3-D convolutions with a narrow output (3*npad <= 256).  An SS-mode tcgen05.mma with M = 128
// costs max(64, N/2) cycles (measured on B200: the A operand streams from shared memory at
// 64 B/cycle), so N = 64 would cap the tensor pipe at 50 %.  Here the weight slab of one
// (dy, dx) holds the three dz taps stacked along N ([3*npad][64]); one MMA on INPUT plane ip
// then feeds three OUTPUT planes at once: block j of the result belongs to output plane
// ip - j.  The accumulators of an item's R output planes sit at descending TMEM columns so the
// three blocks land in consecutive columns: col(O_r) = base + npad * (R - 1 - r).
// Loop order: (dy, dx) slab outer (each slab is fetched once per item), input planes inner.
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_zcat_kernel(const __grid_constant__ CUtensorMap tm_a_hi,
                      const __grid_constant__ CUtensorMap tm_a_lo,
                      const __grid_constant__ CUtensorMap tm_w_hi,
                      const __grid_constant__ CUtensorMap tm_w_lo, const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemMap sm = carve(p, smem_raw);
  auto bar = [&](int i) { return sm.bar_base + 8u * i; };
  const int halves = p.split ? 2 : 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_base = setup_cta(p, sm, &tm_a_hi, &tm_a_lo, &tm_w_hi, &tm_w_lo);
  const int NP = p.ZB;                                   // input planes per item (R + 2)
  const uint32_t plane_bytes = (uint32_t)p.YB * p.XB * 128u;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (warp-uniform)
    int as = 0, aph = 0, ws = 0, wph = 0, it = 0;
    auto load_slab = [&](int s) {
      mbar_wait(bar(B_WEMPTY + ws), wph ^ 1, p.dbg, 2, ws, it * 100 + s);
      if (elect_one()) {
        mbar_expect_tx(bar(B_WFULL + ws), p.w_bytes * halves);
        tma_load_3d(sm.w_base + ws * sm.w_stage_bytes, &tm_w_hi, bar(B_WFULL + ws), 0, 0, s);
        if (p.split)
          tma_load_3d(sm.w_base + ws * sm.w_stage_bytes + sm.w_slab, &tm_w_lo, bar(B_WFULL + ws),
                      0, 0, s);
      }
      __syncwarp();
      if (++ws == p.WS) { ws = 0; wph ^= 1; }
    };
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const ItemCoord c = decode_item(p, item);
      load_slab(0);
      const int plane0 = c.b * p.plane_pitch + c.pl0;
      for (int ip = 0; ip < NP; ++ip) {
        const int bi = as * kMaxPlanes + ip;
        mbar_wait(bar(B_AEMPTY + bi), aph ^ 1, p.dbg, 1, bi, it);
        if (elect_one()) {
          mbar_expect_tx(bar(B_AFULL + bi), plane_bytes * halves);
          const uint32_t dst = sm.a_base + as * sm.a_stage_bytes + ip * plane_bytes;
          tma_load_4d(dst, &tm_a_hi, bar(B_AFULL + bi), 0, c.xb * 8, c.y0, plane0 + ip);
          if (p.split)
            tma_load_4d(dst + p.box_stride, &tm_a_lo, bar(B_AFULL + bi), 0, c.xb * 8, c.y0,
                        plane0 + ip);
        }
        __syncwarp();
      }
      if (++as == p.AS) { as = 0; aph ^= 1; }
      for (int s = 1; s < 9; ++s) load_slab(s);
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (warp-uniform, elected issue)
    int as = 0, aph = 0, ws = 0, wph = 0, ab = 0, abph = 0, it = 0;
    const int R = p.R, npad = p.npad;
    const uint32_t fmtb = p.fmt == 0 ? 1u : 0u;
    const uint32_t hi_a = sdesc_hi_sw128((uint32_t)p.XB * 128u);
    const uint32_t hi_b = sdesc_hi_sw128(1024u);
    const uint32_t id1 = make_idesc_f16((uint32_t)npad, fmtb);
    const uint32_t id2 = make_idesc_f16((uint32_t)(2 * npad), fmtb);
    const uint32_t id3 = make_idesc_f16((uint32_t)(3 * npad), fmtb);
    const uint32_t blk_lo = ((uint32_t)npad * 128u) >> 4;   // one weight block in desc units
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      mbar_wait(bar(B_ACCEMPTY + ab), abph ^ 1, p.dbg, 3, ab, it);
      tc_fence_after();
      const uint32_t a_stage = sm.a_base + as * sm.a_stage_bytes;
      const uint32_t acc0 = tmem_base + (uint32_t)(ab * R * npad);
      for (int s = 0; s < 9; ++s) {
        const int dy = s / 3, dx = s - 3 * dy;
        mbar_wait(bar(B_WFULL + ws), wph, p.dbg, 5, ws, it * 100 + s);
        tc_fence_after();
        const uint32_t wl = sdesc_lo(sm.w_base + ws * sm.w_stage_bytes);
        const uint32_t al0 = sdesc_lo(a_stage + ((uint32_t)(dy * p.XB + dx)) * 128u);
        for (int ip = 0; ip < NP; ++ip) {
          if (s == 0) {
            mbar_wait(bar(B_AFULL + as * kMaxPlanes + ip), aph, p.dbg, 4, ip, it);
            tc_fence_after();
          }
          const int jlo = ip - (R - 1) > 0 ? ip - (R - 1) : 0;
          const int jhi = ip < 2 ? ip : 2;
          const int nblk = jhi - jlo + 1;
          const uint32_t dcol = acc0 + (uint32_t)(npad * (R - 1 - (ip - jlo)));
          const uint32_t al = al0 + ((ip * plane_bytes) >> 4);
          const uint32_t bl = wl + (uint32_t)jlo * blk_lo;
          const uint32_t idn = nblk == 3 ? id3 : (nblk == 2 ? id2 : id1);
          if (elect_one()) {
            if (s == 0 && jlo == 0) {
              // block 0 (output plane ip) starts a new accumulation; the others continue
              umma_f16_new(dcol, mk_desc(al, hi_a), mk_desc(wl, hi_b), id1);
              if (nblk > 1)
                umma_f16_acc(dcol + npad, mk_desc(al, hi_a), mk_desc(wl + blk_lo, hi_b),
                             nblk == 3 ? id2 : id1);
#pragma unroll
              for (int kk = 1; kk < 4; ++kk)
                umma_f16_acc(dcol, mk_desc(al + 2u * kk, hi_a), mk_desc(bl + 2u * kk, hi_b), idn);
            } else {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_acc(dcol, mk_desc(al + 2u * kk, hi_a), mk_desc(bl + 2u * kk, hi_b), idn);
            }
            if (s == 8) umma_commit(bar(B_AEMPTY + as * kMaxPlanes + ip));
          }
          __syncwarp();
        }
        if (elect_one()) {
          umma_commit(bar(B_WEMPTY + ws));
          if (s == 8) umma_commit(bar(B_ACCFULL + ab));
        }
        __syncwarp();
        if (++ws == p.WS) { ws = 0; wph ^= 1; }
      }
      if (++as == p.AS) { as = 0; aph ^= 1; }
      if (++ab == p.acc_bufs) { ab = 0; abph ^= 1; }
    }
  } else {
    int ab = 0, abph = 0, it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const ItemCoord c = decode_item(p, item);
      mbar_wait(bar(B_ACCFULL + ab), abph, p.dbg, 6, ab, it);
      tc_fence_after();
      for (int r = 0; r < p.R; ++r)
        epilogue_tile(p, sm, c, r * p.YB,
                      tmem_base + (uint32_t)(ab * p.R * p.npad + p.npad * (p.R - 1 - r)), warp,
                      lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_ACCEMPTY + ab));
      if (++ab == p.acc_bufs) { ab = 0; abph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static int encode_map(CUtensorMap* tm, const void* base, int fmt, int rank, const uint64_t* dims,
                      const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return S3_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  uint64_t stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
  }
  CUresult r = enc(tm, fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d (rank %d)", (int)r, rank);
    return S3_ERR_CUDA;
  }
  return S3_OK;
}

static DebugRec* debug_rec() {
  static DebugRec* host = nullptr;
  static DebugRec* dev = nullptr;
  if (!host) {
    if (cudaHostAlloc(&host, sizeof(DebugRec), cudaHostAllocMapped) != cudaSuccess) return nullptr;
    memset(host, 0, sizeof(DebugRec));
    if (cudaHostGetDevicePointer(&dev, host, 0) != cudaSuccess) return nullptr;
  }
  return dev;
}

constexpr uint32_t kSmemLimit = 232448;  // 227 KB

}  // namespace s3

using namespace s3;

extern "C" int s3_umma_weight_layout(int ndim, int cout, int split) {
  // 1: zcat layout [9 (dy,dx)][3 (dz)][npad][64];  0: tap-major [taps][npad][64]
  // (the split-precision path keeps the tap-major kernel: its doubled operands do not fit
  // the zcat kernel's shared-memory plan)
  return (ndim == 3 && !split && 3 * s3_umma_npad(cout) <= 256) ? 1 : 0;
}

extern "C" int s3_conv_fwd_umma(const s3_conv_desc* d, const void* x_hi, const void* x_lo,
                                const void* w_hi, const void* w_lo, const float* bias,
                                const float* residual, const float* post_scale,
                                const float* post_shift, float* y, void* y_hi, void* y_lo,
                                const s3_umma_tuning* tune, s3_stream stream) {
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_geom(d, &p.g);
  if (rc) return rc;
  const ConvGeom& g = p.g;
  S3_REQUIRE(x_hi && w_hi && (y || y_hi), "s3_conv_fwd_umma: null operand / no destination");
  S3_REQUIRE((x_lo == nullptr) == (w_lo == nullptr),
             "s3_conv_fwd_umma: x_lo and w_lo must both be given (split) or both be NULL");
  S3_REQUIRE(g.cin == 64, "s3_conv_fwd_umma: cin must be 64, got %d", g.cin);
  S3_REQUIRE(g.cout <= 256, "s3_conv_fwd_umma: cout must be <= 256, got %d", g.cout);
  S3_REQUIRE(g.pad_mode == S3_PAD_REFLECT, "s3_conv_fwd_umma: needs REFLECT padding");
  const int kz = g.ndim == 3 ? 3 : 1;
  for (int i = 0; i < 3; ++i) {
    const int k = (i == 0) ? kz : 3, pd = (i == 0 && kz == 1) ? 0 : 1;
    S3_REQUIRE(g.k[i] == k && g.st[i] == 1 && g.pl[i] == pd && g.ph[i] == pd,
               "s3_conv_fwd_umma: needs kernel 3, stride 1, pad 1 on every convolved dim");
  }
  S3_REQUIRE(g.in[1] >= 2 && g.in[2] >= 2 && (kz == 1 || g.in[0] >= 2),
             "s3_conv_fwd_umma: reflect-1 needs extents >= 2");
  if (y_hi) S3_REQUIRE(g.fd[1] >= 2 && g.fd[2] >= 2, "s3_conv_fwd_umma: padded output too small");
  if (y_hi) S3_REQUIRE(g.r == 1 && g.m == 1, "s3_conv_fwd_umma: 16-bit output needs a plain map");

  s3_umma_tuning t;
  memset(&t, 0, sizeof(t));
  if (tune) t = *tune;
  p.ep = Epilogue{bias, residual, post_scale, post_shift, y, y_hi, y_lo, t.fmt};
  p.kz = kz;
  p.ntaps = kz * 9;
  p.npad = s3_umma_npad(g.cout);
  p.split = x_lo ? 1 : 0;
  p.fmt = t.fmt;
  p.XB = t.box_x > 0 ? t.box_x : 10;
  S3_REQUIRE(p.XB >= 10 && p.XB <= 64, "s3_conv_fwd_umma: box_x must be in [10, 64]");
  const int halves = p.split ? 2 : 1;
  const int Y = g.in[1], X = g.in[2];
  p.planes = kz == 3 ? g.in[0] : g.n;
  p.nb = kz == 3 ? g.n : 1;
  p.plane_pitch = kz == 3 ? g.in[0] + 2 : 0;
  p.nxb = (X + 7) / 8;
  const bool zcat = s3_umma_weight_layout(g.ndim, g.cout, p.split) == 1;
  p.w_bytes = (uint32_t)p.npad * 128u * (zcat ? 3u : 1u);
  const uint32_t w_slab = (p.w_bytes + 1023u) & ~1023u;
  p.WS = t.w_stages > 0 ? t.w_stages : (zcat ? 3 : 4);
  S3_REQUIRE(p.WS >= 1 && p.WS <= kMaxWS, "s3_conv_fwd_umma: w_stages must be in [1, %d]", kMaxWS);
  const uint32_t fixed = 3072u;  // alignment slack + barrier block + staged bias

  bool found = false;
  if (zcat) {
    int r_max = t.tiles > 0 ? t.tiles : 4;
    if (2 * r_max * p.npad > 512) r_max = 512 / (2 * p.npad);
    if (r_max > p.planes) r_max = p.planes;
    if (r_max < 1) r_max = 1;
    const int YB = t.box_y > 0 ? t.box_y : 18;
    S3_REQUIRE(YB >= 18 && YB <= 64, "s3_conv_fwd_umma: box_y must be in [18, 64]");
    for (int ws = p.WS; ws >= 2 && !found; --ws)
      for (int R = r_max; R >= 1 && !found; --R) {
        const int ZB = R + 2;
        const uint32_t box = (uint32_t)ZB * YB * p.XB * 128u;
        const uint32_t boxs = (box + 1023u) & ~1023u;
        if (boxs * halves + (uint32_t)ws * w_slab * halves + fixed > kSmemLimit) continue;
        p.flat = 0; p.R = R; p.YB = YB; p.ZB = ZB; p.TS = YB; p.WS = ws;
        p.box_bytes = box; p.box_stride = boxs;
        p.AS = (2 * boxs * halves + (uint32_t)ws * w_slab * halves + fixed <= kSmemLimit) ? 2 : 1;
        found = true;
      }
  } else {
    const int max_r_tmem = 512 / p.npad;
    int r_max = t.tiles > 0 ? t.tiles : (p.npad >= 128 ? 2 : 4);
    if (r_max > max_r_tmem) r_max = max_r_tmem;
    if (r_max > 8) r_max = 8;
    S3_REQUIRE(r_max >= 1, "s3_conv_fwd_umma: npad %d too wide for TMEM", p.npad);
    const double eff_plane = (double)Y / (((Y + 15) / 16) * 16.0);
    const double eff_flat = (double)Y / (Y + 2.0);
    for (int ws = p.WS; ws >= 1 && !found; --ws)
      for (int use_flat = (eff_flat > eff_plane + 0.02 ? 1 : 0); use_flat >= 0 && !found;
           --use_flat)
        for (int R = r_max; R >= 1 && !found; --R) {
          int YB, ZB, TS;
          if (!use_flat) {
            YB = 18; TS = 18;
            ZB = (kz == 3) ? R + 2 : R;
          } else {
            YB = Y + 2; TS = 16;
            const int rows = (YB - 1) + 16 * R + (kz == 3 ? 2 * YB : 0) + 2;
            ZB = (rows + YB - 1) / YB;
          }
          if (YB > 256 || ZB > 256) continue;
          const uint32_t box = (uint32_t)ZB * YB * p.XB * 128u;
          const uint32_t boxs = (box + 1023u) & ~1023u;
          if (boxs * halves + (uint32_t)ws * w_slab * halves + fixed > kSmemLimit) continue;
          p.flat = use_flat; p.R = R; p.YB = YB; p.ZB = ZB; p.TS = TS; p.WS = ws;
          p.box_bytes = box; p.box_stride = boxs;
          p.AS = (2 * boxs * halves + (uint32_t)ws * w_slab * halves + fixed <= kSmemLimit) ? 2 : 1;
          found = true;
        }
  }
  S3_REQUIRE(found, "s3_conv_fwd_umma: no tile shape fits shared memory (Y=%d, npad=%d)", Y, p.npad);
  p.acc_bufs = (2 * p.R * p.npad <= 512) ? 2 : 1;
  if (!p.flat) {
    p.nyb = (Y + 15) / 16;
    p.groups_per_b = (p.planes + p.R - 1) / p.R;
    p.n_items = p.nxb * p.nyb * p.groups_per_b * p.nb;
  } else {
    p.nyb = 1;
    p.groups_per_b = (p.planes * p.YB + 16 * p.R - 1) / (16 * p.R);
    p.n_items = p.nxb * p.groups_per_b * p.nb;
  }
  p.idesc = make_idesc_f16((uint32_t)p.npad, (uint32_t)(t.fmt == 0 ? 1 : 0));
  p.dbg = debug_rec();

  CUtensorMap tm_a_hi, tm_a_lo, tm_w_hi, tm_w_lo;
  const uint64_t total_planes = kz == 3 ? (uint64_t)g.n * (g.in[0] + 2) : (uint64_t)g.n;
  const uint64_t adims[4] = {64, (uint64_t)X + 2, (uint64_t)Y + 2, total_planes};
  const uint32_t abox[4] = {64, (uint32_t)p.XB, (uint32_t)p.YB, (uint32_t)(zcat ? 1 : p.ZB)};
  const uint64_t wdims[3] = {64, (uint64_t)p.npad * (zcat ? 3 : 1), (uint64_t)(zcat ? 9 : p.ntaps)};
  const uint32_t wbox[3] = {64, (uint32_t)p.npad * (zcat ? 3 : 1), 1};
  if ((rc = encode_map(&tm_a_hi, x_hi, t.fmt, 4, adims, abox))) return rc;
  if ((rc = encode_map(&tm_w_hi, w_hi, t.fmt, 3, wdims, wbox))) return rc;
  tm_a_lo = tm_a_hi;
  tm_w_lo = tm_w_hi;
  if (p.split) {
    if ((rc = encode_map(&tm_a_lo, x_lo, t.fmt, 4, adims, abox))) return rc;
    if ((rc = encode_map(&tm_w_lo, w_lo, t.fmt, 3, wdims, wbox))) return rc;
  }
  const uint32_t smem = p.box_stride * halves * p.AS + (uint32_t)p.WS * w_slab * halves + fixed;
  static bool attr_set = false;
  if (!attr_set) {
    S3_CUDA(cudaFuncSetAttribute(conv_umma_tile_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
    S3_CUDA(cudaFuncSetAttribute(conv_umma_zcat_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
    attr_set = true;
  }
  int ctas = t.max_ctas > 0 ? t.max_ctas : sm_count();
  if (ctas > p.n_items) ctas = p.n_items;
  if (zcat)
    conv_umma_zcat_kernel<<<ctas, kThreads, smem, as_stream(stream)>>>(tm_a_hi, tm_a_lo, tm_w_hi,
                                                                       tm_w_lo, p);
  else
    conv_umma_tile_kernel<<<ctas, kThreads, smem, as_stream(stream)>>>(tm_a_hi, tm_a_lo, tm_w_hi,
                                                                       tm_w_lo, p);
  S3_LAUNCH_CHECK("conv_umma_kernel");
  return S3_OK;
}
