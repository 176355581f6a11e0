// tcgen05 convolution, "tile" scheme: every 128-voxel output tile accumulates all taps itself
// (N = npad).  Used for wide outputs (npad > 80, e.g. the 64 -> 200 head with the fused
// depth_to_space scatter), 2-D convolutions and the split-operand modes (bf16x3: three kind::f16
// MMAs per k-step; fp16c: one kind::f16 + one kind::f8f6f4 MMA per k-step, common.cuh).
#include "conv_umma_common.cuh"

namespace s3 {

// bounded spin without clock reads
__device__ __forceinline__ void tile_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  for (uint32_t n = 0; !mbar_try_wait(bar, parity); ++n)
    if (n > (1u << 26)) __trap();
}

// MMA role for the hot wide-N shape (3-D, plane mode, R = 2 tiles per item, 18 x 10 voxel planes,
// 4 weight stages): one elected thread issues a whole (item, pass) stage (27 taps x 2 tiles x 4
// k-steps) as straight-line code with immediate descriptor offsets, and waits for the next
// tap's weights before the last MMA group of the current tap -- the tcgen05 queue only holds ~2
// MMAs, so every longer stretch of issue-side code is a tensor-pipe bubble (see
// conv_umma_zring.cu).  Two-pass formats (fp16c, p.seq2): the kind::f16 pass over the fp16
// tensors and the kind::f8f6f4 pass over the e4m3 corr tensors run back to back into the same
// accumulators; both stream through the same activation / weight stages.
template <bool kF8, bool kFirst>
__device__ __forceinline__ void tile_issue_pass(uint32_t a_lo, uint32_t w_base_lo, uint32_t w_stage_lo,
                                                uint32_t bar_wfull, uint32_t bar_wempty, int g,
                                                uint32_t d0, uint32_t d1, uint32_t hi_a,
                                                uint32_t hi_b, uint32_t idesc, bool more_slabs) {
  constexpr int kR = 2, kWS = 4;
#pragma unroll
  for (int tap = 0; tap < 27; ++tap) {
    const int dx = tap % 3, dy = (tap / 3) % 3, dz = tap / 9;
    const int gs = g + tap;
    const uint32_t wl = w_base_lo + (uint32_t)(gs & (kWS - 1)) * w_stage_lo;
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const uint32_t al = a_lo + (uint32_t)(((r + dz) * 180 + dy * 10 + dx) * 8);
      const uint32_t dd = r == 0 ? d0 : d1;
      if (r == kR - 1 && (tap < 26 || more_slabs))
        tile_wait(bar_wfull + 8u * ((gs + 1) & (kWS - 1)), (uint32_t)(((gs + 1) >> 2) & 1));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = mk_desc(al + 2u * kk, hi_a), db = mk_desc(wl + 2u * kk, hi_b);
        if (kFirst && tap == 0 && kk == 0) umma_f16_new(dd, da, db, idesc);
        else if (kF8) umma_f8_acc(dd, da, db, idesc);
        else umma_f16_acc(dd, da, db, idesc);
      }
    }
    umma_commit(bar_wempty + 8u * (gs & (kWS - 1)));
  }
}

__device__ __forceinline__ void tile_mma_fast(const UmmaParams& p, const SmemMap& sm,
                                              uint32_t tmem_base) {
  constexpr int kR = 2;
  auto bar = [&](int i) { return sm.bar_base + 8u * i; };
  const uint32_t hi_a = sdesc_hi_sw128(1280u), hi_b = sdesc_hi_sw128(1024u);
  const uint32_t idesc = p.idesc;
  const uint32_t w_base_lo = sdesc_lo(sm.w_base), w_stage_lo = sm.w_stage_bytes >> 4;
  const int npass = p.seq2 ? 2 : 1;
  int as = 0, aph = 0, ab = 0, abph = 0;
  int g = 0;   // weight slabs consumed: stage g & 3, parity (g >> 2) & 1
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    const uint32_t d0 = tmem_base + (uint32_t)(ab * kR * p.npad);
    const uint32_t d1 = d0 + (uint32_t)p.npad;
    const bool has_next = item + (int)gridDim.x < p.n_items;
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
      const uint32_t a_lo = sdesc_lo(sm.a_base + as * sm.a_stage_bytes);
      const bool last = pass == npass - 1;
      if (elect_one()) {
        if (pass == 0) tile_wait(bar(B_ACCEMPTY + ab), (uint32_t)(abph ^ 1));
        tile_wait(bar(B_AFULL + as), (uint32_t)aph);
        tile_wait(bar(B_WFULL + (g & 3)), (uint32_t)((g >> 2) & 1));
        tc_fence_after();
        // (the wait for the first slab of the next stage is taken inside the pass only when a
        // next stage of THIS item exists: its activation box may not have landed yet)
        if (pass == 0)
          tile_issue_pass<false, true>(a_lo, w_base_lo, w_stage_lo, bar(B_WFULL), bar(B_WEMPTY), g,
                                       d0, d1, hi_a, hi_b, idesc, false);
        else
          tile_issue_pass<true, false>(a_lo, w_base_lo, w_stage_lo, bar(B_WFULL), bar(B_WEMPTY), g,
                                       d0, d1, hi_a, hi_b, idesc, false);
        umma_commit(bar(B_AEMPTY + as));
        if (last) umma_commit(bar(B_ACCFULL + ab));
      }
      __syncwarp();
      g += 27;
      if (++as == p.AS) { as = 0; aph ^= 1; }
    }
    (void)has_next;
    if (++ab == p.acc_bufs) { ab = 0; abph ^= 1; }
  }
}

// ============================================================================ kernel "tile"
// Every output tile accumulates all taps itself (N = npad).  Used for 2-D convolutions and
// for wide outputs (npad > 80) where one MMA already has N >= 128.
template <int EPI>
__global__ void __launch_bounds__(kTileThreads, 1)
conv_umma_tile_kernel(const __grid_constant__ CUtensorMap tm_a_hi,
                      const __grid_constant__ CUtensorMap tm_a_lo,
                      const __grid_constant__ CUtensorMap tm_w_hi,
                      const __grid_constant__ CUtensorMap tm_w_lo, const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemMap sm = carve(p, smem_raw);
  auto bar = [&](int i) { return sm.bar_base + 8u * i; };
  const int halves = (p.split && !p.seq2) ? 2 : 1;
  const uint32_t a_tx_bytes = p.box_bytes * halves;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_base = setup_cta(p, sm, &tm_a_hi, &tm_a_lo, &tm_w_hi, &tm_w_lo);

  if (warp == 0 && p.seq2) {
    // ---------------------------------------- TMA producer, two sequential passes per item
    int as = 0, aph = 0, ws = 0, wph = 0, it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const ItemCoord c = decode_item(p, item);
      const int plane = c.b * p.plane_pitch + c.pl0;
      for (int pass = 0; pass < 2; ++pass) {
        mbar_wait(bar(B_AEMPTY + as), aph ^ 1, p.dbg, 1, as, it);
        if (elect_one()) {
          mbar_expect_tx(bar(B_AFULL + as), p.box_bytes);
          tma_load_4d(sm.a_base + as * sm.a_stage_bytes, pass ? &tm_a_lo : &tm_a_hi,
                      bar(B_AFULL + as), 0, c.xb * 8, c.y0, plane);
        }
        __syncwarp();
        if (++as == p.AS) { as = 0; aph ^= 1; }
        for (int tap = 0; tap < p.ntaps; ++tap) {
          mbar_wait(bar(B_WEMPTY + ws), wph ^ 1, p.dbg, 2, ws, it * 100 + tap);
          if (elect_one()) {
            mbar_expect_tx(bar(B_WFULL + ws), p.w_bytes);
            tma_load_3d(sm.w_base + ws * sm.w_stage_bytes, pass ? &tm_w_lo : &tm_w_hi,
                        bar(B_WFULL + ws), 0, 0, tap);
          }
          __syncwarp();
          if (++ws == p.WS) { ws = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 0) {
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
  } else if (warp == 1 && p.tile_fast) {
    tile_mma_fast(p, sm, tmem_base);
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
                if (p.fmt == kFmtFp16c) {   // e4m3 corr rows: lo_x w + x lo_w in one K = 32 MMA
                  umma_f8_acc(d_addr, mk_desc(al_lo + 2u * kk, hi_a),
                              mk_desc(wl_lo + 2u * kk, hi_b), p.idesc);
                } else {
                  umma_f16_acc(d_addr, mk_desc(al_lo + 2u * kk, hi_a), db, p.idesc);
                  umma_f16_acc(d_addr, da, mk_desc(wl_lo + 2u * kk, hi_b), p.idesc);
                }
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
      // two epilogue warps per TMEM lane quarter: they alternate over the R tiles of the item
      for (int r = (warp - 2) >> 2; r < p.R && !(p.dbg_flags & 8); r += 2) {
        const uint32_t ta = tmem_base + (uint32_t)((ab * p.R + r) * p.npad);
        if (EPI == EPI_D2S16)
          epilogue_tile_d2s16(p, sm, c, c.row0 + r * p.TS, ta, warp, lane);
        else
          epilogue_tile<EPI == EPI_D2S16 ? EPI_D2S : EPI>(p, sm, c, c.row0 + r * p.TS, ta, warp, lane);
      }
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


template <int EPI>
static int launch_tile_t(const UmmaParams& p, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                         const CUtensorMap& w_hi, const CUtensorMap& w_lo, int ctas, uint32_t smem,
                         cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    S3_CUDA(cudaFuncSetAttribute(conv_umma_tile_kernel<EPI>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr = true;
  }
  conv_umma_tile_kernel<EPI><<<ctas, kTileThreads, smem, st>>>(a_hi, a_lo, w_hi, w_lo, p);
  S3_CUDA(cudaGetLastError());
  return S3_OK;
}

int launch_umma_tile(const UmmaParams& p, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                     const CUtensorMap& w_hi, const CUtensorMap& w_lo, int epi, int ctas,
                     uint32_t smem, cudaStream_t st) {
  if (epi == EPI_PLAIN) return launch_tile_t<EPI_PLAIN>(p, a_hi, a_lo, w_hi, w_lo, ctas, smem, st);
  if (epi == EPI_D2S && p.g.ndim == 3 && p.g.m == 1 && p.g.cmap == 8 && p.ep.y_hi && !p.ep.y)
    return launch_tile_t<EPI_D2S16>(p, a_hi, a_lo, w_hi, w_lo, ctas, smem, st);
  if (epi == EPI_D2S) return launch_tile_t<EPI_D2S>(p, a_hi, a_lo, w_hi, w_lo, ctas, smem, st);
  return launch_tile_t<EPI_GENERIC>(p, a_hi, a_lo, w_hi, w_lo, ctas, smem, st);
}

}  // namespace s3
