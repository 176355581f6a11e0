// tcgen05 convolution, "zcat" scheme: 3-D convolutions with a narrow output (3*npad <= 256).
// An SS-mode tcgen05.mma with M = 128 costs max(~49..64, N/2) cycles (measured on B200,
// tools/microbench/mma_rate*.cu: the A operand streams from shared memory), so N = 64 would
// cap the tensor pipe near 50 %.  Here the weight slab of one (dy, dx) holds the three dz taps
// stacked along N ([3*npad][64]); one MMA on INPUT plane ip then feeds three OUTPUT planes at
// once: block j of the result belongs to output plane ip - j.  The accumulators of an item's R
// output planes sit at descending TMEM columns so that the three blocks land in consecutive
// columns: col(O_r) = base + npad * (R - 1 - r).
// Loop order: (dy, dx) slab outer (each slab is fetched once per item), input planes inner
// (fully unrolled for the compile-time R: all descriptors offsets are immediates).
#include "conv_umma_common.cuh"

namespace s3 {

// kR > 0: compile-time planes per item; kR == 0: runtime p.R (rare shapes)
template <int kR, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_zcat_kernel(const __grid_constant__ CUtensorMap tm_a_hi,
                      const __grid_constant__ CUtensorMap tm_a_lo,
                      const __grid_constant__ CUtensorMap tm_w_hi,
                      const __grid_constant__ CUtensorMap tm_w_lo, const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemMap sm = carve(p, smem_raw);
  auto bar = [&](int i) { return sm.bar_base + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_base = setup_cta(p, sm, &tm_a_hi, &tm_a_lo, &tm_w_hi, &tm_w_lo);
  const int R = kR > 0 ? kR : p.R;
  const int NP = R + 2;                                  // input planes per item
  const uint32_t plane_bytes = (uint32_t)p.YB * p.XB * 128u;
  const int n_items = p.n_items, WS = p.WS, AS = p.AS, acc_bufs = p.acc_bufs;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (warp-uniform)
    int as = 0, aph = 0, ws = 0, wph = 0, it = 0;
    auto load_slab = [&](int s) {
      mbar_wait(bar(B_WEMPTY + ws), wph ^ 1, p.dbg, 2, ws, it * 100 + s);
      if (elect_one()) {
        mbar_expect_tx(bar(B_WFULL + ws), p.w_bytes);
        tma_load_3d(sm.w_base + ws * sm.w_stage_bytes, &tm_w_hi, bar(B_WFULL + ws), 0, 0, s);
      }
      __syncwarp();
      if (++ws == WS) { ws = 0; wph ^= 1; }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const ItemCoord c = decode_item(p, item);
      load_slab(0);
      const int plane0 = c.b * p.plane_pitch + c.pl0;
      for (int ip = 0; ip < NP; ++ip) {
        const int bi = as * kMaxPlanes + ip;
        mbar_wait(bar(B_AEMPTY + bi), aph ^ 1, p.dbg, 1, bi, it);
        if (elect_one()) {
          mbar_expect_tx(bar(B_AFULL + bi), plane_bytes);
          tma_load_4d(sm.a_base + as * sm.a_stage_bytes + ip * plane_bytes, &tm_a_hi,
                      bar(B_AFULL + bi), 0, c.xb * 8, c.y0, plane0 + ip);
        }
        __syncwarp();
      }
      if (++as == AS) { as = 0; aph ^= 1; }
      for (int s = 1; s < 9; ++s) load_slab(s);
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (warp-uniform, elected issue)
    int as = 0, aph = 0, ws = 0, wph = 0, ab = 0, abph = 0, it = 0;
    const int npad = p.npad;
    const uint32_t fmtb = p.fmt == 0 ? 1u : 0u;
    const uint32_t hi_a = sdesc_hi_sw128((uint32_t)p.XB * 128u);
    const uint32_t hi_b = sdesc_hi_sw128(1024u);
    const uint32_t id1 = make_idesc_f16((uint32_t)npad, fmtb);
    const uint32_t id2 = make_idesc_f16((uint32_t)(2 * npad), fmtb);
    const uint32_t id3 = make_idesc_f16((uint32_t)(3 * npad), fmtb);
    const uint32_t blk_lo = ((uint32_t)npad * 128u) >> 4;   // one weight block in desc units
    const uint32_t plane_lo = plane_bytes >> 4;
    const uint32_t xb128 = (uint32_t)p.XB * 128u;
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0;
    long long t_acc = 0, t_w = 0, t_a = 0, t_issue = 0, t_all0 = tr ? clock64() : 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      long long c0 = tr ? clock64() : 0;
      mbar_wait(bar(B_ACCEMPTY + ab), abph ^ 1, p.dbg, 3, ab, it);
      if (tr) t_acc += clock64() - c0;
      tc_fence_after();
      const uint32_t a_stage = sm.a_base + as * sm.a_stage_bytes;
      const uint32_t acc0 = tmem_base + (uint32_t)(ab * R * npad);
#pragma unroll 1
      for (int s = 0; s < 9; ++s) {
        const int dy = s / 3, dx = s - 3 * dy;
        long long c1 = tr ? clock64() : 0;
        mbar_wait(bar(B_WFULL + ws), wph, p.dbg, 5, ws, it * 100 + s);
        if (tr) t_w += clock64() - c1;
        if (s == 0) {
          long long c3 = tr ? clock64() : 0;
          for (int ip = 0; ip < NP; ++ip)
            mbar_wait(bar(B_AFULL + as * kMaxPlanes + ip), aph, p.dbg, 4, ip, it);
          if (tr) t_a += clock64() - c3;
        }
        tc_fence_after();
        const uint32_t wl = sdesc_lo(sm.w_base + ws * sm.w_stage_bytes);
        const uint32_t al0 = sdesc_lo(a_stage + (uint32_t)dy * xb128 + (uint32_t)dx * 128u);
        long long c2 = tr ? clock64() : 0;
        if (elect_one()) {
          if (kR > 0) {
#pragma unroll
            for (int ip = 0; ip < kR + 2; ++ip) {
              constexpr int Rc = kR > 0 ? kR : 1;
              const int jlo = ip - (Rc - 1) > 0 ? ip - (Rc - 1) : 0;
              const int jhi = ip < 2 ? ip : 2;
              const int nblk = jhi - jlo + 1;
              const uint32_t dcol = acc0 + (uint32_t)(npad * (Rc - 1 - (ip - jlo)));
              const uint32_t al = al0 + (uint32_t)ip * plane_lo;
              const uint32_t bl = wl + (uint32_t)jlo * blk_lo;
              const uint32_t idn = nblk == 3 ? id3 : (nblk == 2 ? id2 : id1);
              if (s == 0 && jlo == 0) {
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
          } else {
            for (int ip = 0; ip < NP; ++ip) {
              const int jlo = ip - (R - 1) > 0 ? ip - (R - 1) : 0;
              const int jhi = ip < 2 ? ip : 2;
              const int nblk = jhi - jlo + 1;
              const uint32_t dcol = acc0 + (uint32_t)(npad * (R - 1 - (ip - jlo)));
              const uint32_t al = al0 + (uint32_t)ip * plane_lo;
              const uint32_t bl = wl + (uint32_t)jlo * blk_lo;
              const uint32_t idn = nblk == 3 ? id3 : (nblk == 2 ? id2 : id1);
              if (s == 0 && jlo == 0) {
                umma_f16_new(dcol, mk_desc(al, hi_a), mk_desc(wl, hi_b), id1);
                if (nblk > 1)
                  umma_f16_acc(dcol + npad, mk_desc(al, hi_a), mk_desc(wl + blk_lo, hi_b),
                               nblk == 3 ? id2 : id1);
                for (int kk = 1; kk < 4; ++kk)
                  umma_f16_acc(dcol, mk_desc(al + 2u * kk, hi_a), mk_desc(bl + 2u * kk, hi_b), idn);
              } else {
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16_acc(dcol, mk_desc(al + 2u * kk, hi_a), mk_desc(bl + 2u * kk, hi_b), idn);
              }
              if (s == 8) umma_commit(bar(B_AEMPTY + as * kMaxPlanes + ip));
            }
          }
          umma_commit(bar(B_WEMPTY + ws));
          if (s == 8) umma_commit(bar(B_ACCFULL + ab));
        }
        __syncwarp();
        if (tr) t_issue += clock64() - c2;
        if (++ws == WS) { ws = 0; wph ^= 1; }
      }
      if (++as == AS) { as = 0; aph ^= 1; }
      if (++ab == acc_bufs) { ab = 0; abph ^= 1; }
    }
    if (tr) {
      p.trace[0] = clock64() - t_all0; p.trace[1] = t_acc; p.trace[2] = t_w; p.trace[3] = t_a;
      p.trace[4] = t_issue; p.trace[5] = it;
    }
  } else {
    // ---------------------------------------------------------------------------- epilogue
    int ab = 0, abph = 0, it = 0;
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
    long long t_wait = 0, t_work = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const ItemCoord c = decode_item(p, item);
      long long c0 = tr ? clock64() : 0;
      mbar_wait(bar(B_ACCFULL + ab), abph, p.dbg, 6, ab, it);
      long long c1 = tr ? clock64() : 0;
      if (tr) t_wait += c1 - c0;
      tc_fence_after();
      for (int r = 0; r < R; ++r)
        epilogue_tile<EPI>(p, sm, c, r * p.YB,
                           tmem_base + (uint32_t)(ab * R * p.npad + p.npad * (R - 1 - r)), warp,
                           lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_ACCEMPTY + ab));
      if (++ab == acc_bufs) { ab = 0; abph ^= 1; }
      if (tr) t_work += clock64() - c1;
    }
    if (tr) { p.trace[8] = t_wait; p.trace[9] = t_work; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int kR, int EPI>
static int launch_zcat_t(const UmmaParams& p, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                         const CUtensorMap& w_hi, const CUtensorMap& w_lo, int ctas, uint32_t smem,
                         cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    S3_CUDA(cudaFuncSetAttribute(conv_umma_zcat_kernel<kR, EPI>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr = true;
  }
  conv_umma_zcat_kernel<kR, EPI><<<ctas, kThreads, smem, st>>>(a_hi, a_lo, w_hi, w_lo, p);
  S3_CUDA(cudaGetLastError());
  return S3_OK;
}

int launch_umma_zcat(const UmmaParams& p, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                     const CUtensorMap& w_hi, const CUtensorMap& w_lo, int epi, int ctas,
                     uint32_t smem, cudaStream_t st) {
  if (epi == EPI_PLAIN && p.R == 4)
    return launch_zcat_t<4, EPI_PLAIN>(p, a_hi, a_lo, w_hi, w_lo, ctas, smem, st);
  if (epi == EPI_PLAIN)
    return launch_zcat_t<0, EPI_PLAIN>(p, a_hi, a_lo, w_hi, w_lo, ctas, smem, st);
  return launch_zcat_t<0, EPI_GENERIC>(p, a_hi, a_lo, w_hi, w_lo, ctas, smem, st);
}

}  // namespace s3
