// tcgen05 convolution, "zring" scheme: the z-concatenated-N kernel (see conv_umma_zcat.cu for
// the MMA arrangement: one (dy, dx) weight slab holds the three dz taps stacked along N, so one
// MMA on INPUT plane ip feeds OUTPUT planes ip, ip-1, ip-2) with a software pipeline that never
// drains between work items:
//   * activations live in a RING of P plane slots (one TMA box + full/empty mbarrier pair per
//     plane).  A CTA owns a contiguous range of work items ordered z-fastest inside a
//     (batch, y block, x block) column, so consecutive items share two input planes (no z-halo
//     re-read) and the planes of item i+1 stream in while the last slab of item i is still
//     being multiplied; the MMA warp waits per plane, not per item;
//   * weight slabs stream through a 2-deep ring (one slab is consumed for ~1.8 k cycles, the
//     next one lands meanwhile);
//   * accumulators are double-buffered in TMEM and drained by EIGHT epilogue warps (two per
//     TMEM lane quarter, alternating output planes), which run with 216 registers each
//     (setmaxnreg) while the producer / issuer warpgroup keeps 64.
// Reference semantics: FlexiblePadding(3, REFLECT) -> Conv3D(valid) -> Cropping3D(2)
// [-> LeakyReLU] [-> nearest repeat] [-> SkipConnection add] as executed by
// sup3r/models/abstract.py:1081-1092 over sup3r/configs/spatiotemporal/gen_*.json.
#include "conv_umma_common.cuh"

namespace s3 {

constexpr int kRingThreads = 384;    // WG0: TMA + MMA (+2 idle warps); WG1, WG2: epilogue
constexpr int kRingMaxP = 8;
constexpr int kRingMaxWS = 4;
constexpr int RB_PFULL = 0;
constexpr int RB_PEMPTY = RB_PFULL + kRingMaxP;
constexpr int RB_WFULL = RB_PEMPTY + kRingMaxP;
constexpr int RB_WEMPTY = RB_WFULL + kRingMaxWS;
constexpr int RB_ACCFULL = RB_WEMPTY + kRingMaxWS;
constexpr int RB_ACCEMPTY = RB_ACCFULL + 2;
constexpr int RB_TMEMPTR = RB_ACCEMPTY + 2;

template <int N>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

struct RingItem {
  int b, yb, xb, grp;   // column + z group
  int pl0;              // first output plane
  int ri;               // output planes in this item (<= R)
  bool cont;            // continues the previous item's column (two planes already in the ring)
  bool next_cont;       // the next item of this CTA continues this column
};

__device__ __forceinline__ RingItem ring_decode(const UmmaParams& p, int i, int i0, int i1) {
  RingItem c;
  const int G = p.groups_per_b;
  const int col = i / G;
  c.grp = i - col * G;
  c.xb = col % p.nxb;
  const int rest = col / p.nxb;
  c.yb = rest % p.nyb;
  c.b = rest / p.nyb;
  c.pl0 = c.grp * p.R;
  c.ri = min(p.R, p.planes - c.pl0);
  c.cont = i > i0 && c.grp > 0;
  c.next_cont = (i + 1 < i1) && (c.grp + 1 < G);
  return c;
}

// ---------------------------------------------------------------------------------- epilogue
// One thread = one output voxel (64 fp32 accumulator columns).  The residual row (if any) is
// requested before the TMEM loads, the whole 64-channel row is finished in registers and then
// written with back-to-back 16-byte stores (full 128-B lines per voxel).
__device__ __forceinline__ void ring_epilogue_row64(const ConvGeom& g, const Epilogue& ep,
                                                    const float* sbias, uint32_t t_addr,
                                                    const RowPlan& rp) {
  float4 rpre[16];
  const bool has_res = ep.residual != nullptr;
  if (rp.valid && has_res) {
    const float4* rr = reinterpret_cast<const float4*>(ep.residual + rp.conv_vox * 64);
#pragma unroll
    for (int q = 0; q < 16; ++q) rpre[q] = __ldg(rr + q);
  }
  uint32_t raw[64];
#pragma unroll
  for (int cc = 0; cc < 4; ++cc)
    tmem_ld16(t_addr + cc * 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[cc * 16]));
  tmem_ld_wait();
  if (!rp.valid) return;
  float v[64];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float4 bv = *reinterpret_cast<const float4*>(sbias + 4 * q);
    v[4 * q] = __uint_as_float(raw[4 * q]) + bv.x;
    v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + bv.y;
    v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + bv.z;
    v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + bv.w;
  }
  if (g.act == S3_ACT_LEAKY) {
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = v[j] >= 0.f ? v[j] : g.alpha * v[j];
  } else if (g.act == S3_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (g.act != S3_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = apply_act(v[j], g.act, g.alpha);
  }
  if (has_res) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      v[4 * q] += rpre[q].x; v[4 * q + 1] += rpre[q].y;
      v[4 * q + 2] += rpre[q].z; v[4 * q + 3] += rpre[q].w;
    }
  }
  const int rep = g.rep[2];
  const int fmt = ep.fmt;
  if (ep.y) {
#pragma unroll 1
    for (int rx = 0; rx < rep; ++rx) {
      float4* dst = reinterpret_cast<float4*>(ep.y + rp.base32 + (size_t)rx * 64);
#pragma unroll
      for (int q = 0; q < 16; ++q)
        dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  }
  if (ep.y_hi) {
    uint16_t* yh = reinterpret_cast<uint16_t*>(ep.y_hi);
    uint16_t* yl = reinterpret_cast<uint16_t*>(ep.y_lo);
    uint4 h[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      h[q].x = pack2(v[8 * q], v[8 * q + 1], fmt);
      h[q].y = pack2(v[8 * q + 2], v[8 * q + 3], fmt);
      h[q].z = pack2(v[8 * q + 4], v[8 * q + 5], fmt);
      h[q].w = pack2(v[8 * q + 6], v[8 * q + 7], fmt);
    }
    if (yl) {
      // low half of the split: reuse v[] for the rounding residue
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = v[j] - from16(to16(v[j], fmt), fmt);
    }
#pragma unroll 1
    for (int rx = 0; rx < rep; ++rx) {
      const int ox = rp.x * rep + rx;
      const long long mx = ox == 1 ? -128LL : (ox == g.fd[2] - 2 ? 128LL : 0LL);
      const long long o0 = (long long)rp.base16 + (long long)rx * 64;
#pragma unroll 1
      for (int combo = 0; combo < 8; ++combo) {
        const bool a = combo & 4, bq = combo & 2, cq = combo & 1;
        if ((a && rp.mz == 0) || (bq && rp.my == 0) || (cq && mx == 0)) continue;
        const long long o = o0 + (a ? rp.mz : 0) + (bq ? rp.my : 0) + (cq ? mx : 0);
        uint4* d = reinterpret_cast<uint4*>(yh + o);
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] = h[q];
        if (yl) {
          uint4* dl = reinterpret_cast<uint4*>(yl + o);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint4 u;
            u.x = pack2(v[8 * q], v[8 * q + 1], fmt);
            u.y = pack2(v[8 * q + 2], v[8 * q + 3], fmt);
            u.z = pack2(v[8 * q + 4], v[8 * q + 5], fmt);
            u.w = pack2(v[8 * q + 6], v[8 * q + 7], fmt);
            dl[q] = u;
          }
        }
        if ((rp.mz | rp.my | mx) == 0) break;
      }
    }
  }
}

template <int EPI>
__device__ __forceinline__ void ring_epilogue_tile(const UmmaParams& p, const float* sbias,
                                                   const RingItem& c, int r, uint32_t t_addr,
                                                   int q, int lane) {
  const ConvGeom& g = p.g;
  const int mrow = q * 32 + lane;
  const int yq = mrow >> 3, xl = mrow & 7;
  RowPlan rp;
  rp.y = c.yb * 16 + yq;
  rp.x = c.xb * 8 + xl;
  rp.b = c.b;
  rp.z = c.pl0 + r;
  rp.valid = rp.y < g.in[1] && rp.x < g.in[2];
  rp.conv_vox = (((size_t)rp.b * g.in[0] + rp.z) * g.in[1] + rp.y) * g.in[2] + rp.x;
  plan_plain(g, p.ep, rp);
  const uint32_t ta = t_addr + ((uint32_t)(q * 32) << 16);
  if (EPI == EPI_PLAIN && g.cout == 64 && g.cstride == 64 && g.coff == 0)
    ring_epilogue_row64(g, p.ep, sbias, ta, rp);
  else
    epilogue_row<EPI>(g, p.ep, sbias, ta, rp);
}

// Issue the MMAs of one (dy, dx) slab for the hot configuration (R = 4 planes per item,
// npad = 64, 18 x 10 voxel planes, 7 ring slots) with every descriptor offset an immediate:
// S0 = ring slot of input plane 0.  One elected thread calls this; with ~4 uniform-datapath
// instructions per tcgen05.mma the issue stream stays ahead of the tensor pipe (a dynamic slot
// computation per plane cost ~60 issue cycles per MMA and starved it, profiles/r01_zring_v1).
template <int S0, bool kLast>
__device__ __forceinline__ void ring_issue_slab_fast(uint32_t a_tap, uint32_t wl, uint32_t hi_a,
                                                     uint32_t hi_b, uint32_t acc0, uint32_t id1,
                                                     uint32_t id2, uint32_t id3, bool keep_tail,
                                                     uint32_t pempty0) {
  constexpr int kR = 4, kP = 7;
  constexpr uint32_t kPlaneLo = (18u * 10u * 128u) >> 4, kBlkLo = (64u * 128u) >> 4;
#pragma unroll
  for (int ip = 0; ip < kR + 2; ++ip) {
    const int slot = (S0 + ip) % kP;
    const int jlo = ip - (kR - 1) > 0 ? ip - (kR - 1) : 0;
    const int jhi = ip < 2 ? ip : 2;
    const int nblk = jhi - jlo + 1;
    const uint32_t dcol = acc0 + (uint32_t)(64 * (kR - 1 - (ip - jlo)));
    const uint32_t al = a_tap + (uint32_t)slot * kPlaneLo;
    const uint32_t bl = wl + (uint32_t)jlo * kBlkLo;
    const uint32_t idn = nblk == 3 ? id3 : (nblk == 2 ? id2 : id1);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      umma_f16_acc(dcol, mk_desc(al + 2u * kk, hi_a), mk_desc(bl + 2u * kk, hi_b), idn);
    if (kLast && !(keep_tail && ip >= kR)) umma_commit(pempty0 + 8u * slot);
  }
}

template <bool kLast>
__device__ __forceinline__ void ring_issue_slab_fast_sw(int slot0, uint32_t a_tap, uint32_t wl,
                                                        uint32_t hi_a, uint32_t hi_b,
                                                        uint32_t acc0, uint32_t id1, uint32_t id2,
                                                        uint32_t id3, bool keep_tail,
                                                        uint32_t pempty0) {
  switch (slot0) {
    case 0: ring_issue_slab_fast<0, kLast>(a_tap, wl, hi_a, hi_b, acc0, id1, id2, id3, keep_tail, pempty0); break;
    case 1: ring_issue_slab_fast<1, kLast>(a_tap, wl, hi_a, hi_b, acc0, id1, id2, id3, keep_tail, pempty0); break;
    case 2: ring_issue_slab_fast<2, kLast>(a_tap, wl, hi_a, hi_b, acc0, id1, id2, id3, keep_tail, pempty0); break;
    case 3: ring_issue_slab_fast<3, kLast>(a_tap, wl, hi_a, hi_b, acc0, id1, id2, id3, keep_tail, pempty0); break;
    case 4: ring_issue_slab_fast<4, kLast>(a_tap, wl, hi_a, hi_b, acc0, id1, id2, id3, keep_tail, pempty0); break;
    case 5: ring_issue_slab_fast<5, kLast>(a_tap, wl, hi_a, hi_b, acc0, id1, id2, id3, keep_tail, pempty0); break;
    default: ring_issue_slab_fast<6, kLast>(a_tap, wl, hi_a, hi_b, acc0, id1, id2, id3, keep_tail, pempty0); break;
  }
}

// ------------------------------------------------------------------------------------ kernel
// kR > 0: compile-time planes per item (fully unrolled issue loop); kR == 0: runtime p.R
template <int kR, int EPI>
__global__ void __launch_bounds__(kRingThreads, 1)
conv_umma_zring_kernel(const __grid_constant__ CUtensorMap tm_a,
                       const __grid_constant__ CUtensorMap tm_w, const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t plane_bytes = (uint32_t)p.YB * p.XB * 128u;
  const int P = p.AS, WS = p.WS;
  const uint32_t a_base = base;
  const uint32_t w_slab = (p.w_bytes + 1023u) & ~1023u;
  const uint32_t w_base = (a_base + (uint32_t)P * plane_bytes + 1023u) & ~1023u;
  const uint32_t bar_base = w_base + (uint32_t)WS * w_slab;
  float* sbias = reinterpret_cast<float*>(smem_raw + (bar_base + 1024u - smem_u32(smem_raw)));
  auto bar = [&](int i) { return bar_base + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ------------------------------------------------------------------------------- set-up
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRingMaxP; ++i) {
      mbar_init(bar(RB_PFULL + i), 1);
      mbar_init(bar(RB_PEMPTY + i), 1);
    }
    for (int i = 0; i < kRingMaxWS; ++i) {
      mbar_init(bar(RB_WFULL + i), 1);
      mbar_init(bar(RB_WEMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(RB_ACCFULL + i), 1);
      mbar_init(bar(RB_ACCEMPTY + i), 8);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.npad; i += blockDim.x)
    sbias[i] = (p.ep.bias && i < p.g.cout) ? p.ep.bias[i] : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1) {
    tmem_alloc(bar(RB_TMEMPTR), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bar(RB_TMEMPTR)));

  const int R = kR > 0 ? kR : p.R;
  const int npad = p.npad;
  // contiguous item range of this CTA (z groups fastest inside a column)
  const int i0 = (int)(((long long)p.n_items * blockIdx.x) / gridDim.x);
  const int i1 = (int)(((long long)p.n_items * (blockIdx.x + 1)) / gridDim.x);

  if (warp < 4) {
    reg_dec<64>();
    if (warp == 0) {
      // ----------------------------------------------------------- TMA producer (warp-uniform)
      int head = 0, ws = 0, wph = 0;
      uint32_t pe_phase = 0;     // bit per plane slot: parity to wait for on its EMPTY barrier
      auto load_slab = [&](int s, int it) {
        mbar_wait_inl(bar(RB_WEMPTY + ws), wph ^ 1, p.dbg, 2, ws, it * 100 + s);
        if (elect_one()) {
          if ((p.dbg_flags & 4) && (it > 0 || s >= 2)) {
            mbar_arrive(bar(RB_WFULL + ws));
          } else {
            mbar_expect_tx(bar(RB_WFULL + ws), p.w_bytes);
            tma_load_3d(w_base + ws * w_slab, &tm_w, bar(RB_WFULL + ws), 0, 0, s);
          }
        }
        __syncwarp();
        if (++ws == WS) { ws = 0; wph ^= 1; }
      };
      for (int i = i0; i < i1; ++i) {
        const RingItem c = ring_decode(p, i, i0, i1);
        load_slab(0, i - i0);
        const int plane0 = c.b * p.plane_pitch + c.pl0;
        for (int ip = c.cont ? 2 : 0; ip < c.ri + 2; ++ip) {
          mbar_wait_inl(bar(RB_PEMPTY + head), ((pe_phase >> head) & 1u) ^ 1u, p.dbg, 1, head, i - i0);
          pe_phase ^= 1u << head;
          if (elect_one()) {
            if ((p.dbg_flags & 2) && i > i0) {
              mbar_arrive(bar(RB_PFULL + head));
            } else {
              mbar_expect_tx(bar(RB_PFULL + head), plane_bytes);
              tma_load_4d(a_base + head * plane_bytes, &tm_a, bar(RB_PFULL + head), 0, c.xb * 8,
                          c.yb * 16, plane0 + ip);
            }
          }
          __syncwarp();
          if (++head == P) head = 0;
        }
        for (int s = 1; s < 9; ++s) load_slab(s, i - i0);
      }
    } else if (warp == 1) {
      // ---------------------------------------------- MMA issuer (warp-uniform, elected issue)
      int ws = 0, wph = 0, ab = 0, abph = 0;
      int slot0 = 0;             // ring slot of input plane ip = 0 of the current item
      uint32_t pf_phase = 0;     // bit per plane slot: parity to wait for on its FULL barrier
      const uint32_t fmtb = p.fmt == 0 ? 1u : 0u;
      const uint32_t hi_a = sdesc_hi_sw128((uint32_t)p.XB * 128u);
      const uint32_t hi_b = sdesc_hi_sw128(1024u);
      const bool tiny = (p.dbg_flags & 32) != 0;   // experiment: N = 16 everywhere (issue cost)
      const uint32_t id1 = make_idesc_f16(tiny ? 16u : (uint32_t)npad, fmtb);
      const uint32_t id2 = make_idesc_f16(tiny ? 16u : (uint32_t)(2 * npad), fmtb);
      const uint32_t id3 = make_idesc_f16(tiny ? 16u : (uint32_t)(3 * npad), fmtb);
      const uint32_t blk_lo = ((uint32_t)npad * 128u) >> 4;   // one weight block in desc units
      const uint32_t xb128 = (uint32_t)p.XB * 128u;
      const bool fast = npad == 64 && p.XB == 10 && P == 7;
      const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0;
      long long t_acc = 0, t_w = 0, t_a = 0, t_all0 = tr ? clock64() : 0;
      for (int i = i0; i < i1; ++i) {
        const RingItem c = ring_decode(p, i, i0, i1);
        const int np = c.ri + 2;
        long long c0 = tr ? clock64() : 0;
        mbar_wait_inl(bar(RB_ACCEMPTY + ab), abph ^ 1, p.dbg, 3, ab, i - i0);
        if (tr) t_acc += clock64() - c0;
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(ab * R * npad);
#pragma unroll 1
        for (int s = 0; s < 9; ++s) {
          const int dy = s / 3, dx = s - 3 * dy;
          long long c1 = tr ? clock64() : 0;
          mbar_wait_inl(bar(RB_WFULL + ws), wph, p.dbg, 5, ws, (i - i0) * 100 + s);
          if (tr) t_w += clock64() - c1;
          tc_fence_after();
          const uint32_t wl = sdesc_lo(w_base + ws * w_slab);
          const uint32_t tap_off = (uint32_t)dy * xb128 + (uint32_t)dx * 128u;
          // one input plane: up to three output planes (blocks jlo..jhi of the slab)
          auto issue_plane = [&](int ip, int ri) {
            int slot = slot0 + ip;
            if (slot >= P) slot -= P;
            const int jlo = ip - (ri - 1) > 0 ? ip - (ri - 1) : 0;
            const int jhi = ip < 2 ? ip : 2;
            const int nblk = jhi - jlo + 1;
            const uint32_t dcol = acc0 + (uint32_t)(npad * (R - 1 - (ip - jlo)));
            const uint32_t al = sdesc_lo(a_base + (uint32_t)slot * plane_bytes + tap_off);
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
            if (s == 8 && !(c.next_cont && ip >= ri)) umma_commit(bar(RB_PEMPTY + slot));
          };
          if (s == 0) {
            // first slab: wait plane by plane so that the MMAs start as soon as the first
            // planes of the item have landed
            for (int ip = 0; ip < np; ++ip) {
              if (!(c.cont && ip < 2)) {
                int slot = slot0 + ip;
                if (slot >= P) slot -= P;
                long long c3 = tr ? clock64() : 0;
                mbar_wait_inl(bar(RB_PFULL + slot), (pf_phase >> slot) & 1u, p.dbg, 4, slot, i - i0);
                if (tr) t_a += clock64() - c3;
                pf_phase ^= 1u << slot;
                tc_fence_after();
              }
              if (elect_one()) issue_plane(ip, c.ri);
              __syncwarp();
            }
            if (elect_one()) umma_commit(bar(RB_WEMPTY + ws));
            __syncwarp();
          } else {
            if (elect_one()) {
              if (kR == 4 && fast && c.ri == 4) {
                const uint32_t a_tap = sdesc_lo(a_base + tap_off);
                if (s < 8)
                  ring_issue_slab_fast_sw<false>(slot0, a_tap, wl, hi_a, hi_b, acc0, id1, id2, id3,
                                                 false, bar(RB_PEMPTY));
                else
                  ring_issue_slab_fast_sw<true>(slot0, a_tap, wl, hi_a, hi_b, acc0, id1, id2, id3,
                                                c.next_cont, bar(RB_PEMPTY));
              } else {
                for (int ip = 0; ip < np; ++ip) issue_plane(ip, c.ri);
              }
              umma_commit(bar(RB_WEMPTY + ws));
              if (s == 8) umma_commit(bar(RB_ACCFULL + ab));
            }
            __syncwarp();
          }
          if (++ws == WS) { ws = 0; wph ^= 1; }
        }
        // the next item starts at the carried planes (continuing) or after all of this item's
        slot0 += c.next_cont ? c.ri : np;
        while (slot0 >= P) slot0 -= P;
        if (++ab == 2) { ab = 0; abph ^= 1; }
      }
      if (tr) {
        p.trace[0] = clock64() - t_all0; p.trace[1] = t_acc; p.trace[2] = t_w; p.trace[3] = t_a;
        p.trace[4] = 0; p.trace[5] = i1 - i0;
      }
    }
  } else {
    // ------------------------------------------------------------------------------ epilogue
    reg_inc<216>();
    const int wg = (warp - 4) >> 2, q = warp & 3;
    int ab = 0, abph = 0;
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 128;
    long long t_wait = 0, t_work = 0;
    for (int i = i0; i < i1; ++i) {
      const RingItem c = ring_decode(p, i, i0, i1);
      long long c0 = tr ? clock64() : 0;
      mbar_wait_inl(bar(RB_ACCFULL + ab), abph, p.dbg, 6, ab, i - i0);
      long long c1 = tr ? clock64() : 0;
      if (tr) t_wait += c1 - c0;
      tc_fence_after();
      if (!(p.dbg_flags & 8))
      for (int r = wg; r < c.ri; r += 2)
        ring_epilogue_tile<EPI>(p, sbias, c, r,
                                tmem_base + (uint32_t)(ab * R * npad + npad * (R - 1 - r)), q,
                                lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(RB_ACCEMPTY + ab));
      if (++ab == 2) { ab = 0; abph ^= 1; }
      if (tr) t_work += clock64() - c1;
    }
    if (tr) { p.trace[8] = t_wait; p.trace[9] = t_work; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int kR, int EPI>
static int launch_zring_t(const UmmaParams& p, const CUtensorMap& a, const CUtensorMap& w,
                          int ctas, uint32_t smem, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    S3_CUDA(cudaFuncSetAttribute(conv_umma_zring_kernel<kR, EPI>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr = true;
  }
  conv_umma_zring_kernel<kR, EPI><<<ctas, kRingThreads, smem, st>>>(a, w, p);
  S3_CUDA(cudaGetLastError());
  return S3_OK;
}

int launch_umma_zring(const UmmaParams& p, const CUtensorMap& a, const CUtensorMap& w, int epi,
                      int ctas, uint32_t smem, cudaStream_t st) {
  if (epi == EPI_PLAIN && p.R == 4) return launch_zring_t<4, EPI_PLAIN>(p, a, w, ctas, smem, st);
  if (epi == EPI_PLAIN) return launch_zring_t<0, EPI_PLAIN>(p, a, w, ctas, smem, st);
  return launch_zring_t<0, EPI_GENERIC>(p, a, w, ctas, smem, st);
}

}  // namespace s3
