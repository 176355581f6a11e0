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
constexpr int kRingThreadsV4 = 640;  // ... WG1..WG4: sixteen epilogue warps (EPI_V4)
__host__ __device__ constexpr int ring_threads(int epi) { return epi == 6 ? kRingThreadsV4 : kRingThreads; }
constexpr int kRingMaxP = 8;
constexpr int kRingMaxWS = 4;
constexpr int RB_PFULL = 0;
constexpr int RB_PEMPTY = RB_PFULL + kRingMaxP;
constexpr int RB_WFULL = RB_PEMPTY + kRingMaxP;
constexpr int RB_WEMPTY = RB_WFULL + kRingMaxWS;
constexpr int RB_ACCFULL = RB_WEMPTY + kRingMaxWS;
constexpr int RB_ACCEMPTY = RB_ACCFULL + 2;
constexpr int RB_TMEMPTR = RB_ACCEMPTY + 2;
constexpr int RB_EPILD = RB_TMEMPTR + 2;     // 2 x 8 per-warp residual load barriers (TMA epilogue)

struct EpiMaps {
  CUtensorMap res_hi, res_lo, y_hi, y_lo;   // interior views, box 32 ch x 8 x 4 x 1 (TMA epilogue)
  CUtensorMap row_hi, row_lo;               // padded views, box 32 ch x 8 x 1 x 1 (y halo rows)
};

template <int N>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// bounded spin without clock reads (one try_wait + branch on the hot path)
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  for (uint32_t n = 0; !mbar_try_wait(bar, parity); ++n)
    if (n > (1u << 26)) __trap();
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
                                                    const RowPlan& rp, int dbg = 0) {
  float4 rpre[16];
  const bool has_res = ep.residual != nullptr;
  if (rp.valid && has_res) {
    const float4* rr = reinterpret_cast<const float4*>(ep.residual + rp.conv_vox * 64);
#pragma unroll
    for (int q = 0; q < 16; ++q) rpre[q] = __ldg(rr + q);
  }
  uint32_t raw[64];
  if (dbg & 64) {
#pragma unroll
    for (int j = 0; j < 64; ++j) raw[j] = t_addr + j;
  } else {
#pragma unroll
  for (int cc = 0; cc < 4; ++cc)
    tmem_ld16(t_addr + cc * 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[cc * 16]));
  tmem_ld_wait();
  }
  if (!rp.valid) return;
  if ((dbg & 32) && raw[5] != 0x7fffffffu) return;
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
    for (int j = 0; j < 64; ++j) v[j] = apply_act_slow(v[j], g.act, g.alpha);
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

// ------------------------------------------------------------------- coalescing epilogue
// Measured (role trace, B200): global loads / stores issued thread-per-row (32 different
// 128-B lines per warp instruction) slow the concurrently running MMA stream almost 1:1 with
// their L1 wavefront count -- the SS-mode tcgen05.mma already uses ~85 % of the shared-memory
// bandwidth and the LSU shares that data path.  This epilogue therefore moves every global
// access to a row-coalesced mapping (8 lanes x 16 B = one 128-B voxel row, 4 rows = 512
// contiguous bytes per instruction) through a 2 KiB per-warp staging buffer with a 16-byte
// XOR swizzle (conflict-free for both the thread-per-row and the coalesced side).
//   * output: 16-bit padded rows (hi [+ lo = rounding residue]) incl. the REFLECT halo mirrors;
//   * residual: 16-bit hi + lo pair of the skip tensor (same padded layout), prefetched into
//     registers before the accumulator is waited for.
// One call = one warp = 32 accumulator rows = 4 y rows x 8 x voxels of one output plane.
struct TileGeom {
  long long sy, sz;      // byte strides of the padded 16-bit tensor along y / z
  long long base;        // byte offset of voxel (y0, x0) of this plane (interior position)
  long long mz;          // z mirror delta in bytes (0 = none), warp-uniform
  int y0, x0;
};

__device__ __forceinline__ uint32_t stage_off(int r, int k) {
  return (uint32_t)(r * 128 + ((k ^ (r & 7)) << 4));
}

__device__ __forceinline__ void unpack_add8(float* v, const uint4& u, int fmt) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] += from16((uint16_t)(w[j] & 0xffffu), fmt);
    v[2 * j + 1] += from16((uint16_t)(w[j] >> 16), fmt);
  }
}

template <bool kRes>
__device__ __forceinline__ void ring_epilogue_warp_v2(const ConvGeom& g, const Epilogue& ep,
                                                      const float* sbias, uint32_t t_addr,
                                                      const TileGeom& tg, uint8_t* stage,
                                                      int lane) {
  const int fmt = ep.fmt;
  const int FY = g.fd[1], FX = g.fd[2];
  const int lr = lane >> 3, lk = lane & 7;      // coalesced side: row within a group of 4, chunk
  const int half_of_lane = lane >> 4, rrow = lane & 15;   // thread-per-row side
  const bool has_lo = ep.y_lo != nullptr;
  // coalesced side: (h, i) -> m = 16 h + 4 i + lr, y = y0 + (m >> 3), x = x0 + (m & 7);
  // byte offset of this lane's 16-byte chunk of row m = rowoff + (m >> 3) * sy + (m & 7) * 128
  const long long lane_off = tg.base + (long long)(lr & 1 ? 0 : 0) + lk * 16;
  auto row_off = [&](int h, int i) -> long long {
    const int m = 16 * h + 4 * i + lr;
    return lane_off + (long long)(m >> 3) * tg.sy + (long long)(m & 7) * 128;
  };
  auto row_ok = [&](int h, int i) -> bool {
    const int m = 16 * h + 4 * i + lr;
    return tg.y0 + (m >> 3) < FY && tg.x0 + (m & 7) < FX;
  };

  // ---- residual: rolling register prefetch, two (operand, half) rounds in flight
  uint4 rres[2][4];
  const bool res_has_lo = kRes && ep.res_lo != nullptr;
  auto res_load = [&](int op, int h, uint4 (&dst)[4]) {
    const uint8_t* src = reinterpret_cast<const uint8_t*>(op == 0 ? ep.res_hi : ep.res_lo);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      dst[i] = row_ok(h, i) ? __ldg(reinterpret_cast<const uint4*>(src + row_off(h, i)))
                            : make_uint4(0, 0, 0, 0);
  };
  if (kRes) {
    res_load(0, 0, rres[0]);
    res_load(0, 1, rres[1]);
  }

  // ---- accumulator row -> registers, bias, activation
  float v[64];
  {
    uint32_t raw[64];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
      tmem_ld16(t_addr + cc * 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[cc * 16]));
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float4 bv = *reinterpret_cast<const float4*>(sbias + 4 * q);
      v[4 * q] = __uint_as_float(raw[4 * q]) + bv.x;
      v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + bv.y;
      v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + bv.z;
      v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + bv.w;
    }
  }
  if (g.act == S3_ACT_LEAKY) {
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = v[j] >= 0.f ? v[j] : g.alpha * v[j];
  } else if (g.act == S3_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (g.act != S3_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = apply_act_slow(v[j], g.act, g.alpha);
  }

  // ---- residual: coalesced registers -> staging -> own row (rounds: hi h0, hi h1, lo h0, lo h1)
  if (kRes) {
#pragma unroll
    for (int rnd = 0; rnd < 4; ++rnd) {
      const int op = rnd >> 1, h = rnd & 1;
      if (op == 0 || res_has_lo) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(stage + stage_off(4 * i + lr, lk)) = rres[rnd & 1][i];
        __syncwarp();
        if (rnd < 2 && res_has_lo) res_load(1, h, rres[rnd & 1]);   // next use: round rnd + 2
        if (half_of_lane == h) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint4 u = *reinterpret_cast<const uint4*>(stage + stage_off(rrow, k));
            unpack_add8(&v[8 * k], u, fmt);
          }
        }
        __syncwarp();
      }
    }
  }

  // ---- outputs: own row -> staging -> coalesced stores (+ halo mirrors)
#pragma unroll
  for (int op = 0; op < 2; ++op) {
    if (op == 0 || has_lo) {
    uint8_t* dst = reinterpret_cast<uint8_t*>(op == 0 ? ep.y_hi : ep.y_lo);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (half_of_lane == h) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float a[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = v[8 * k + j];
            a[j] = op == 0 ? x : x - from16(to16(x, fmt), fmt);
          }
          uint4 hk;
          hk.x = pack2(a[0], a[1], fmt);
          hk.y = pack2(a[2], a[3], fmt);
          hk.z = pack2(a[4], a[5], fmt);
          hk.w = pack2(a[6], a[7], fmt);
          *reinterpret_cast<uint4*>(stage + stage_off(rrow, k)) = hk;
        }
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (row_ok(h, i)) {
          const int m = 16 * h + 4 * i + lr;
          const int y = tg.y0 + (m >> 3), x = tg.x0 + (m & 7);
          const uint4 u = *reinterpret_cast<const uint4*>(stage + stage_off(4 * i + lr, lk));
          const long long off = row_off(h, i);
          const long long my = y == 1 ? -2 * tg.sy : (y == FY - 2 ? 2 * tg.sy : 0);
          const long long mx = x == 1 ? -256LL : (x == FX - 2 ? 256LL : 0LL);
          *reinterpret_cast<uint4*>(dst + off) = u;
          if ((tg.mz | my | mx) != 0) {
#pragma unroll 1
            for (int combo = 1; combo < 8; ++combo) {
              const bool a = combo & 4, bq = combo & 2, cq = combo & 1;
              if ((a && tg.mz == 0) || (bq && my == 0) || (cq && mx == 0)) continue;
              *reinterpret_cast<uint4*>(dst + off + (a ? tg.mz : 0) + (bq ? my : 0) + (cq ? mx : 0)) = u;
            }
          }
        }
      }
      __syncwarp();
    }
    }
  }
}

// ------------------------------------------------------------------------- TMA epilogue
// Same job as ring_epilogue_warp_v2, but every bulk transfer goes through the TMA unit, whose
// shared-memory traffic was measured NOT to slow the MMA stream (the plane / weight loads are
// free), unlike LSU wavefronts (~2 MMA cycles lost per wavefront).  Per warp: one 2 KiB
// SWIZZLE_128B staging box = 2 y rows x 8 x voxels x 64 channels; the 32 accumulator rows of the
// warp go through it in two halves.  Residual halves are TMA-loaded (mbarrier), output halves
// TMA-stored (bulk group), incl. the z mirror (same box, plane +-2).  The y / x REFLECT mirrors
// (rows y = 1, FY-2 / voxels x = 1, FX-2 only) are stored from registers.
struct EpiTma {
  const CUtensorMap* res[2];   // hi, lo (interior view of the padded 16-bit tensor)
  const CUtensorMap* out[2];
  const CUtensorMap* row[2];   // y-halo rows by TMA (nullptr: from registers)
  uint32_t stage_s0, stage_s1; // shared addresses of this warp's staging boxes (1 KiB aligned)
  uint8_t *stage0, *stage1;    // generic pointers to the same
  uint32_t bar0, bar1;         // load barrier of each box
  uint32_t phase0, phase1;     // (scalars, not arrays: a runtime box index must not force the
                               //  struct into local memory)
  int nb;                      // boxes in use: 1, or 2 (residual layers: two transfers in flight)
  bool trace;                  // role timing (CTA 0, first epilogue warp)
  long long t_load, t_store;
  long long t_ph[4];           // tmem load, bias + act, residual, output
  int dbg;
};

// staging box: [32 rows (4 y x 8 x)][32 channels = 64 B], SWIZZLE_64B (16-byte chunk index XOR
// bits 7..8 of the address): conflict-free for thread-per-row 16-byte accesses
__device__ __forceinline__ uint32_t stage64_off(int row, int k) {
  return (uint32_t)(row * 64 + ((k ^ ((row >> 1) & 3)) << 4));
}

template <bool kRes>
__device__ __forceinline__ void ring_epilogue_warp_v3(const ConvGeom& g, const Epilogue& ep,
                                                      const float* sbias, uint32_t t_addr,
                                                      const TileGeom& tg, int plane_coord,
                                                      int mz_planes, EpiTma& et, int lane) {
  const int fmt = ep.fmt;
  const int FY = g.fd[1], FX = g.fd[2];
  const int yl = lane >> 3, xl = lane & 7;
  const int y = tg.y0 + yl, x = tg.x0 + xl;
  const bool row_valid = y < FY && x < FX;
  const bool has_lo = ep.y_lo != nullptr;
  const bool res_has_lo = kRes && ep.res_lo != nullptr;
  const int nb = et.nb;

  // the staging boxes may still be read by the previous tile's stores
  if (lane == 0) {
    const long long ts0 = et.trace ? clock64() : 0;
    tma_store_wait_read();
    if (et.trace) et.t_store += clock64() - ts0;
  }
  __syncwarp();

  // round r of a 4-round sequence: operand r >> 1 (hi, lo), channel half r & 1, box r % nb
  auto issue_res = [&](int r) {
    if (lane == 0) {
      const int bi = nb == 2 ? (r & 1) : 0;
      const uint32_t lb = bi ? et.bar1 : et.bar0;
      mbar_expect_tx(lb, 2048u);
      tma_load_4d(bi ? et.stage_s1 : et.stage_s0, (r >> 1) ? et.res[1] : et.res[0], lb,
                  32 * (r & 1), tg.x0, tg.y0, plane_coord);
    }
  };
  const int n_res = kRes ? (res_has_lo ? 4 : 2) : 0;
  if (kRes) {
    issue_res(0);
    if (nb == 2) issue_res(1);
  }

  const long long tp0 = et.trace ? clock64() : 0;
  float v[64];
  {
    uint32_t raw[64];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
      tmem_ld16(t_addr + cc * 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[cc * 16]));
    tmem_ld_wait();
    if (et.trace) et.t_ph[0] += clock64() - tp0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float4 bv = *reinterpret_cast<const float4*>(sbias + 4 * q);
      v[4 * q] = __uint_as_float(raw[4 * q]) + bv.x;
      v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + bv.y;
      v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + bv.z;
      v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + bv.w;
    }
  }
  if (g.act == S3_ACT_LEAKY) {
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = v[j] >= 0.f ? v[j] : g.alpha * v[j];
  } else if (g.act == S3_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (g.act != S3_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = apply_act_slow(v[j], g.act, g.alpha);
  }

  const long long tp1 = et.trace ? clock64() : 0;
  if (et.trace) et.t_ph[1] += tp1 - tp0;
  if (kRes) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r < n_res) {
        const int bi = nb == 2 ? (r & 1) : 0;
        const long long tl0 = et.trace ? clock64() : 0;
        mbar_wait_lean(bi ? et.bar1 : et.bar0, bi ? et.phase1 : et.phase0);
        if (et.trace) et.t_load += clock64() - tl0;
        if (bi) et.phase1 ^= 1u; else et.phase0 ^= 1u;
        const uint8_t* sb = bi ? et.stage1 : et.stage0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint4 u = *reinterpret_cast<const uint4*>(sb + stage64_off(lane, k));
          unpack_add8(&v[32 * (r & 1) + 8 * k], u, fmt);
        }
        __syncwarp();
        if (r + nb < n_res) issue_res(r + nb);
      }
    }
  }

  const long long tp2 = et.trace ? clock64() : 0;
  if (et.trace) et.t_ph[2] += tp2 - tp1;
  const long long row_off = tg.base + (long long)yl * tg.sy + (long long)xl * 128;
  const long long my = y == 1 ? -2 * tg.sy : (y == FY - 2 ? 2 * tg.sy : 0);
  const long long mx = x == 1 ? -256LL : (x == FX - 2 ? 256LL : 0LL);
  const long long mzb = (long long)mz_planes * tg.sz;
  const int n_out = has_lo ? 4 : 2;
#pragma unroll
  for (int op = 0; op < 2; ++op) {
    if (op == 0 || has_lo) {
      uint4 hrow[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xv = v[8 * k + j];
          a[j] = op == 0 ? xv : xv - from16(to16(xv, fmt), fmt);
        }
        hrow[k].x = pack2(a[0], a[1], fmt);
        hrow[k].y = pack2(a[2], a[3], fmt);
        hrow[k].z = pack2(a[4], a[5], fmt);
        hrow[k].w = pack2(a[6], a[7], fmt);
      }
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int r = 2 * op + c2;
        const int bi = nb == 2 ? (r & 1) : 0;
        // box reuse inside the tile: the store issued nb rounds ago must have read it
        if (r >= nb) {
          if (lane == 0) {
            const long long ts0 = et.trace ? clock64() : 0;
            if (nb == 2) tma_store_wait_read1(); else tma_store_wait_read();
            if (et.trace) et.t_store += clock64() - ts0;
          }
          __syncwarp();
        }
        uint8_t* sb = bi ? et.stage1 : et.stage0;
        const uint32_t sbs = bi ? et.stage_s1 : et.stage_s0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<uint4*>(sb + stage64_off(lane, k)) = hrow[4 * c2 + k];
        if (!(et.dbg & 512)) fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(et.dbg & 2048)) {
          tma_store_4d(et.out[op], sbs, 32 * c2, tg.x0, tg.y0, plane_coord);
          if (mz_planes != 0)
            tma_store_4d(et.out[op], sbs, 32 * c2, tg.x0, tg.y0, plane_coord + mz_planes);
          tma_store_commit();
        }
        (void)n_out;
      }
      // y / x halo mirrors of this row (and their z-mirrored copies)
      if (row_valid && (my | mx) != 0 && !(et.dbg & 1024)) {
        uint8_t* dst = reinterpret_cast<uint8_t*>(op == 0 ? ep.y_hi : ep.y_lo) + row_off;
#pragma unroll 1
        for (int combo = 1; combo < 8; ++combo) {
          const bool a = combo & 4, bq = combo & 2, cq = combo & 1;
          if (!(bq || cq)) continue;
          if ((a && mzb == 0) || (bq && my == 0) || (cq && mx == 0)) continue;
          uint4* d = reinterpret_cast<uint4*>(dst + (a ? mzb : 0) + (bq ? my : 0) + (cq ? mx : 0));
#pragma unroll
          for (int k = 0; k < 8; ++k) d[k] = hrow[k];
        }
      }
    }
  }
  if (et.trace) et.t_ph[3] += clock64() - tp2;
}

// ---------------------------------------------------------- TMA epilogue, 16 warps (V4)
// The V3 epilogue is bound by the serial load -> add -> store chain of each warp (two 32-row
// tiles per warp and item, ~12 k cycles each against 16 k cycles of MMA per item).  V4 runs
// SIXTEEN epilogue warps (one per output plane and TMEM lane quarter, so every warp has exactly
// one tile per item) on 104 registers each: the row is processed in two 32-channel passes
// (v[32] instead of v[64]) and every pass moves through ONE 2 KiB SWIZZLE_64B staging box per
// warp: residual hi, residual lo (TMA loads), output hi, output lo (TMA stores).
template <bool kRes>
__device__ __forceinline__ void ring_epilogue_warp_v4(const ConvGeom& g, const Epilogue& ep,
                                                      const float* sbias, uint32_t t_addr,
                                                      const TileGeom& tg, int plane_coord,
                                                      int mz_planes, EpiTma& et, int lane,
                                                      bool prefetched) {
  const int fmt = ep.fmt;
  const int FY = g.fd[1], FX = g.fd[2];
  const int yl = lane >> 3, xl = lane & 7;
  const int y = tg.y0 + yl, x = tg.x0 + xl;
  const bool row_valid = y < FY && x < FX;
  const bool has_lo = ep.y_lo != nullptr;
  const bool res_has_lo = kRes && ep.res_lo != nullptr;
  uint8_t* const sb = et.stage0;
  const uint32_t sbs = et.stage_s0;
  const bool row_tma = et.row[0] != nullptr;

  auto box_free = [&]() {   // the last store issued by this warp has read the box
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
  };
  auto issue_res = [&](int op, int c2) {
    if (lane == 0) {
      mbar_expect_tx(et.bar0, 2048u);
      tma_load_4d(sbs, op ? et.res[1] : et.res[0], et.bar0, 32 * c2, tg.x0, tg.y0, plane_coord);
    }
  };
  if (!prefetched) {   // (the caller may have done this before waiting for the accumulator)
    box_free();
    if (kRes) issue_res(0, 0);
  }

  const long long row_off = tg.base + (long long)yl * tg.sy + (long long)xl * 128;
  const long long my = y == 1 ? -2 * tg.sy : (y == FY - 2 ? 2 * tg.sy : 0);
  const long long mx = x == 1 ? -256LL : (x == FX - 2 ? 256LL : 0LL);
  const long long mzb = (long long)mz_planes * tg.sz;

#pragma unroll
  for (int c2 = 0; c2 < 2; ++c2) {
    if (kRes && c2 == 1) {
      box_free();
      issue_res(0, 1);
    }
    float v[32];
    {
      uint32_t raw[32];
      tmem_ld16(t_addr + 32 * c2, *reinterpret_cast<uint32_t(*)[16]>(&raw[0]));
      tmem_ld16(t_addr + 32 * c2 + 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[16]));
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 bv = *reinterpret_cast<const float4*>(sbias + 32 * c2 + 4 * q);
        v[4 * q] = __uint_as_float(raw[4 * q]) + bv.x;
        v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + bv.y;
        v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + bv.z;
        v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + bv.w;
      }
    }
    if (g.act == S3_ACT_LEAKY) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = v[j] >= 0.f ? v[j] : g.alpha * v[j];
    } else if (g.act == S3_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (g.act != S3_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = apply_act_slow(v[j], g.act, g.alpha);
    }
    if (kRes) {
#pragma unroll
      for (int op = 0; op < 2; ++op) {
        if (op == 0 || res_has_lo) {
          mbar_wait_lean(et.bar0, et.phase0);
          et.phase0 ^= 1u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint4 u = *reinterpret_cast<const uint4*>(sb + stage64_off(lane, k));
            unpack_add8(&v[8 * k], u, fmt);
          }
          __syncwarp();
          if (op == 0 && res_has_lo) issue_res(1, c2);
        }
      }
    }
#pragma unroll
    for (int op = 0; op < 2; ++op) {
      if (op == 0 || has_lo) {
        uint4 hrow[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float a[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xv = v[8 * k + j];
            a[j] = op == 0 ? xv : xv - from16(to16(xv, fmt), fmt);
          }
          hrow[k].x = pack2(a[0], a[1], fmt);
          hrow[k].y = pack2(a[2], a[3], fmt);
          hrow[k].z = pack2(a[4], a[5], fmt);
          hrow[k].w = pack2(a[6], a[7], fmt);
        }
        if (op == 1 || (c2 == 1 && !kRes)) box_free();   // (after a residual round the box is free)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<uint4*>(sb + stage64_off(lane, k)) = hrow[k];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(et.out[op], sbs, 32 * c2, tg.x0, tg.y0, plane_coord);
          if (mz_planes != 0)
            tma_store_4d(et.out[op], sbs, 32 * c2, tg.x0, tg.y0, plane_coord + mz_planes);
          if (row_tma) {
            // REFLECT halo rows y = -1 (copy of y = 1) and y = FY (copy of FY - 2): one 8-voxel
            // row of the box each, stored at the padded row index (and its z mirror)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int yy = e == 0 ? 1 : FY - 2;
              if (yy >= tg.y0 && yy < tg.y0 + 4) {
                const uint32_t src = sbs + (uint32_t)(yy - tg.y0) * 512u;
                const int ypad = e == 0 ? 0 : FY + 1;
                tma_store_4d(et.row[op], src, 32 * c2, tg.x0 + 1, ypad, plane_coord);
                if (mz_planes != 0)
                  tma_store_4d(et.row[op], src, 32 * c2, tg.x0 + 1, ypad, plane_coord + mz_planes);
              }
            }
          }
          tma_store_commit();
        }
        if (row_valid && (row_tma ? mx : (my | mx)) != 0) {
          uint8_t* dst = reinterpret_cast<uint8_t*>(op == 0 ? ep.y_hi : ep.y_lo) + row_off + 64 * c2;
#pragma unroll 1
          for (int combo = 1; combo < 8; ++combo) {
            const bool a = combo & 4, bq = combo & 2, cq = combo & 1;
            if (!(bq || cq) || (row_tma && !cq)) continue;
            if ((a && mzb == 0) || (bq && my == 0) || (cq && mx == 0)) continue;
            uint4* d = reinterpret_cast<uint4*>(dst + (a ? mzb : 0) + (bq ? my : 0) + (cq ? mx : 0));
#pragma unroll
            for (int k = 0; k < 4; ++k) d[k] = hrow[k];
          }
        }
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
    ring_epilogue_row64(g, p.ep, sbias, ta, rp, p.dbg_flags);
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

// MMA role for the hot configuration (every item has R = 4 output planes, npad = 64, 18 x 10
// voxel planes, 7 ring slots, 2 weight slots).  Measured on B200 (tools/microbench/mma_rate4.cu
// and the role trace): the tcgen05 queue only holds ~2 MMAs beyond the executing one, so every
// stretch of issue-side code between two MMAs that is longer than ~2 MMA durations is a tensor
// pipe bubble (the per-slab loop overhead of the generic role, ~430-580 cycles, cost ~20 %).
// Here one elected thread issues an entire item (9 slabs x 6 planes x 4 k-steps) as straight
// line code; the barrier waits for the NEXT slab / item are taken before the last MMA group of
// the current slab (a N = 192 group, 4 x 96 cycles), so that the first MMA of the next slab
// follows the last one of this slab back to back.
template <int kP>
__device__ __forceinline__ void ring_mma_fast(const UmmaParams& p, uint32_t bar_base,
                                              uint32_t a_base, uint32_t w_base, uint32_t w_slab,
                                              uint32_t tmem_base, int i0, int i1, int lane) {
  constexpr int kR = 4;
  constexpr uint32_t kPlaneLo = (18u * 10u * 128u) >> 4, kBlkLo = (64u * 128u) >> 4;
  // plane order inside a slab: slab 0 ascending (each accumulator block is zero-initialised by
  // its dz = 0 contribution); other slabs end with a N = 192 group; the last slab releases the
  // planes the producer needs first (0, 1, 2) early
  // (with 6 slots every slot is needed again by the next item: release 0..3 in order)
  constexpr int kOrdMid[6] = {0, 1, 4, 5, 2, 3};
  constexpr int kOrdLast[6] = {0, 1, 2, kP == 6 ? 3 : 4, 5, kP == 6 ? 4 : 3};
  auto bar = [&](int i) { return bar_base + 8u * i; };
  const uint32_t fmtb = p.fmt == 0 ? 1u : 0u;
  const uint32_t hi_a = sdesc_hi_sw128(1280u), hi_b = sdesc_hi_sw128(1024u);
  const uint32_t id1 = make_idesc_f16(64u, fmtb), id2 = make_idesc_f16(128u, fmtb),
                 id3 = make_idesc_f16(192u, fmtb);
  const uint32_t a_lo0 = sdesc_lo(a_base);
  const uint32_t w_lo0 = sdesc_lo(w_base), w_lo1 = sdesc_lo(w_base + w_slab);
  int g = 0;   // weight slabs consumed so far: slot g & 1, parity (g >> 1) & 1
  int ab = 0, abph = 0, slot0 = 0;
  uint32_t pf_phase = 0;
  const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0;
  const long long t_all0 = tr ? clock64() : 0;
  for (int i = i0; i < i1; ++i) {
    const RingItem c = ring_decode(p, i, i0, i1);
    const bool has_next = i + 1 < i1;
    const uint32_t acc0 = tmem_base + (uint32_t)(ab * kR * 64);
    const uint32_t next_acc_par = (uint32_t)((ab == 1 ? abph ^ 1 : abph) ^ 1);
    uint32_t al[kR + 2], pfb[kR + 2], pfp[kR + 2], peb[kR + 2];
#pragma unroll
    for (int ip = 0; ip < kR + 2; ++ip) {
      int slot = slot0 + ip;
      if (slot >= kP) slot -= kP;
      al[ip] = a_lo0 + (uint32_t)slot * kPlaneLo;
      pfb[ip] = bar(RB_PFULL + slot);
      peb[ip] = bar(RB_PEMPTY + slot);
      pfp[ip] = (pf_phase >> slot) & 1u;
      if (!(c.cont && ip < 2)) pf_phase ^= 1u << slot;
    }
    if (elect_one()) {
      if (i == i0) {
        mbar_wait_lean(bar(RB_ACCEMPTY + ab), (uint32_t)(abph ^ 1));
        mbar_wait_lean(bar(RB_WFULL + 0), 0u);
        tc_fence_after();
      }
#pragma unroll
      for (int s = 0; s < 9; ++s) {
        const int gs = g + s;
        const uint32_t wl = (gs & 1) ? w_lo1 : w_lo0;
        const uint32_t tap = (uint32_t)((s / 3) * 80 + (s % 3) * 8);
#pragma unroll
        for (int q = 0; q < kR + 2; ++q) {
          const int ip = s == 0 ? q : (s == 8 ? kOrdLast[q] : kOrdMid[q]);
          const int jlo = ip - (kR - 1) > 0 ? ip - (kR - 1) : 0;
          const int jhi = ip < 2 ? ip : 2;
          const int nblk = jhi - jlo + 1;
          const uint32_t dcol = acc0 + (uint32_t)(64 * (kR - 1 - (ip - jlo)));
          const uint32_t idn = nblk == 3 ? id3 : (nblk == 2 ? id2 : id1);
          const uint32_t a0 = al[ip] + tap;
          const uint32_t b0 = wl + (uint32_t)jlo * kBlkLo;
          if (q == kR + 1) {
            // operands of the next slab / item before the last group of this slab goes out
            if (s < 8) {
              mbar_wait_lean(bar(RB_WFULL + ((gs + 1) & 1)), (uint32_t)(((gs + 1) >> 1) & 1));
            } else if (has_next) {
              mbar_wait_lean(bar(RB_ACCEMPTY + (ab ^ 1)), next_acc_par);
              mbar_wait_lean(bar(RB_WFULL + ((gs + 1) & 1)), (uint32_t)(((gs + 1) >> 1) & 1));
              tc_fence_after();
            }
          }
          if (s == 0) {
            if (!(c.cont && ip < 2)) mbar_wait_lean(pfb[ip], pfp[ip]);
            if (jlo == 0) {
              umma_f16_new(dcol, mk_desc(a0, hi_a), mk_desc(wl, hi_b), id1);
              if (nblk > 1)
                umma_f16_acc(dcol + 64u, mk_desc(a0, hi_a), mk_desc(wl + kBlkLo, hi_b),
                             nblk == 3 ? id2 : id1);
#pragma unroll
              for (int kk = 1; kk < 4; ++kk)
                umma_f16_acc(dcol, mk_desc(a0 + 2u * kk, hi_a), mk_desc(b0 + 2u * kk, hi_b), idn);
            } else {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_acc(dcol, mk_desc(a0 + 2u * kk, hi_a), mk_desc(b0 + 2u * kk, hi_b), idn);
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16_acc(dcol, mk_desc(a0 + 2u * kk, hi_a), mk_desc(b0 + 2u * kk, hi_b), idn);
          }
          if (s == 8 && !(c.next_cont && ip >= kR)) umma_commit(peb[ip]);
        }
        umma_commit(bar(RB_WEMPTY + (gs & 1)));
        if (s == 8) umma_commit(bar(RB_ACCFULL + ab));
      }
    }
    __syncwarp();
    g += 9;
    slot0 += c.next_cont ? kR : kR + 2;
    while (slot0 >= kP) slot0 -= kP;
    if (++ab == 2) { ab = 0; abph ^= 1; }
  }
  if (tr) {
    p.trace[0] = clock64() - t_all0; p.trace[1] = 0; p.trace[2] = 0; p.trace[3] = 0;
    p.trace[4] = 0; p.trace[5] = i1 - i0;
  }
}

// ------------------------------------------------------------------------------------ kernel
// kR > 0: compile-time planes per item (fully unrolled issue loop); kR == 0: runtime p.R
template <int kR, int EPI>
__global__ void __launch_bounds__(ring_threads(EPI), 1)
conv_umma_zring_kernel(const __grid_constant__ CUtensorMap tm_a,
                       const __grid_constant__ CUtensorMap tm_w,
                       const __grid_constant__ EpiMaps em, const UmmaParams p) {
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
      mbar_init(bar(RB_ACCEMPTY + i), EPI == EPI_V4 ? 16 : 8);
    }
    for (int i = 0; i < 16; ++i) mbar_init(bar(RB_EPILD + i), 1);
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
    } else if (warp == 1 && kR == 4 && p.ring_fast) {
      if (P == 6) ring_mma_fast<6>(p, bar_base, a_base, w_base, w_slab, tmem_base, i0, i1, lane);
      else ring_mma_fast<7>(p, bar_base, a_base, w_base, w_slab, tmem_base, i0, i1, lane);
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
    if (EPI == EPI_V4) reg_inc<104>(); else reg_inc<216>();
    constexpr int kWgs = EPI == EPI_V4 ? 4 : 2;   // epilogue warpgroups
    const int wg = (warp - 4) >> 2, q = warp & 3;
    int ab = 0, abph = 0;
    EpiTma et;
    et.res[0] = &em.res_hi; et.res[1] = &em.res_lo;
    et.out[0] = &em.y_hi; et.out[1] = &em.y_lo;
    et.row[0] = p.epi_row_tma ? &em.row_hi : nullptr;
    et.row[1] = p.epi_row_tma ? &em.row_lo : nullptr;
    et.nb = EPI == EPI_V4 ? 1 : p.epi_bufs;
    et.stage_s0 = bar_base + 2048u + (uint32_t)((warp - 4) * et.nb) * 2048u;
    et.stage_s1 = et.stage_s0 + 2048u;
    et.stage0 = smem_raw + (et.stage_s0 - smem_u32(smem_raw));
    et.stage1 = et.stage0 + 2048;
    et.bar0 = bar(RB_EPILD + (EPI == EPI_V4 ? 1 : 2) * (warp - 4));
    et.bar1 = et.bar0 + 8u;
    et.phase0 = et.phase1 = 0;
    et.trace = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 128;
    et.t_load = et.t_store = 0;
    et.dbg = p.dbg_flags;
    et.t_ph[0] = et.t_ph[1] = et.t_ph[2] = et.t_ph[3] = 0;
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 128;
    long long t_wait = 0, t_work = 0;
    for (int i = i0; i < i1; ++i) {
      const RingItem c = ring_decode(p, i, i0, i1);
      long long c0 = tr ? clock64() : 0;
      bool prefetched = false;
      if (EPI == EPI_V4 && wg < c.ri && !(p.dbg_flags & 8)) {
        // the residual tile does not depend on the accumulator: start its first TMA load (and
        // retire the previous tile's stores) before waiting for the MMAs of this item
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        if (p.ep.res_hi && lane == 0) {
          const ConvGeom& g = p.g;
          const int plane_coord = c.b * (g.fd[0] + 2) + c.pl0 + wg + 1;
          mbar_expect_tx(et.bar0, 2048u);
          tma_load_4d(et.stage_s0, et.res[0], et.bar0, 0, c.xb * 8, c.yb * 16 + q * 4, plane_coord);
        }
        prefetched = true;
      }
      mbar_wait_inl(bar(RB_ACCFULL + ab), abph, p.dbg, 6, ab, i - i0);
      long long c1 = tr ? clock64() : 0;
      if (tr) t_wait += c1 - c0;
      tc_fence_after();
      if (EPI == EPI_V3 || EPI == EPI_V4) {
        if (!(p.dbg_flags & 8)) {
          const ConvGeom& g = p.g;
          TileGeom tg;
          tg.sy = (long long)(g.fd[2] + 2) * 128;
          tg.sz = (long long)(g.fd[1] + 2) * tg.sy;
          tg.y0 = c.yb * 16 + q * 4;
          tg.x0 = c.xb * 8;
          tg.mz = 0;
          for (int r = wg; r < c.ri; r += kWgs) {
            const int z = c.pl0 + r;
            const int plane_coord = c.b * (g.fd[0] + 2) + z + 1;
            tg.base = (((long long)plane_coord * (g.fd[1] + 2) + tg.y0 + 1) * (g.fd[2] + 2) +
                       tg.x0 + 1) * 128;
            const int mzp = z == 1 ? -2 : (z == g.fd[0] - 2 ? 2 : 0);
            const uint32_t ta = tmem_base + (uint32_t)(ab * R * npad + npad * (R - 1 - r)) +
                                ((uint32_t)(q * 32) << 16);
            if (EPI == EPI_V4) {
              if (p.ep.res_hi)
                ring_epilogue_warp_v4<true>(g, p.ep, sbias, ta, tg, plane_coord, mzp, et, lane,
                                            prefetched);
              else
                ring_epilogue_warp_v4<false>(g, p.ep, sbias, ta, tg, plane_coord, mzp, et, lane,
                                             prefetched);
            } else if (p.ep.res_hi)
              ring_epilogue_warp_v3<true>(g, p.ep, sbias, ta, tg, plane_coord, mzp, et, lane);
            else
              ring_epilogue_warp_v3<false>(g, p.ep, sbias, ta, tg, plane_coord, mzp, et, lane);
          }
        }
      } else if (EPI == EPI_V2) {
        if (!(p.dbg_flags & 8)) {
        const ConvGeom& g = p.g;
        uint8_t* stage = smem_raw + (bar_base + 2048u - smem_u32(smem_raw)) + (warp - 4) * 2048;
        TileGeom tg;
        tg.sy = (long long)(g.fd[2] + 2) * 128;
        tg.sz = (long long)(g.fd[1] + 2) * tg.sy;
        tg.y0 = c.yb * 16 + q * 4;
        tg.x0 = c.xb * 8;
        for (int r = wg; r < c.ri; r += 2) {
          const int z = c.pl0 + r;
          tg.base = ((((long long)c.b * (g.fd[0] + 2) + z + 1) * (g.fd[1] + 2) + tg.y0 + 1) *
                         (g.fd[2] + 2) + tg.x0 + 1) * 128;
          tg.mz = z == 1 ? -2 * tg.sz : (z == g.fd[0] - 2 ? 2 * tg.sz : 0);
          const uint32_t ta = tmem_base + (uint32_t)(ab * R * npad + npad * (R - 1 - r)) +
                              ((uint32_t)(q * 32) << 16);
          if (p.ep.res_hi)
            ring_epilogue_warp_v2<true>(g, p.ep, sbias, ta, tg, stage, lane);
          else
            ring_epilogue_warp_v2<false>(g, p.ep, sbias, ta, tg, stage, lane);
        }
        }
      } else if (!(p.dbg_flags & 8)) {
        for (int r = wg; r < c.ri; r += 2)
          ring_epilogue_tile<(EPI == EPI_V2 || EPI == EPI_V3 || EPI == EPI_V4) ? EPI_PLAIN : EPI>(
              p, sbias, c, r, tmem_base + (uint32_t)(ab * R * npad + npad * (R - 1 - r)), q, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(RB_ACCEMPTY + ab));
      if (++ab == 2) { ab = 0; abph ^= 1; }
      if (tr) t_work += clock64() - c1;
    }
    // the staging boxes must outlive the TMA stores that read them: retire this thread's bulk
    // groups before the CTA (and its shared memory) goes away
    if ((EPI == EPI_V3 || EPI == EPI_V4) && lane == 0) tma_store_wait_read();
    if (tr) { p.trace[8] = t_wait; p.trace[9] = t_work; p.trace[10] = et.t_load; p.trace[11] = et.t_store;
              for (int k = 0; k < 4; ++k) p.trace[12 + k] = et.t_ph[k]; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int kR, int EPI>
static int launch_zring_t(const UmmaParams& p, const CUtensorMap& a, const CUtensorMap& w,
                          const EpiMaps& em, int ctas, uint32_t smem, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    S3_CUDA(cudaFuncSetAttribute(conv_umma_zring_kernel<kR, EPI>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr = true;
  }
  conv_umma_zring_kernel<kR, EPI><<<ctas, ring_threads(EPI), smem, st>>>(a, w, em, p);
  S3_CUDA(cudaGetLastError());
  return S3_OK;
}

int launch_umma_zring(const UmmaParams& p, const CUtensorMap& a, const CUtensorMap& w,
                      const CUtensorMap* epi_maps, int epi, int ctas, uint32_t smem,
                      cudaStream_t st) {
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  if (epi_maps) {
    em.res_hi = epi_maps[0]; em.res_lo = epi_maps[1]; em.y_hi = epi_maps[2]; em.y_lo = epi_maps[3];
    em.row_hi = epi_maps[4]; em.row_lo = epi_maps[5];
  }
  if (p.epi_v2 == 3 && p.R == 4) return launch_zring_t<4, EPI_V4>(p, a, w, em, ctas, smem, st);
  if (p.epi_v2 == 2 && p.R == 4) return launch_zring_t<4, EPI_V3>(p, a, w, em, ctas, smem, st);
  if (p.epi_v2 == 1 && p.R == 4) return launch_zring_t<4, EPI_V2>(p, a, w, em, ctas, smem, st);
  if (epi == EPI_PLAIN && p.R == 4) return launch_zring_t<4, EPI_PLAIN>(p, a, w, em, ctas, smem, st);
  if (epi == EPI_PLAIN) return launch_zring_t<0, EPI_PLAIN>(p, a, w, em, ctas, smem, st);
  return launch_zring_t<0, EPI_GENERIC>(p, a, w, em, ctas, smem, st);
}

}  // namespace s3
