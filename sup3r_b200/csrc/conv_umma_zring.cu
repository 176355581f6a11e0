// tcgen05 convolution, "zring" scheme for narrow (64-channel) 3-D outputs.
// MMA arrangement ("z-concatenated N"): one (dy, dx) weight slab holds the three dz taps stacked
// along N ([3 x 64 rows][64 ch], 24 KB), so ONE tcgen05.mma (M = 128 voxels = 16 y x 8 x of an
// INPUT plane ip, K = 16) feeds OUTPUT planes ip, ip-1, ip-2 at once (N = 192; 64 / 128 on the
// edge planes of an item).  Around it a software pipeline that never drains between work items:
//   * activations live in a RING of P plane slots (one TMA box + full/empty mbarrier pair per
//     plane).  A CTA owns a contiguous range of work items ordered z-fastest inside a
//     (batch, y block, x block) column, so consecutive items share two input planes (no z-halo
//     re-read) and the planes of item i+1 stream in while the last slab of item i is still
//     being multiplied; the MMA warp waits per plane, not per item;
//   * weight slabs stream through a 2-deep ring (one slab is consumed for ~1.8 k cycles, the
//     next one lands meanwhile);
//   * accumulators are double-buffered in TMEM and drained by SIXTEEN epilogue warps on 104
//     registers each (V4: 16-bit output, TMA tile I/O) or eight on 216 (thread-per-row: f32
//     destinations, nearest-repeat layers), while the producer / issuer warpgroup keeps 64
//     (setmaxnreg);
//   * two-pass operand formats (fp16c, common.cuh): every item runs a kind::f16 pass over the
//     fp16 tensors and a kind::f8f6f4 pass over the e4m3 corr tensors into the same accumulator;
//     both passes stream through the same plane / weight slots.
// Reference semantics: FlexiblePadding(3, REFLECT) -> Conv3D(valid) -> Cropping3D(2)
// [-> LeakyReLU] [-> nearest repeat] [-> SkipConnection add] as executed by
// sup3r/models/abstract.py:1081-1092 over sup3r/configs/spatiotemporal/gen_*.json.
#include "conv_umma_common.cuh"

namespace s3 {

constexpr int kRingThreads = 384;    // WG0: TMA + MMA (+2 idle warps); WG1, WG2: epilogue
constexpr int kRingThreadsV4 = 640;  // ... WG1..WG4: sixteen epilogue warps (EPI_V4)
__host__ __device__ constexpr int ring_threads(int epi) { return epi_is_v4(epi) ? kRingThreadsV4 : kRingThreads; }
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
  // (two-pass formats stream both operand tensors through the same slots: no planes carry over)
  c.cont = p.npass == 1 && i > i0 && c.grp > 0;
  c.next_cont = p.npass == 1 && (i + 1 < i1) && (c.grp + 1 < G);
  return c;
}

// ---------------------------------------------------------------------------------- epilogue
// One thread = one output voxel (64 fp32 accumulator columns).  The residual row (if any) is
// requested before the TMEM loads, the whole 64-channel row is finished in registers and then
// written with back-to-back 16-byte stores (full 128-B lines per voxel).  Used by the
// 384-thread variant: nearest-repeat layers, f32 destinations, f32 residuals.
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
  const float sc = ep.acc_scale;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float4 bv = *reinterpret_cast<const float4*>(sbias + 4 * q);
    v[4 * q] = fmaf(__uint_as_float(raw[4 * q]), sc, bv.x);
    v[4 * q + 1] = fmaf(__uint_as_float(raw[4 * q + 1]), sc, bv.y);
    v[4 * q + 2] = fmaf(__uint_as_float(raw[4 * q + 2]), sc, bv.z);
    v[4 * q + 3] = fmaf(__uint_as_float(raw[4 * q + 3]), sc, bv.w);
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
    uint4 h[8], l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      h[q].x = pack2(v[8 * q], v[8 * q + 1], fmt);
      h[q].y = pack2(v[8 * q + 2], v[8 * q + 3], fmt);
      h[q].z = pack2(v[8 * q + 4], v[8 * q + 5], fmt);
      h[q].w = pack2(v[8 * q + 6], v[8 * q + 7], fmt);
    }
    if (yl) {
      if (fmt == kFmtFp16c) {
        // corr row: per 32-channel half [lo8 ch 0-15 | lo8 ch 16-31 | a8 ch 0-15 | a8 ch 16-31]
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          corr16(&v[32 * hf], l[4 * hf], l[4 * hf + 2]);
          corr16(&v[32 * hf + 16], l[4 * hf + 1], l[4 * hf + 3]);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float e[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) e[j] = v[8 * q + j] - from16(to16(v[8 * q + j], fmt), fmt);
          l[q].x = pack2(e[0], e[1], fmt);
          l[q].y = pack2(e[2], e[3], fmt);
          l[q].z = pack2(e[4], e[5], fmt);
          l[q].w = pack2(e[6], e[7], fmt);
        }
      }
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
          for (int q = 0; q < 8; ++q) dl[q] = l[q];
        }
        if ((rp.mz | rp.my | mx) == 0) break;
      }
    }
  }
}

// ------------------------------------------------------------------------- TMA epilogue
// Measured (role trace, B200): global loads / stores issued thread-per-row (32 different
// 128-B lines per warp instruction) slow the concurrently running MMA stream almost 1:1 with
// their L1 wavefront count -- the SS-mode tcgen05.mma already uses ~85 % of the shared-memory
// bandwidth and the LSU shares that data path -- whereas shared-memory traffic of the TMA unit
// does not (the plane / weight loads are free).  Every bulk transfer of the epilogue therefore
// goes through TMA: residual tiles are TMA-loaded (mbarrier), output tiles TMA-stored (bulk
// group), incl. the z mirror (same box, plane +-2) and the y halo rows; only the x REFLECT
// mirrors (voxels x = 1, FX-2) are stored from registers.
struct TileGeom {
  long long sy, sz;      // byte strides of the padded 16-bit tensor along y / z
  long long base;        // byte offset of voxel (y0, x0) of this plane (interior position)
  int y0, x0;
};

__device__ __forceinline__ void unpack_add8(float* v, const uint4& u, int fmt) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] += from16((uint16_t)(w[j] & 0xffffu), fmt);
    v[2 * j + 1] += from16((uint16_t)(w[j] >> 16), fmt);
  }
}

struct EpiTma {
  const CUtensorMap* res[2];   // hi, lo (interior view of the padded 16-bit tensor)
  const CUtensorMap* out[2];
  const CUtensorMap* row[2];   // y-halo rows by TMA (nullptr: from registers)
  uint32_t stage_s0;           // shared address of this warp's staging box (1 KiB aligned)
  uint8_t* stage0;             // generic pointer to the same
  uint32_t bar0;               // load barrier of the box
  uint32_t phase0;
  int dbg;                     // experiment flags (timing runs only, results are wrong)
};

// staging box: [32 rows (4 y x 8 x)][32 channels = 64 B], SWIZZLE_64B (16-byte chunk index XOR
// bits 7..8 of the address): conflict-free for thread-per-row 16-byte accesses
__device__ __forceinline__ uint32_t stage64_off(int row, int k) {
  return (uint32_t)(row * 64 + ((k ^ ((row >> 1) & 3)) << 4));
}

// SIXTEEN epilogue warps (one per output plane and TMEM lane quarter, so every warp has exactly
// one 32-row tile per item) on 104 registers each: the 64-channel row is processed in two
// 32-channel passes and every pass moves through ONE 2 KiB SWIZZLE_64B staging box per warp:
// residual hi, residual lo (TMA loads), output hi, output lo (TMA stores).  In the fp16c format
// the "lo" 64 bytes of a 32-channel half are [lo8 x 32 | a8 x 32] (common.cuh), so the same
// boxes and tensor maps serve both formats.
template <bool kRes, bool kRep>
__device__ __forceinline__ void ring_epilogue_warp_v4(const ConvGeom& g, const Epilogue& ep,
                                                      const float* sbias, uint32_t t_addr,
                                                      const TileGeom& tg, int plane_coord,
                                                      int mz_planes, EpiTma& et, int lane,
                                                      bool prefetched) {
  const int fmt = ep.fmt;
  const bool corr = fmt == kFmtFp16c;
  // nearest repeat along x (SpatioTemporalExpansion, temporal_method "nearest"): conv voxel x
  // becomes output voxels x * rep + rx; the tile is stored once per replica through 5-D maps
  // whose x axis is split into (x, rx)
  const int rep = kRep ? g.rep[2] : 1;
  const int FY = g.fd[1], FX = g.in[2];
  const int yl = lane >> 3, xl = lane & 7;
  const int y = tg.y0 + yl, x = tg.x0 + xl;
  const bool row_valid = y < FY && x < FX;
  const bool has_lo = ep.y_lo != nullptr;
  const bool res_has_lo = kRes && ep.res_lo != nullptr;
  uint8_t* const sb = et.stage0;
  const uint32_t sbs = et.stage_s0;
  const bool row_tma = et.row[0] != nullptr;
  const float sc = ep.acc_scale;

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

  const long long row_off = tg.base + (long long)yl * tg.sy + (long long)xl * rep * 128;
  const long long my = y == 1 ? -2 * tg.sy : (y == FY - 2 ? 2 * tg.sy : 0);
  // REFLECT halo in x: output voxel 1 -> -1 and FX_out - 2 -> FX_out.  With a repeat >= 2 those
  // are replicas of the first / last conv voxel: its row goes one voxel before its first /
  // after its last replica.
  const long long mx = !kRep ? (x == 1 ? -256LL : (x == FX - 2 ? 256LL : 0LL))
                             : (x == 0 ? -128LL : (x == FX - 1 ? 128LL * rep : 0LL));
  auto store_tile = [&](const CUtensorMap* m, uint32_t src, int c, int xc, int yc, int pc,
                        bool padded_x) {
    if (!kRep) {
      tma_store_4d(m, src, c, xc + (padded_x ? 1 : 0), yc, pc);
    } else {
#pragma unroll 1
      for (int rx = 0; rx < rep; ++rx) tma_store_5d(m, src, c, rx, xc, yc, pc);
    }
  };
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
        v[4 * q] = fmaf(__uint_as_float(raw[4 * q]), sc, bv.x);
        v[4 * q + 1] = fmaf(__uint_as_float(raw[4 * q + 1]), sc, bv.y);
        v[4 * q + 2] = fmaf(__uint_as_float(raw[4 * q + 2]), sc, bv.z);
        v[4 * q + 3] = fmaf(__uint_as_float(raw[4 * q + 3]), sc, bv.w);
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
          if (op == 1 && corr) {
            // chunks 0, 1 = lo8 of the 32 channels (chunks 2, 3 = a8: an MMA operand only)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint4 u = *reinterpret_cast<const uint4*>(sb + stage64_off(lane, k));
              corr_add16(&v[16 * k], u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint4 u = *reinterpret_cast<const uint4*>(sb + stage64_off(lane, k));
              unpack_add8(&v[8 * k], u, fmt);
            }
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
        if (op == 1 && corr) {
          corr16(&v[0], hrow[0], hrow[2]);
          corr16(&v[16], hrow[1], hrow[3]);
        } else {
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
        }
        if (op == 1 || (c2 == 1 && !kRes)) box_free();   // (after a residual round the box is free)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<uint4*>(sb + stage64_off(lane, k)) = hrow[k];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(et.dbg & 64)) {
          store_tile(et.out[op], sbs, 32 * c2, tg.x0, tg.y0, plane_coord, false);
          if (mz_planes != 0)
            store_tile(et.out[op], sbs, 32 * c2, tg.x0, tg.y0, plane_coord + mz_planes, false);
          if (row_tma) {
            // REFLECT halo rows y = -1 (copy of y = 1) and y = FY (copy of FY - 2): one 8-voxel
            // row of the box each, stored at the padded row index (and its z mirror)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int yy = e == 0 ? 1 : FY - 2;
              if (yy >= tg.y0 && yy < tg.y0 + 4) {
                const uint32_t src = sbs + (uint32_t)(yy - tg.y0) * 512u;
                const int ypad = e == 0 ? 0 : FY + 1;
                store_tile(et.row[op], src, 32 * c2, tg.x0, ypad, plane_coord, true);
                if (mz_planes != 0)
                  store_tile(et.row[op], src, 32 * c2, tg.x0, ypad, plane_coord + mz_planes, true);
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
    ring_epilogue_row64(g, p.ep, sbias, ta, rp);
  else
    epilogue_row<EPI>(g, p.ep, sbias, ta, rp);
}

// One (item, pass) stage of the hot configuration (every item has R = 4 output planes,
// npad = 64, 18 x 10 voxel planes, 2 weight slots) issued by one elected thread as straight-line
// code (9 slabs x 6 planes x 4 k-steps).  Measured on B200 (tools/microbench/mma_rate4.cu and the
// role trace): the tcgen05 queue only holds ~2 MMAs beyond the executing one, so every stretch
// of issue-side code between two MMAs that is longer than ~2 MMA durations is a tensor pipe
// bubble (the per-slab loop overhead of the generic role, ~430-580 cycles, cost ~20 %).  The
// barrier waits for the NEXT slab / stage are taken before the last MMA group of the current
// slab (a N = 192 group, 4 x 96 cycles), so that the first MMA of the next slab follows the last
// one of this slab back to back.
//   kF8    : kind::f8f6f4 MMAs on the corr tensors (second pass of the fp16c format); measured
//            (tools/microbench/f8_probe.cu): same cycles per MMA as kind::f16 at K = 32, no
//            penalty for alternating kinds, same accumulator
//   kFirst : this stage zero-initialises the accumulators (first pass of an item)
//   kStatic: the stage's input planes sit in ring slots 0 .. 5 (P = 6 and no carried planes, i.e.
//            every stage of a two-pass format): descriptor / barrier addresses are immediates
//            off one base register and all six FULL barriers share one parity -- the issue
//            role runs on 64 registers (setmaxnreg), the per-plane arrays of the general form
//            do not fit next to the straight-line code's temporaries
struct RingStage {
  uint32_t al[6], pfb[6], pfp[6], peb[6];   // per input plane: descriptor low word, barriers
  uint32_t a_lo0, bar_base, par;            // kStatic form of the same
  uint32_t w_lo0, w_lo1;                    // weight slot descriptors
  uint32_t acc0;
  uint32_t bar_wfull, bar_wempty, bar_accfull, bar_accempty_next;
  uint32_t next_acc_par;
  int g;                                    // weight slabs consumed before this stage
  bool cont, next_cont, last_pass, has_next;
};

template <int kP, bool kF8, bool kFirst, bool kStatic>
__device__ __forceinline__ void ring_issue_stage(const RingStage& st, uint32_t hi_a, uint32_t hi_b,
                                                 uint32_t id1, uint32_t id2, uint32_t id3) {
  constexpr int kR = 4;
  constexpr uint32_t kBlkLo = (64u * 128u) >> 4, kPlaneLo = (18u * 10u * 128u) >> 4;
  // plane order inside a slab: slab 0 ascending (each accumulator block is zero-initialised by
  // its dz = 0 contribution); other slabs end with a N = 192 group; the last slab releases the
  // planes the producer needs first (0, 1, 2) early
  // (with 6 slots every slot is needed again by the next stage: release 0..3 in order)
  constexpr int kOrdMid[6] = {0, 1, 4, 5, 2, 3};
  constexpr int kOrdLast[6] = {0, 1, 2, kP == 6 ? 3 : 4, 5, kP == 6 ? 4 : 3};
  auto mma_acc = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
    if (kF8) umma_f8_acc(d, a, b, id); else umma_f16_acc(d, a, b, id);
  };
#pragma unroll
  for (int s = 0; s < 9; ++s) {
    const int gs = st.g + s;
    const uint32_t wl = (gs & 1) ? st.w_lo1 : st.w_lo0;
    const uint32_t tap = (uint32_t)((s / 3) * 80 + (s % 3) * 8);
#pragma unroll
    for (int q = 0; q < kR + 2; ++q) {
      const int ip = s == 0 ? q : (s == 8 ? kOrdLast[q] : kOrdMid[q]);
      const int jlo = ip - (kR - 1) > 0 ? ip - (kR - 1) : 0;
      const int jhi = ip < 2 ? ip : 2;
      const int nblk = jhi - jlo + 1;
      const uint32_t dcol = st.acc0 + (uint32_t)(64 * (kR - 1 - (ip - jlo)));
      const uint32_t idn = nblk == 3 ? id3 : (nblk == 2 ? id2 : id1);
      const uint32_t a0 = (kStatic ? st.a_lo0 + (uint32_t)ip * kPlaneLo : st.al[ip]) + tap;
      const uint32_t b0 = wl + (uint32_t)jlo * kBlkLo;
      if (q == kR + 1) {
        // operands of the next slab / stage before the last group of this slab goes out
        const uint32_t nb = st.bar_wfull + 8u * ((gs + 1) & 1);
        const uint32_t np = (uint32_t)(((gs + 1) >> 1) & 1);
        if (s < 8 || !st.last_pass) {
          mbar_wait_lean(nb, np);
        } else if (st.has_next) {
          mbar_wait_lean(st.bar_accempty_next, st.next_acc_par);
          mbar_wait_lean(nb, np);
          tc_fence_after();
        }
      }
      if (s == 0) {
        if (kStatic) mbar_wait_lean(st.bar_base + 8u * (RB_PFULL + ip), st.par);
        else if (!(st.cont && ip < 2)) mbar_wait_lean(st.pfb[ip], st.pfp[ip]);
        if (kFirst && jlo == 0) {
          umma_f16_new(dcol, mk_desc(a0, hi_a), mk_desc(wl, hi_b), id1);
          if (nblk > 1)
            umma_f16_acc(dcol + 64u, mk_desc(a0, hi_a), mk_desc(wl + kBlkLo, hi_b),
                         nblk == 3 ? id2 : id1);
#pragma unroll
          for (int kk = 1; kk < 4; ++kk)
            mma_acc(dcol, mk_desc(a0 + 2u * kk, hi_a), mk_desc(b0 + 2u * kk, hi_b), idn);
        } else {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_acc(dcol, mk_desc(a0 + 2u * kk, hi_a), mk_desc(b0 + 2u * kk, hi_b), idn);
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          mma_acc(dcol, mk_desc(a0 + 2u * kk, hi_a), mk_desc(b0 + 2u * kk, hi_b), idn);
      }
      if (s == 8) {
        if (kStatic) umma_commit(st.bar_base + 8u * (RB_PEMPTY + ip));
        else if (!(st.next_cont && ip >= kR)) umma_commit(st.peb[ip]);
      }
    }
    umma_commit(st.bar_wempty + 8u * (gs & 1));
    if (s == 8 && st.last_pass) umma_commit(st.bar_accfull);
  }
}

template <int kP>
__device__ __forceinline__ void ring_mma_fast(const UmmaParams& p, uint32_t bar_base,
                                              uint32_t a_base, uint32_t w_base, uint32_t w_slab,
                                              uint32_t tmem_base, int i0, int i1, int lane) {
  constexpr int kR = 4;
  constexpr uint32_t kPlaneLo = (18u * 10u * 128u) >> 4;
  auto bar = [&](int i) { return bar_base + 8u * i; };
  const uint32_t fmtb = p.fmt == 0 ? 1u : 0u;
  const uint32_t hi_a = sdesc_hi_sw128(1280u), hi_b = sdesc_hi_sw128(1024u);
  const uint32_t id1 = make_idesc_f16(64u, fmtb), id2 = make_idesc_f16(128u, fmtb),
                 id3 = make_idesc_f16(192u, fmtb);
  // (kind::f8f6f4: format field 0 = e4m3, i.e. the bits of the fp16 descriptor)
  const uint32_t a_lo0 = sdesc_lo(a_base);
  const int npass = p.npass;
  RingStage st;
  st.w_lo0 = sdesc_lo(w_base);
  st.w_lo1 = sdesc_lo(w_base + w_slab);
  st.bar_wfull = bar(RB_WFULL);
  st.bar_wempty = bar(RB_WEMPTY);
  st.a_lo0 = a_lo0;
  st.bar_base = bar_base;
  st.par = 0;
  st.g = 0;
  const bool is_static = kP == 6 && npass == 2;
  int ab = 0, abph = 0, slot0 = 0;
  uint32_t pf_phase = 0;
  const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0;
  const long long t_all0 = tr ? clock64() : 0;
  for (int i = i0; i < i1; ++i) {
    const RingItem c = ring_decode(p, i, i0, i1);
    st.has_next = i + 1 < i1;
    st.cont = c.cont;
    st.next_cont = c.next_cont;
    st.acc0 = tmem_base + (uint32_t)(ab * kR * 64);
    st.next_acc_par = (uint32_t)((ab == 1 ? abph ^ 1 : abph) ^ 1);
    st.bar_accfull = bar(RB_ACCFULL + ab);
    st.bar_accempty_next = bar(RB_ACCEMPTY + (ab ^ 1));
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
      st.last_pass = pass == npass - 1;
      if (!is_static) {
#pragma unroll
        for (int ip = 0; ip < kR + 2; ++ip) {
          int slot = slot0 + ip;
          if (slot >= kP) slot -= kP;
          st.al[ip] = a_lo0 + (uint32_t)slot * kPlaneLo;
          st.pfb[ip] = bar(RB_PFULL + slot);
          st.peb[ip] = bar(RB_PEMPTY + slot);
          st.pfp[ip] = (pf_phase >> slot) & 1u;
          if (!(c.cont && ip < 2)) pf_phase ^= 1u << slot;
        }
      }
      if (elect_one()) {
        if (i == i0 && pass == 0) {
          mbar_wait_lean(bar(RB_ACCEMPTY + ab), (uint32_t)(abph ^ 1));
          mbar_wait_lean(bar(RB_WFULL + 0), 0u);
          tc_fence_after();
        }
        if (kP == 6 && is_static) {
          if (pass == 0) ring_issue_stage<6, false, true, true>(st, hi_a, hi_b, id1, id2, id3);
          else ring_issue_stage<6, true, false, true>(st, hi_a, hi_b, id1, id2, id3);
        } else {
          if (pass == 0) ring_issue_stage<kP, false, true, false>(st, hi_a, hi_b, id1, id2, id3);
          else ring_issue_stage<kP, true, false, false>(st, hi_a, hi_b, id1, id2, id3);
        }
      }
      __syncwarp();
      st.g += 9;
      st.par ^= 1u;
      slot0 += c.next_cont ? kR : kR + 2;
      while (slot0 >= kP) slot0 -= kP;
    }
    if (++ab == 2) { ab = 0; abph ^= 1; }
  }
  if (tr) {
    p.trace[0] = clock64() - t_all0; p.trace[1] = 0; p.trace[2] = 0; p.trace[3] = 0;
    p.trace[4] = 0; p.trace[5] = i1 - i0;
  }
}

// ------------------------------------------------------------------------------------ kernel
// kR > 0: compile-time planes per item (fully unrolled issue loop); kR == 0: runtime p.R
// tm_a / tm_w: operands of pass 0 (16-bit); tm_a2 / tm_w2: operands of pass 1 (p.npass == 2: the
// e4m3 corr tensors of the fp16c format).
template <int kR, int EPI>
__global__ void __launch_bounds__(ring_threads(EPI), 1)
conv_umma_zring_kernel(const __grid_constant__ CUtensorMap tm_a,
                       const __grid_constant__ CUtensorMap tm_w,
                       const __grid_constant__ CUtensorMap tm_a2,
                       const __grid_constant__ CUtensorMap tm_w2,
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
  const int npass = p.npass;

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
      mbar_init(bar(RB_ACCEMPTY + i), epi_is_v4(EPI) ? 16 : 8);
    }
    for (int i = 0; i < 16; ++i) mbar_init(bar(RB_EPILD + i), 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.npad; i += blockDim.x)
    sbias[i] = (p.ep.bias && i < p.g.cout) ? p.ep.bias[i] : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    if (npass > 1) {
      tma_prefetch_desc(&tm_a2);
      tma_prefetch_desc(&tm_w2);
    }
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
      auto load_slab = [&](int s, int pass, int it) {
        mbar_wait_inl(bar(RB_WEMPTY + ws), wph ^ 1, p.dbg, 2, ws, it * 100 + s);
        if (elect_one()) {
          if ((p.dbg_flags & 4) && (it > 0 || pass > 0 || s >= 2)) {   // timing experiment
            mbar_arrive(bar(RB_WFULL + ws));
          } else {
            mbar_expect_tx(bar(RB_WFULL + ws), p.w_bytes);
            tma_load_3d(w_base + ws * w_slab, pass ? &tm_w2 : &tm_w, bar(RB_WFULL + ws), 0, 0, s);
          }
        }
        __syncwarp();
        if (++ws == WS) { ws = 0; wph ^= 1; }
      };
      for (int i = i0; i < i1; ++i) {
        const RingItem c = ring_decode(p, i, i0, i1);
        const int plane0 = c.b * p.plane_pitch + c.pl0;
        for (int pass = 0; pass < npass; ++pass) {
          load_slab(0, pass, i - i0);
          for (int ip = c.cont ? 2 : 0; ip < c.ri + 2; ++ip) {
            mbar_wait_inl(bar(RB_PEMPTY + head), ((pe_phase >> head) & 1u) ^ 1u, p.dbg, 1, head, i - i0);
            pe_phase ^= 1u << head;
            if (elect_one()) {
              if ((p.dbg_flags & 2) && (i > i0 || pass > 0)) {   // timing experiment
                mbar_arrive(bar(RB_PFULL + head));
              } else {
                mbar_expect_tx(bar(RB_PFULL + head), plane_bytes);
                tma_load_4d(a_base + head * plane_bytes, pass ? &tm_a2 : &tm_a,
                            bar(RB_PFULL + head), 0, c.xb * 8, c.yb * 16, plane0 + ip);
              }
            }
            __syncwarp();
            if (++head == P) head = 0;
          }
          for (int s = 1; s < 9; ++s) load_slab(s, pass, i - i0);
        }
      }
    } else if (warp == 1 && kR == 4 && p.ring_fast) {
      if (P == 6) ring_mma_fast<6>(p, bar_base, a_base, w_base, w_slab, tmem_base, i0, i1, lane);
      else ring_mma_fast<7>(p, bar_base, a_base, w_base, w_slab, tmem_base, i0, i1, lane);
    } else if (warp == 1) {
      // ---------------------------------------------- MMA issuer (warp-uniform, elected issue)
      // generic role: any R / ragged last group / ring depth; one loop iteration per slab
      int ws = 0, wph = 0, ab = 0, abph = 0;
      int slot0 = 0;             // ring slot of input plane ip = 0 of the current stage
      uint32_t pf_phase = 0;     // bit per plane slot: parity to wait for on its FULL barrier
      const uint32_t fmtb = p.fmt == 0 ? 1u : 0u;
      const uint32_t hi_a = sdesc_hi_sw128((uint32_t)p.XB * 128u);
      const uint32_t hi_b = sdesc_hi_sw128(1024u);
      const uint32_t id1 = make_idesc_f16((uint32_t)npad, fmtb);
      const uint32_t id2 = make_idesc_f16((uint32_t)(2 * npad), fmtb);
      const uint32_t id3 = make_idesc_f16((uint32_t)(3 * npad), fmtb);
      const uint32_t blk_lo = ((uint32_t)npad * 128u) >> 4;   // one weight block in desc units
      const uint32_t xb128 = (uint32_t)p.XB * 128u;
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
        for (int pass = 0; pass < npass; ++pass) {
          const bool f8 = pass == 1, last_pass = pass == npass - 1;
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
              if (s == 0 && jlo == 0 && pass == 0) {
                umma_f16_new(dcol, mk_desc(al, hi_a), mk_desc(wl, hi_b), id1);
                if (nblk > 1)
                  umma_f16_acc(dcol + npad, mk_desc(al, hi_a), mk_desc(wl + blk_lo, hi_b),
                               nblk == 3 ? id2 : id1);
#pragma unroll
                for (int kk = 1; kk < 4; ++kk)
                  umma_f16_acc(dcol, mk_desc(al + 2u * kk, hi_a), mk_desc(bl + 2u * kk, hi_b), idn);
              } else if (f8) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f8_acc(dcol, mk_desc(al + 2u * kk, hi_a), mk_desc(bl + 2u * kk, hi_b), idn);
              } else {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16_acc(dcol, mk_desc(al + 2u * kk, hi_a), mk_desc(bl + 2u * kk, hi_b), idn);
              }
              if (s == 8 && !(c.next_cont && ip >= ri)) umma_commit(bar(RB_PEMPTY + slot));
            };
            if (s == 0) {
              // first slab: wait plane by plane so that the MMAs start as soon as the first
              // planes of the stage have landed
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
                for (int ip = 0; ip < np; ++ip) issue_plane(ip, c.ri);
                umma_commit(bar(RB_WEMPTY + ws));
                if (s == 8 && last_pass) umma_commit(bar(RB_ACCFULL + ab));
              }
              __syncwarp();
            }
            if (++ws == WS) { ws = 0; wph ^= 1; }
          }
          // the next stage starts at the carried planes (continuing) or after all of this one's
          slot0 += c.next_cont ? c.ri : np;
          while (slot0 >= P) slot0 -= P;
        }
        if (++ab == 2) { ab = 0; abph ^= 1; }
      }
      if (tr) {
        p.trace[0] = clock64() - t_all0; p.trace[1] = t_acc; p.trace[2] = t_w; p.trace[3] = t_a;
        p.trace[4] = 0; p.trace[5] = i1 - i0;
      }
    }
  } else {
    // ------------------------------------------------------------------------------ epilogue
    if (epi_is_v4(EPI)) reg_inc<104>(); else reg_inc<216>();
    constexpr int kWgs = epi_is_v4(EPI) ? 4 : 2;   // epilogue warpgroups
    const int wg = (warp - 4) >> 2, q = warp & 3;
    int ab = 0, abph = 0;
    EpiTma et;
    et.res[0] = &em.res_hi; et.res[1] = &em.res_lo;
    et.out[0] = &em.y_hi; et.out[1] = &em.y_lo;
    et.row[0] = p.epi_row_tma ? &em.row_hi : nullptr;
    et.row[1] = p.epi_row_tma ? &em.row_lo : nullptr;
    et.stage_s0 = bar_base + 2048u + (uint32_t)(warp - 4) * 2048u;
    et.stage0 = smem_raw + (et.stage_s0 - smem_u32(smem_raw));
    et.bar0 = bar(RB_EPILD + (warp - 4));
    et.phase0 = 0;
    et.dbg = p.dbg_flags;
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 128;
    long long t_wait = 0, t_work = 0;
    for (int i = i0; i < i1; ++i) {
      const RingItem c = ring_decode(p, i, i0, i1);
      long long c0 = tr ? clock64() : 0;
      bool prefetched = false;
      if (epi_is_v4(EPI) && wg < c.ri && !(p.dbg_flags & 8)) {
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
      if (epi_is_v4(EPI)) {
        if (!(p.dbg_flags & 8)) {
          const ConvGeom& g = p.g;
          TileGeom tg;
          tg.sy = (long long)(g.fd[2] + 2) * 128;
          tg.sz = (long long)(g.fd[1] + 2) * tg.sy;
          tg.y0 = c.yb * 16 + q * 4;
          tg.x0 = c.xb * 8;
          for (int r = wg; r < c.ri; r += kWgs) {
            const int z = c.pl0 + r;
            const int plane_coord = c.b * (g.fd[0] + 2) + z + 1;
            tg.base = (((long long)plane_coord * (g.fd[1] + 2) + tg.y0 + 1) * (g.fd[2] + 2) +
                       (long long)tg.x0 * (EPI == EPI_V4R ? g.rep[2] : 1) + 1) * 128;
            const int mzp = z == 1 ? -2 : (z == g.fd[0] - 2 ? 2 : 0);
            const uint32_t ta = tmem_base + (uint32_t)(ab * R * npad + npad * (R - 1 - r)) +
                                ((uint32_t)(q * 32) << 16);
            if (EPI == EPI_V4R)
              ring_epilogue_warp_v4<false, true>(g, p.ep, sbias, ta, tg, plane_coord, mzp, et, lane,
                                                 prefetched);
            else if (p.ep.res_hi)
              ring_epilogue_warp_v4<true, false>(g, p.ep, sbias, ta, tg, plane_coord, mzp, et, lane,
                                                 prefetched);
            else
              ring_epilogue_warp_v4<false, false>(g, p.ep, sbias, ta, tg, plane_coord, mzp, et, lane,
                                                  prefetched);
          }
        }
      } else if (!(p.dbg_flags & 8)) {
        for (int r = wg; r < c.ri; r += 2)
          ring_epilogue_tile<epi_is_v4(EPI) ? EPI_PLAIN : EPI>(
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
    if (epi_is_v4(EPI) && lane == 0) tma_store_wait_read();
    if (tr) { p.trace[8] = t_wait; p.trace[9] = t_work; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int kR, int EPI>
static int launch_zring_t(const UmmaParams& p, const CUtensorMap* maps, const EpiMaps& em,
                          int ctas, uint32_t smem, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    S3_CUDA(cudaFuncSetAttribute(conv_umma_zring_kernel<kR, EPI>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr = true;
  }
  conv_umma_zring_kernel<kR, EPI><<<ctas, ring_threads(EPI), smem, st>>>(maps[0], maps[1], maps[2],
                                                                           maps[3], em, p);
  S3_CUDA(cudaGetLastError());
  return S3_OK;
}

int launch_umma_zring(const UmmaParams& p, const CUtensorMap* maps /* a, w, a2, w2 */,
                      const CUtensorMap* epi_maps, int epi, int ctas, uint32_t smem,
                      cudaStream_t st) {
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  if (epi_maps) {
    em.res_hi = epi_maps[0]; em.res_lo = epi_maps[1]; em.y_hi = epi_maps[2]; em.y_lo = epi_maps[3];
    em.row_hi = epi_maps[4]; em.row_lo = epi_maps[5];
  }
  if (p.epi_v4 && p.R == 4 && p.g.rep[2] > 1)
    return launch_zring_t<4, EPI_V4R>(p, maps, em, ctas, smem, st);
  if (p.epi_v4 && p.R == 4) return launch_zring_t<4, EPI_V4>(p, maps, em, ctas, smem, st);
  if (epi == EPI_PLAIN && p.R == 4) return launch_zring_t<4, EPI_PLAIN>(p, maps, em, ctas, smem, st);
  if (epi == EPI_PLAIN) return launch_zring_t<0, EPI_PLAIN>(p, maps, em, ctas, smem, st);
  return launch_zring_t<0, EPI_GENERIC>(p, maps, em, ctas, smem, st);
}

}  // namespace s3
