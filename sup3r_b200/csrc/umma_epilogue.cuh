// Lean epilogue for the tcgen05 convolutions: one thread owns one accumulator row (= one conv
// output voxel, `cout` fp32 columns in TMEM).  Everything per-row lives in registers (the CTA
// keeps ~215 KB of shared memory, which leaves almost no L1, so local-memory or repeated global
// loads would each cost an L2 round trip): destinations are a base offset plus at most one
// mirror delta per dim; the bias sits in shared memory; the residual row is prefetched.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace s3 {

// epilogue specialisations (one per kernel instantiation keeps the SASS small: the generic
// scatter is ~10x the code of the fast paths and would thrash the instruction cache)
enum { EPI_PLAIN = 0, EPI_D2S = 1, EPI_GENERIC = 2, EPI_V2 = 3, EPI_V3 = 4, EPI_D2S16 = 5, EPI_V4 = 6,
       EPI_V4R = 7 };   // V4R: V4 with nearest repeat along x (one TMA store per replica)
__host__ __device__ constexpr bool epi_is_v4(int epi) { return epi == EPI_V4 || epi == EPI_V4R; }  // zring: V2 LSU-coalescing, V3 TMA tile I/O

struct RowPlan {
  bool valid, slow;
  int b, z, y, x;
  size_t conv_vox;
  size_t base32;         // f32 element offset of the rx = 0 copy
  size_t base16;         // 16-bit element offset of the rx = 0 copy (interior position)
  long long mz, my;      // halo mirror deltas along z / y in elements (0 = none)
};

// plain mapping (no depth_to_space / depth_to_time); nearest repeat along x only
__device__ __forceinline__ void plan_plain(const ConvGeom& g, const Epilogue& ep, RowPlan& rp) {
  rp.slow = (g.rep[0] != 1 || g.rep[1] != 1 || g.rep[2] > 3 || g.r != 1 || g.m != 1);
  rp.base32 = rp.base16 = 0;
  rp.mz = rp.my = 0;
  if (!rp.valid || rp.slow) return;
  const int FZ = g.fd[0], FY = g.fd[1], FX = g.fd[2];
  const int pz = (g.ndim == 3) ? 1 : 0;
  const long long PY = FY + 2, PX = FX + 2;
  const int ox0 = rp.x * g.rep[2];
  rp.base32 = ((((size_t)rp.b * FZ + rp.z) * FY + rp.y) * FX + ox0) * g.cstride + g.coff;
  rp.base16 = ((((size_t)rp.b * (FZ + 2 * pz) + rp.z + pz) * PY + rp.y + 1) * PX + ox0 + 1) *
                  g.cstride + g.coff;
  const long long sy = PX * g.cstride, sz = PY * sy;
  if (pz) {
    const bool lo = rp.z == 1, hi = rp.z == FZ - 2;
    if (lo && hi) rp.slow = true;
    rp.mz = lo ? -2 * sz : (hi ? 2 * sz : 0);
  }
  {
    const bool lo = rp.y == 1, hi = rp.y == FY - 2;
    if (lo && hi) rp.slow = true;
    rp.my = lo ? -2 * sy : (hi ? 2 * sy : 0);
  }
  if (FX == 3) rp.slow = true;
}

__device__ __forceinline__ uint32_t pack2(float a, float b, int fmt) {
  if (fmt == 0) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  if (fmt == kFmtFp16c) return f16x2_sat(a, b);
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// fp16c: the "lo" half of 16 consecutive channels c0 .. c0 + 15 (c0 % 16 == 0) of one voxel:
// l8 = e4m3((v - fp16(v)) 2^11), a8 = e4m3(v); destination = byte offsets corr_byte(0 / 1, c0)
// inside the voxel's 128-byte corr row
__device__ __forceinline__ void corr16(const float* v, uint4& l8, uint4& a8) {
  float e[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) e[j] = (v[j] - from16(to16(v[j], kFmtFp16c), kFmtFp16c)) * kCorrScale;
  l8.x = e4m3x4(e[0], e[1], e[2], e[3]);     l8.y = e4m3x4(e[4], e[5], e[6], e[7]);
  l8.z = e4m3x4(e[8], e[9], e[10], e[11]);   l8.w = e4m3x4(e[12], e[13], e[14], e[15]);
  a8.x = e4m3x4(v[0], v[1], v[2], v[3]);     a8.y = e4m3x4(v[4], v[5], v[6], v[7]);
  a8.z = e4m3x4(v[8], v[9], v[10], v[11]);   a8.w = e4m3x4(v[12], v[13], v[14], v[15]);
}
// v[0..15] += lo8 2^-11 of 16 channels (one uint4 of a corr row)
__device__ __forceinline__ void corr_add16(float* v, const uint4& u) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = e4m3x2_to_f32(w[j]), b = e4m3x2_to_f32(w[j] >> 16);
    v[4 * j] += a.x * kCorrInv;     v[4 * j + 1] += a.y * kCorrInv;
    v[4 * j + 2] += b.x * kCorrInv; v[4 * j + 3] += b.y * kCorrInv;
  }
}

__device__ __forceinline__ void store16x2(uint16_t* base, long long off, const uint4& a,
                                          const uint4& b) {
  uint4* d = reinterpret_cast<uint4*>(base + off);
  d[0] = a;
  d[1] = b;
}

// Process all column chunks of one accumulator row.  t_addr: TMEM address of column 0 of this
// warp's lane quarter.  sbias: bias staged in shared memory (zeros when absent).
// All 32 lanes must call (LDTM is warp-collective).
// Hot path: plain mapping, cout == 64.  All four TMEM chunks and the residual row are in
// flight before anything is consumed (one TMEM wait, one DRAM/L2 round trip per tile).
__device__ __forceinline__ void epilogue_row_plain64(const ConvGeom& g, const Epilogue& ep,
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
  const int rep = g.rep[2];
  const int fmt = ep.fmt;
  uint16_t* yh = reinterpret_cast<uint16_t*>(ep.y_hi);
  uint16_t* yl = reinterpret_cast<uint16_t*>(ep.y_lo);
  const float sc = ep.acc_scale;
  const bool corr = fmt == kFmtFp16c;
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    float v[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 bv = *reinterpret_cast<const float4*>(sbias + cc * 16 + 4 * q);
      v[4 * q] = fmaf(__uint_as_float(raw[cc * 16 + 4 * q]), sc, bv.x);
      v[4 * q + 1] = fmaf(__uint_as_float(raw[cc * 16 + 4 * q + 1]), sc, bv.y);
      v[4 * q + 2] = fmaf(__uint_as_float(raw[cc * 16 + 4 * q + 2]), sc, bv.z);
      v[4 * q + 3] = fmaf(__uint_as_float(raw[cc * 16 + 4 * q + 3]), sc, bv.w);
    }
    if (has_res && g.res_pre) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 r4 = rpre[cc * 4 + q];
        v[4 * q] += r4.x; v[4 * q + 1] += r4.y; v[4 * q + 2] += r4.z; v[4 * q + 3] += r4.w;
      }
    }
    if (g.act == S3_ACT_LEAKY) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = v[j] >= 0.f ? v[j] : g.alpha * v[j];
    } else if (g.act == S3_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (g.act != S3_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = apply_act_slow(v[j], g.act, g.alpha);
    }
    if (has_res && !g.res_pre) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 r4 = rpre[cc * 4 + q];
        v[4 * q] += r4.x; v[4 * q + 1] += r4.y; v[4 * q + 2] += r4.z; v[4 * q + 3] += r4.w;
      }
    }
    uint4 h0, h1, l0, l1;
    if (yh) {
      h0.x = pack2(v[0], v[1], fmt);   h0.y = pack2(v[2], v[3], fmt);
      h0.z = pack2(v[4], v[5], fmt);   h0.w = pack2(v[6], v[7], fmt);
      h1.x = pack2(v[8], v[9], fmt);   h1.y = pack2(v[10], v[11], fmt);
      h1.z = pack2(v[12], v[13], fmt); h1.w = pack2(v[14], v[15], fmt);
      if (yl && corr) {
        corr16(v, l0, l1);   // l0 = lo8 x 16, l1 = a8 x 16
      } else if (yl) {
        float e[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) e[j] = v[j] - from16(to16(v[j], fmt), fmt);
        l0.x = pack2(e[0], e[1], fmt);   l0.y = pack2(e[2], e[3], fmt);
        l0.z = pack2(e[4], e[5], fmt);   l0.w = pack2(e[6], e[7], fmt);
        l1.x = pack2(e[8], e[9], fmt);   l1.y = pack2(e[10], e[11], fmt);
        l1.z = pack2(e[12], e[13], fmt); l1.w = pack2(e[14], e[15], fmt);
      }
    }
    // lo destination of this 16-channel chunk: 16-bit lo = same offsets as hi; corr rows =
    // 16 B of lo8 at byte corr_byte(0, 16 cc) and 16 B of a8 at corr_byte(1, 16 cc) of the voxel
    auto store_lo = [&](long long o) {   // o = element offset of this chunk in the hi tensor
      if (corr) {
        uint8_t* row = reinterpret_cast<uint8_t*>(yl) + (o - cc * 16) * 2;
        *reinterpret_cast<uint4*>(row + corr_byte(0, cc * 16)) = l0;
        *reinterpret_cast<uint4*>(row + corr_byte(1, cc * 16)) = l1;
      } else {
        store16x2(yl, o, l0, l1);
      }
    };
#pragma unroll 1
    for (int rx = 0; rx < rep; ++rx) {
      if (ep.y) {
        float4* dst = reinterpret_cast<float4*>(ep.y + rp.base32 + (size_t)rx * 64 + cc * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      if (yh) {
        const int ox = rp.x * rep + rx;
        const long long mx = ox == 1 ? -128LL : (ox == g.fd[2] - 2 ? 128LL : 0LL);
        const long long o0 = (long long)rp.base16 + (long long)rx * 64 + cc * 16;
        store16x2(yh, o0, h0, h1);
        if (yl) store_lo(o0);
        if ((rp.mz | rp.my | mx) != 0) {
#pragma unroll 1
          for (int combo = 1; combo < 8; ++combo) {
            const bool a = combo & 4, bq = combo & 2, cq = combo & 1;
            if ((a && rp.mz == 0) || (bq && rp.my == 0) || (cq && mx == 0)) continue;
            const long long o = o0 + (a ? rp.mz : 0) + (bq ? rp.my : 0) + (cq ? mx : 0);
            store16x2(yh, o, h0, h1);
            if (yl) store_lo(o);
          }
        }
      }
    }
  }
}

//   EPI_PLAIN  : host guarantees plain mapping, cout % 16 == 0, aligned strides, extents >= 4
//                (no "slow" rows), nearest repeat along x only (<= 3 copies)
//   EPI_D2S    : host guarantees depth_to_space / depth_to_time with cmap in {4, 8, 16},
//                cout % cmap == 0, f32 destination only, no repeat
//   EPI_GENERIC: anything
template <int EPI>
__device__ __forceinline__ void epilogue_row(const ConvGeom& g, const Epilogue& ep,
                                             const float* sbias, uint32_t t_addr, RowPlan& rp) {
  if (EPI == EPI_PLAIN && g.cout == 64 && g.cstride == 64 && g.coff == 0 && !ep.post_scale) {
    epilogue_row_plain64(g, ep, sbias, t_addr, rp);
    return;
  }
  constexpr bool kPrefetchResidual = false;
  const bool mapped = EPI == EPI_PLAIN ? false : (EPI == EPI_D2S ? true : (g.r != 1 || g.m != 1));
  const bool vec32 = EPI != EPI_GENERIC ? true : ((g.cstride % 4 == 0) && (g.coff % 4 == 0));
  const bool vec16 = EPI != EPI_GENERIC ? true : ((g.cstride % 8 == 0) && (g.coff % 8 == 0));
  const bool fast_plain = EPI == EPI_PLAIN ? true
                                           : (EPI == EPI_D2S ? false
                                                             : (!mapped && !rp.slow && vec32 && vec16));

  // residual row prefetch (cout <= 64): 16 x LDG.128 in flight before the first TMEM load
  float4 rpre[kPrefetchResidual ? 16 : 1];
  if (kPrefetchResidual) {
    if (rp.valid && ep.residual && g.cout <= 64) {
      const float4* rr = reinterpret_cast<const float4*>(ep.residual + rp.conv_vox * g.cout);
#pragma unroll
      for (int q = 0; q < 16; ++q)
        if (q * 4 < g.cout) rpre[q] = __ldg(rr + q);
    }
  }

#pragma unroll 1
  for (int c0 = 0; c0 < g.cout; c0 += 16) {
    uint32_t raw[16];
    tmem_ld16(t_addr + c0, raw);
    tmem_ld_wait();
    if (!rp.valid) continue;
    const int len = min(16, g.cout - c0);
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]) * ep.acc_scale;
    if (EPI == EPI_PLAIN || len == 16) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 bv = *reinterpret_cast<const float4*>(sbias + c0 + 4 * q);
        v[4 * q] += bv.x; v[4 * q + 1] += bv.y; v[4 * q + 2] += bv.z; v[4 * q + 3] += bv.w;
      }
      if (g.act == S3_ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = v[j] >= 0.f ? v[j] : g.alpha * v[j];
      } else if (g.act == S3_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (g.act != S3_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = apply_act_slow(v[j], g.act, g.alpha);
      }
      if (ep.residual) {
        if (kPrefetchResidual && g.cout <= 64) {
          // select the prefetched quad for this chunk with static indices
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            if (c0 == cc * 16) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 r4 = rpre[cc * 4 + q];
                v[4 * q] += r4.x; v[4 * q + 1] += r4.y; v[4 * q + 2] += r4.z; v[4 * q + 3] += r4.w;
              }
            }
        } else {
          const float4* rr =
              reinterpret_cast<const float4*>(ep.residual + rp.conv_vox * g.cout + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 r4 = __ldg(rr + q);
            v[4 * q] += r4.x; v[4 * q + 1] += r4.y; v[4 * q + 2] += r4.z; v[4 * q + 3] += r4.w;
          }
        }
      }
      if (ep.post_scale) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          v[j] = v[j] * ep.post_scale[c0 + j] + (ep.post_shift ? ep.post_shift[c0 + j] : 0.f);
      }
    } else if (EPI != EPI_PLAIN) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        v[j] = j < len ? finish(g, ep, v[j], c0 + j, rp.conv_vox) : 0.f;
    }

    if (fast_plain && (EPI == EPI_PLAIN || len == 16)) {
      // ---- fast path: the 16-channel run goes to every destination row
      uint4 h0, h1, l0, l1;
      if (ep.y_hi) {
        h0.x = pack2(v[0], v[1], ep.fmt);   h0.y = pack2(v[2], v[3], ep.fmt);
        h0.z = pack2(v[4], v[5], ep.fmt);   h0.w = pack2(v[6], v[7], ep.fmt);
        h1.x = pack2(v[8], v[9], ep.fmt);   h1.y = pack2(v[10], v[11], ep.fmt);
        h1.z = pack2(v[12], v[13], ep.fmt); h1.w = pack2(v[14], v[15], ep.fmt);
        if (ep.y_lo) {
          float e[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) e[j] = v[j] - from16(to16(v[j], ep.fmt), ep.fmt);
          l0.x = pack2(e[0], e[1], ep.fmt);   l0.y = pack2(e[2], e[3], ep.fmt);
          l0.z = pack2(e[4], e[5], ep.fmt);   l0.w = pack2(e[6], e[7], ep.fmt);
          l1.x = pack2(e[8], e[9], ep.fmt);   l1.y = pack2(e[10], e[11], ep.fmt);
          l1.z = pack2(e[12], e[13], ep.fmt); l1.w = pack2(e[14], e[15], ep.fmt);
        }
      }
#pragma unroll
      for (int rx = 0; rx < 3; ++rx) {
        if (rx >= g.rep[2]) break;
        if (ep.y) {
          float4* dst = reinterpret_cast<float4*>(ep.y + rp.base32 + (size_t)rx * g.cstride + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        if (ep.y_hi) {
          const int ox = rp.x * g.rep[2] + rx;
          const long long mx = ox == 1 ? -2LL * g.cstride
                                       : (ox == g.fd[2] - 2 ? 2LL * g.cstride : 0LL);
          const long long o0 = (long long)rp.base16 + (long long)rx * g.cstride + c0;
          uint16_t* yh = reinterpret_cast<uint16_t*>(ep.y_hi);
          uint16_t* yl = reinterpret_cast<uint16_t*>(ep.y_lo);
          store16x2(yh, o0, h0, h1);
          if (ep.y_lo) store16x2(yl, o0, l0, l1);
          if ((rp.mz | rp.my | mx) != 0) {
#pragma unroll 1
            for (int combo = 1; combo < 8; ++combo) {
              const bool a = combo & 4, bq = combo & 2, cq = combo & 1;
              if ((a && rp.mz == 0) || (bq && rp.my == 0) || (cq && mx == 0)) continue;
              const long long o = o0 + (a ? rp.mz : 0) + (bq ? rp.my : 0) + (cq ? mx : 0);
              store16x2(yh, o, h0, h1);
              if (ep.y_lo) store16x2(yl, o, l0, l1);
            }
          }
        }
      }
    } else if (EPI == EPI_D2S ||
               (EPI == EPI_GENERIC && mapped && g.rep[0] * g.rep[1] * g.rep[2] == 1 &&
                (ep.y || ep.y_hi) && (g.cmap == 4 || g.cmap == 8 || g.cmap % 16 == 0) &&
                len % (g.cmap < 16 ? g.cmap : 16) == 0 && g.cbase % 16 == 0)) {
      // ---- depth_to_space / depth_to_time fast path: runs of cmap channels; f32 (ep.y) or an
      // unpadded 16-bit mapped tensor (ep.y_hi; 8-channel runs of 8 consecutive x voxels form
      // full 128-byte lines per warp instruction)
      // runs of min(cmap, 16) channels: wider mapped voxels (cmap 32, 64, ...) take one
      // 16-channel piece per step at channel offset d.c
      const int rs = g.cmap < 16 ? g.cmap : 16;
      const int nrun = len / rs;
      if (g.cmap == 8 && ep.y_hi && !ep.y) {
        // hot case (5x spatial head -> bf16 high-resolution tensor): static register indices
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (s < nrun) {
            const Dest d = map_dest(g, rp.z, rp.y, rp.x, c0 + s * 8);
            const size_t doff = ((((size_t)rp.b * g.fd[0] + d.z) * g.fd[1] + d.y) * g.fd[2] + d.x) *
                                    g.cstride + g.coff;
            uint4 u;
            u.x = pack2(v[8 * s], v[8 * s + 1], ep.fmt);
            u.y = pack2(v[8 * s + 2], v[8 * s + 3], ep.fmt);
            u.z = pack2(v[8 * s + 4], v[8 * s + 5], ep.fmt);
            u.w = pack2(v[8 * s + 6], v[8 * s + 7], ep.fmt);
            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(ep.y_hi) + doff) = u;
          }
        }
        continue;
      }
      for (int s = 0; s < nrun; ++s) {
        const Dest d = map_dest(g, rp.z, rp.y, rp.x, c0 + s * rs);
        const size_t doff = ((((size_t)rp.b * g.fd[0] + d.z) * g.fd[1] + d.y) * g.fd[2] + d.x) *
                                g.cstride + g.coff + d.c;
        if (ep.y_hi) {
          uint16_t* d16 = reinterpret_cast<uint16_t*>(ep.y_hi) + doff;
          const float* vs = v + s * rs;   // (s * rs is a multiple of 4: register select below)
          if (g.cmap == 8) {
            const int o = s * 8;
            uint4 u;
            u.x = pack2(o == 0 ? v[0] : v[8], o == 0 ? v[1] : v[9], ep.fmt);
            u.y = pack2(o == 0 ? v[2] : v[10], o == 0 ? v[3] : v[11], ep.fmt);
            u.z = pack2(o == 0 ? v[4] : v[12], o == 0 ? v[5] : v[13], ep.fmt);
            u.w = pack2(o == 0 ? v[6] : v[14], o == 0 ? v[7] : v[15], ep.fmt);
            *reinterpret_cast<uint4*>(d16) = u;
          } else {
            for (int k = 0; k < rs; ++k) d16[k] = to16(v[(s * rs + k) & 15], ep.fmt);
          }
          (void)vs;
          if (!ep.y) continue;
        }
        float* dst = ep.y + doff;
        if (g.cmap == 4) {
          float4 o;
          if (s == 0) o = make_float4(v[0], v[1], v[2], v[3]);
          else if (s == 1) o = make_float4(v[4], v[5], v[6], v[7]);
          else if (s == 2) o = make_float4(v[8], v[9], v[10], v[11]);
          else o = make_float4(v[12], v[13], v[14], v[15]);
          if (vec32) *reinterpret_cast<float4*>(dst) = o;
          else { dst[0] = o.x; dst[1] = o.y; dst[2] = o.z; dst[3] = o.w; }
        } else if (g.cmap == 8) {
          float4 o0, o1;
          if (s == 0) { o0 = make_float4(v[0], v[1], v[2], v[3]); o1 = make_float4(v[4], v[5], v[6], v[7]); }
          else { o0 = make_float4(v[8], v[9], v[10], v[11]); o1 = make_float4(v[12], v[13], v[14], v[15]); }
          if (vec32) {
            reinterpret_cast<float4*>(dst)[0] = o0;
            reinterpret_cast<float4*>(dst)[1] = o1;
          } else {
            dst[0] = o0.x; dst[1] = o0.y; dst[2] = o0.z; dst[3] = o0.w;
            dst[4] = o1.x; dst[5] = o1.y; dst[6] = o1.z; dst[7] = o1.w;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (vec32)
              reinterpret_cast<float4*>(dst)[q] =
                  make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            else { dst[4 * q] = v[4 * q]; dst[4 * q + 1] = v[4 * q + 1];
                   dst[4 * q + 2] = v[4 * q + 2]; dst[4 * q + 3] = v[4 * q + 3]; }
          }
        }
      }
    } else if (EPI == EPI_GENERIC) {
      // ---- generic scatter (rare shapes): element runs through store_run
      int j = 0;
      while (j < len) {
        const Dest d = map_dest(g, rp.z, rp.y, rp.x, c0 + j);
        const int run = mapped ? min(g.cmap - d.c, len - j) : len - j;
        float seg[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) seg[k] = v[(j + k) & 15];
        store_run<16>(g, ep, rp.b, d, seg, run);
        j += run;
      }
    }
  }
}

}  // namespace s3
