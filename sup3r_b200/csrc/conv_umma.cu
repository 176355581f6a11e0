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

namespace s3 {

struct UmmaParams {
  ConvGeom g;
  Epilogue ep;
  int kz, ntaps, npad, split, fmt;
  int R, TS, XB, YB, ZB, WS, AS, acc_bufs, bo_mode;
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

__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tm_a_hi,
                 const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_w_hi,
                 const __grid_constant__ CUtensorMap tm_w_lo, const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A stages (hi, lo)] [W stages (hi, lo)] [barriers]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int halves = p.split ? 2 : 1;
  const uint32_t a_stage_bytes = p.box_stride * halves;
  const uint32_t a_tx_bytes = p.box_bytes * halves;
  const uint32_t w_slab = (p.w_bytes + 1023u) & ~1023u;
  const uint32_t w_stage_bytes = w_slab * halves;
  const uint32_t a_base = smem_base;
  const uint32_t w_base = a_base + a_stage_bytes * p.AS;
  const uint32_t bar_base = w_base + w_stage_bytes * p.WS;
  // barrier slots (8 B each)
  auto bar = [&](int i) { return bar_base + 8u * i; };
  const int B_AFULL = 0, B_AEMPTY = 2, B_WFULL = 4, B_WEMPTY = 4 + kMaxWS,
            B_ACCFULL = 4 + 2 * kMaxWS, B_ACCEMPTY = 6 + 2 * kMaxWS, B_TMEMPTR = 8 + 2 * kMaxWS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_AFULL + i), 1);
      mbar_init(bar(B_AEMPTY + i), 1);
      mbar_init(bar(B_ACCFULL + i), 1);
      mbar_init(bar(B_ACCEMPTY + i), 4);
    }
    for (int i = 0; i < kMaxWS; ++i) {
      mbar_init(bar(B_WFULL + i), 1);
      mbar_init(bar(B_WEMPTY + i), 1);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_w_hi);
    if (p.split) {
      tma_prefetch_desc(&tm_a_lo);
      tma_prefetch_desc(&tm_w_lo);
    }
  }
  if (warp == 1) {
    tmem_alloc(bar(B_TMEMPTR), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bar(B_TMEMPTR)));

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int as = 0, aph = 0, ws = 0, wph = 0, it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const ItemCoord c = decode_item(p, item);
        mbar_wait(bar(B_AEMPTY + as), aph ^ 1, p.dbg, 1, as, it);
        mbar_expect_tx(bar(B_AFULL + as), a_tx_bytes);
        const int plane = c.b * p.plane_pitch + c.pl0;
        tma_load_4d(a_base + as * a_stage_bytes, &tm_a_hi, bar(B_AFULL + as), 0, c.xb * 8, c.y0,
                    plane);
        if (p.split)
          tma_load_4d(a_base + as * a_stage_bytes + p.box_stride, &tm_a_lo, bar(B_AFULL + as), 0,
                      c.xb * 8, c.y0, plane);
        if (++as == p.AS) { as = 0; aph ^= 1; }
        for (int tap = 0; tap < p.ntaps; ++tap) {
          mbar_wait(bar(B_WEMPTY + ws), wph ^ 1, p.dbg, 2, ws, it * 100 + tap);
          mbar_expect_tx(bar(B_WFULL + ws), p.w_bytes * halves);
          tma_load_3d(w_base + ws * w_stage_bytes, &tm_w_hi, bar(B_WFULL + ws), 0, 0, tap);
          if (p.split)
            tma_load_3d(w_base + ws * w_stage_bytes + w_slab, &tm_w_lo, bar(B_WFULL + ws), 0, 0,
                        tap);
          if (++ws == p.WS) { ws = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================= MMA issuer
    if (lane == 0) {
      int as = 0, aph = 0, ws = 0, wph = 0, ab = 0, abph = 0, it = 0;
      const uint32_t sbo_a = (uint32_t)p.XB * 128u;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const ItemCoord c = decode_item(p, item);
        mbar_wait(bar(B_ACCEMPTY + ab), abph ^ 1, p.dbg, 3, ab, it);
        mbar_wait(bar(B_AFULL + as), aph, p.dbg, 4, as, it);
        tc_fence_after();
        const uint32_t a_hi = a_base + as * a_stage_bytes;
        const uint32_t a_lo = a_hi + p.box_stride;
        const uint32_t d_base = tmem_base + (uint32_t)(ab * p.R * p.npad);
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int dx = tap % 3, dy = (tap / 3) % 3, dz = tap / 9;
          mbar_wait(bar(B_WFULL + ws), wph, p.dbg, 5, ws, it * 100 + tap);
          tc_fence_after();
          const uint32_t w_hi = w_base + ws * w_stage_bytes;
          const uint32_t w_lo = w_hi + w_slab;
          for (int r = 0; r < p.R; ++r) {
            const uint32_t row = (uint32_t)(c.row0 + r * p.TS + dz * p.YB + dy);
            const uint32_t a_off = (row * p.XB + dx) * 128u;
            const uint32_t d_addr = d_base + (uint32_t)(r * p.npad);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t aa = a_hi + a_off + kk * 32u;
              const uint32_t bo = p.bo_mode ? ((aa >> 7) & 7u) : 0u;
              const uint64_t da = make_sdesc_sw128(aa, sbo_a, bo);
              const uint64_t db = make_sdesc_sw128(w_hi + kk * 32u, 1024u, 0);
              umma_f16(d_addr, da, db, p.idesc, (tap | kk) != 0 ? 1u : 0u);
              if (p.split) {
                const uint32_t al = a_lo + a_off + kk * 32u;
                const uint64_t dal = make_sdesc_sw128(al, sbo_a, p.bo_mode ? ((al >> 7) & 7u) : 0u);
                const uint64_t dbl = make_sdesc_sw128(w_lo + kk * 32u, 1024u, 0);
                umma_f16(d_addr, dal, db, p.idesc, 1u);
                umma_f16(d_addr, da, dbl, p.idesc, 1u);
              }
            }
          }
          umma_commit(bar(B_WEMPTY + ws));
          if (++ws == p.WS) { ws = 0; wph ^= 1; }
        }
        umma_commit(bar(B_AEMPTY + as));
        umma_commit(bar(B_ACCFULL + ab));
        if (++as == p.AS) { as = 0; aph ^= 1; }
        if (++ab == p.acc_bufs) { ab = 0; abph ^= 1; }
      }
    }
  } else {
    // ========================================================================= epilogue
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;        // accumulator row
    const int grp = m >> 3, xl = m & 7;
    const ConvGeom& g = p.g;
    int ab = 0, abph = 0, it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const ItemCoord c = decode_item(p, item);
      mbar_wait(bar(B_ACCFULL + ab), abph, p.dbg, 6, ab, it);
      tc_fence_after();
      for (int r = 0; r < p.R; ++r) {
        const int fr = c.row0 + r * p.TS + grp;
        const int zq = fr / p.YB, yq = fr - zq * p.YB;
        const int plane = c.pl0 + zq;
        const int y = c.y0 + yq, x = c.xb * 8 + xl;
        const bool valid = yq <= p.YB - 3 && y < g.in[1] && x < g.in[2] && plane < p.planes;
        int b, z;
        if (g.ndim == 3) { b = c.b; z = plane; } else { b = plane; z = 0; }
        const size_t conv_vox = (((size_t)b * g.in[0] + z) * g.in[1] + y) * g.in[2] + x;
        const uint32_t t_addr =
            tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((ab * p.R + r) * p.npad);
        for (int c0 = 0; c0 < g.cout; c0 += 16) {
          uint32_t raw[16];
          tmem_ld16(t_addr + c0, raw);
          tmem_ld_wait();
          if (!valid) continue;
          float v[16];
          const int len = min(16, g.cout - c0);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = j < len ? finish(g, p.ep, __uint_as_float(raw[j]), c0 + j, conv_vox) : 0.f;
          if (g.r == 1 && g.m == 1) {
            Dest d = map_dest(g, z, y, x, c0);
            store_run<16>(g, p.ep, b, d, v, len);
          } else {
            int j = 0;
            while (j < len) {
              Dest d = map_dest(g, z, y, x, c0 + j);
              int run = min(g.cmap - d.c, len - j);
              float seg[16];
#pragma unroll
              for (int k = 0; k < 16; ++k) seg[k] = (j + k < 16) ? v[(j + k) & 15] : 0.f;
              store_run<16>(g, p.ep, b, d, seg, run);
              j += run;
            }
          }
        }
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

  s3_umma_tuning t;
  memset(&t, 0, sizeof(t));
  if (tune) t = *tune;
  p.ep = Epilogue{bias, residual, post_scale, post_shift, y, y_hi, y_lo, t.fmt};
  p.kz = kz;
  p.ntaps = kz * 9;
  p.npad = s3_umma_npad(g.cout);
  p.split = x_lo ? 1 : 0;
  p.fmt = t.fmt;
  p.bo_mode = t.base_offset_mode;
  p.XB = t.box_x > 0 ? t.box_x : 10;
  S3_REQUIRE(p.XB >= 10 && p.XB <= 64, "s3_conv_fwd_umma: box_x must be in [10, 64]");
  const int halves = p.split ? 2 : 1;
  const int Y = g.in[1], X = g.in[2];
  p.planes = kz == 3 ? g.in[0] : g.n;
  p.nb = kz == 3 ? g.n : 1;
  p.plane_pitch = kz == 3 ? g.in[0] + 2 : 0;
  p.nxb = (X + 7) / 8;
  p.w_bytes = (uint32_t)p.npad * 128u;
  const uint32_t w_slab = (p.w_bytes + 1023u) & ~1023u;
  p.WS = t.w_stages > 0 ? t.w_stages : 4;
  S3_REQUIRE(p.WS >= 1 && p.WS <= kMaxWS, "s3_conv_fwd_umma: w_stages must be in [1, %d]", kMaxWS);

  // ---- choose the work-item shape: plane mode (16-row y blocks) vs flat mode (whole height)
  const int max_r_tmem = 512 / p.npad;
  int r_max = t.tiles > 0 ? t.tiles : 4;
  if (r_max > max_r_tmem) r_max = max_r_tmem;
  if (r_max > 8) r_max = 8;
  S3_REQUIRE(r_max >= 1, "s3_conv_fwd_umma: npad %d too wide for TMEM", p.npad);
  const double eff_plane = (double)Y / (((Y + 15) / 16) * 16.0);
  const double eff_flat = (double)Y / (Y + 2.0);
  bool found = false;
  for (int use_flat = (eff_flat > eff_plane + 0.02 ? 1 : 0); use_flat >= 0 && !found; --use_flat) {
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
      const uint32_t need = boxs * halves + (uint32_t)p.WS * w_slab * halves + 2048u;
      if (need > kSmemLimit) continue;
      p.flat = use_flat; p.R = R; p.YB = YB; p.ZB = ZB; p.TS = TS;
      p.box_bytes = box;
      p.box_stride = boxs;
      p.AS = (2 * boxs * halves + (uint32_t)p.WS * w_slab * halves + 2048u <= kSmemLimit) ? 2 : 1;
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
  const uint32_t abox[4] = {64, (uint32_t)p.XB, (uint32_t)p.YB, (uint32_t)p.ZB};
  const uint64_t wdims[3] = {64, (uint64_t)p.npad, (uint64_t)p.ntaps};
  const uint32_t wbox[3] = {64, (uint32_t)p.npad, 1};
  if ((rc = encode_map(&tm_a_hi, x_hi, t.fmt, 4, adims, abox))) return rc;
  if ((rc = encode_map(&tm_w_hi, w_hi, t.fmt, 3, wdims, wbox))) return rc;
  tm_a_lo = tm_a_hi;
  tm_w_lo = tm_w_hi;
  if (p.split) {
    if ((rc = encode_map(&tm_a_lo, x_lo, t.fmt, 4, adims, abox))) return rc;
    if ((rc = encode_map(&tm_w_lo, w_lo, t.fmt, 3, wdims, wbox))) return rc;
  }
  const uint32_t smem = p.box_stride * halves * p.AS + (uint32_t)p.WS * w_slab * halves + 2048u;
  static uint32_t smem_set = 0;
  if (smem > smem_set) {
    S3_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kSmemLimit));
    smem_set = kSmemLimit;
  }
  int ctas = t.max_ctas > 0 ? t.max_ctas : sm_count();
  if (ctas > p.n_items) ctas = p.n_items;
  conv_umma_kernel<<<ctas, kThreads, smem, as_stream(stream)>>>(tm_a_hi, tm_a_lo, tm_w_hi, tm_w_lo,
                                                                p);
  S3_LAUNCH_CHECK("conv_umma_kernel");
  return S3_OK;
}
