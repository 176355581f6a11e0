// tcgen05 implicit-GEMM convolution for the 64-channel body of the sup3r generators:
// 3x3x3 (5-D tensors) or 3x3 (4-D tensors), stride 1, reflect-pad-1 semantics.  This is the
// fused form of the reference's  FlexiblePadding(3, REFLECT) -> Conv(valid) -> Cropping(2)
// [-> LeakyReLU] [-> SpatioTemporalExpansion] [-> SkipConnection add]  layer runs
// (sup3r/configs/spatiotemporal/gen_*.json, executed by sup3r/models/abstract.py:1081-1092).
//
// This file: host side (work-item shape selection, TMA descriptors, kernel choice).  Kernels:
// conv_umma_zring.cu (narrow 3-D outputs) and conv_umma_tile.cu (wide / 2-D / bf16x3);
// data layouts are described in conv_umma_common.cuh and DESIGN.md section 3.
#include <cstring>

#include "conv_umma_common.cuh"

namespace s3 {

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

static int encode_map_strided(CUtensorMap* tm, const void* base, int fmt, int rank,
                              const uint64_t* dims, const uint64_t* strides_bytes,
                              const uint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return S3_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i < rank - 1) gstr[i] = strides_bytes[i];
  }
  CUresult r = enc(tm, fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (strided) failed with %d (rank %d)", (int)r, rank);
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
  // 1: z-concatenated layout [9 (dy,dx)][3 (dz)][npad][64] (ring kernel);  0: tap-major
  // [taps][npad][64] (tile kernel).  `split` = the three-pass bf16x3 operands, which keep the
  // tile kernel (the fp16c pair runs on the ring kernel: pass split = 0 for it).
  return (ndim == 3 && !split && 3 * s3_umma_npad(cout) <= 256) ? 1 : 0;
}

extern "C" int s3_conv_fwd_umma(const s3_conv_desc* d, const void* x_hi, const void* x_lo,
                                const void* w_hi, const void* w_lo, const float* bias,
                                const float* residual, const void* res_hi, const void* res_lo,
                                const float* post_scale, const float* post_shift, float* y,
                                void* y_hi, void* y_lo, const s3_umma_tuning* tune,
                                s3_stream stream) {
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
  const bool mapped16 = y_hi && (g.r > 1 || g.m > 1);   // unpadded 16-bit mapped destination
  if (y_hi && !mapped16)
    S3_REQUIRE(g.fd[1] >= 2 && g.fd[2] >= 2, "s3_conv_fwd_umma: padded output too small");
  if (mapped16)
    S3_REQUIRE(!y_lo && g.rep[0] * g.rep[1] * g.rep[2] == 1 && g.cmap == 8 && g.cout % 8 == 0 &&
                   g.cstride % 8 == 0 && g.coff % 8 == 0 && !post_scale && !residual && !res_hi,
               "s3_conv_fwd_umma: a 16-bit depth_to_space destination needs 8-channel runs, "
               "aligned strides and no lo / residual / affine");

  s3_umma_tuning t;
  memset(&t, 0, sizeof(t));
  if (tune) t = *tune;
  S3_REQUIRE(!(residual && res_hi), "s3_conv_fwd_umma: give the residual as f32 OR as a 16-bit pair");
  S3_REQUIRE(!(res_lo && !res_hi), "s3_conv_fwd_umma: res_lo without res_hi");
  S3_REQUIRE(t.fmt >= 0 && t.fmt <= 2, "s3_conv_fwd_umma: fmt must be 0 (bf16), 1 (fp16) or 2 (fp16c)");
  const bool fp16c = t.fmt == kFmtFp16c;
  p.ep = Epilogue{bias, residual, post_scale, post_shift, y, y_hi, y_lo, t.fmt, res_hi, res_lo,
                  t.acc_scale > 0.f ? t.acc_scale : 1.f};
  p.kz = kz;
  p.ntaps = kz * 9;
  p.npad = s3_umma_npad(g.cout);
  p.split = x_lo ? 1 : 0;
  p.fmt = t.fmt;
  p.XB = t.box_x > 0 ? t.box_x : 10;
  S3_REQUIRE(p.XB >= 10 && p.XB <= 64, "s3_conv_fwd_umma: box_x must be in [10, 64]");
  int halves = p.split ? 2 : 1;
  const int Y = g.in[1], X = g.in[2];
  p.planes = kz == 3 ? g.in[0] : g.n;
  p.nb = kz == 3 ? g.n : 1;
  p.plane_pitch = kz == 3 ? g.in[0] + 2 : 0;
  p.nxb = (X + 7) / 8;
  const bool zcat = s3_umma_weight_layout(g.ndim, g.cout, p.split && !fp16c) == 1;
  p.npass = (zcat && p.split) ? 2 : 1;
  p.w_bytes = (uint32_t)p.npad * 128u * (zcat ? 3u : 1u);
  const uint32_t w_slab = (p.w_bytes + 1023u) & ~1023u;
  p.WS = t.w_stages > 0 ? t.w_stages : (zcat ? 3 : 4);
  S3_REQUIRE(p.WS >= 1 && p.WS <= kMaxWS, "s3_conv_fwd_umma: w_stages must be in [1, %d]", kMaxWS);
  const uint32_t fixed = 3072u;  // alignment slack + barrier block + staged bias

  bool found = false;
  const bool zring = zcat;
  if (zring) {
    // plane-ring pipeline: P plane slots (18 x XB voxels each) + a WS-deep slab ring
    int R = t.tiles > 0 ? t.tiles : 4;
    if (2 * R * p.npad > 512) R = 512 / (2 * p.npad);
    if (R > p.planes) R = p.planes;
    if (R < 1) R = 1;
    const int ws = t.w_stages > 0 ? t.w_stages : 2;
    S3_REQUIRE(ws >= 2 && ws <= 4, "s3_conv_fwd_umma: zring needs w_stages in [2, 4]");
    const uint32_t plane = 18u * (uint32_t)p.XB * 128u;
    // coalescing 16-bit epilogue: 8 warps x 2 KiB staging (see conv_umma_zring.cu)
    // (nearest repeat along x, rep <= 4: the TMA epilogue stores every tile once per replica;
    //  needs the y-halo rows by TMA, i.e. full 8-voxel x tiles)
    const int xrep = g.rep[2];
    const bool rep_ok = g.rep[0] == 1 && g.rep[1] == 1 && xrep >= 1 && xrep <= 4 &&
                        (xrep == 1 || (g.in[2] % 8 == 0 && !(t.box_y & 1024) && !res_hi));
    const bool v2_shape = g.cout == 64 && g.cstride == 64 && g.coff == 0 && g.r == 1 && g.m == 1 &&
                          rep_ok && g.fd[0] >= 4 && g.fd[1] >= 4 && g.in[2] >= 4 && !post_scale;
    p.epi_v4 = (v2_shape && y_hi && !y && !residual && !(t.box_y & 16) && (t.tiles <= 0 || t.tiles == 4) &&
                p.planes >= 4) ? 1 : 0;
    S3_REQUIRE(!res_hi || p.epi_v4, "s3_conv_fwd_umma: a 16-bit residual pair needs the plain "
               "64-channel 16-bit-output configuration");
    // sixteen epilogue warps with one 2 KiB staging box each
    const uint32_t stage_bytes = p.epi_v4 ? 32768u : 0u;
    int P = (int)((kSmemLimit - fixed - 1024u - stage_bytes - (uint32_t)ws * w_slab) / plane);
    if (P > 8) P = 8;
    if (t.ring_slots > 0 && t.ring_slots < P) P = t.ring_slots;
    if (R + 2 > P) R = P - 2;
    S3_REQUIRE(R >= 1, "s3_conv_fwd_umma: zring does not fit shared memory (npad %d)", p.npad);
    p.flat = 0; p.R = R; p.YB = 18; p.ZB = 1; p.TS = 18; p.WS = ws; p.AS = P;
    p.dbg_flags = t.box_y;   // zring: box_y carries experiment flags (see kernel)
    p.ring_fast = (R == 4 && p.planes % 4 == 0 && p.npad == 64 && p.XB == 10 && (P == 7 || P == 6) && ws == 2 &&
                   !(p.dbg_flags & 16)) ? 1 : 0;
    p.box_bytes = plane; p.box_stride = plane;
    found = true;
  } else {
    const int max_r_tmem = 512 / p.npad;
    int r_max = t.tiles > 0 ? t.tiles : (p.npad >= 128 ? 2 : 4);
    if (r_max > max_r_tmem) r_max = max_r_tmem;
    if (r_max > 8) r_max = 8;
    S3_REQUIRE(r_max >= 1, "s3_conv_fwd_umma: npad %d too wide for TMEM", p.npad);
    const double eff_plane = (double)Y / (((Y + 15) / 16) * 16.0);
    const double eff_flat = (double)Y / (Y + 2.0);
    if (fp16c && p.split && kz == 3 && t.tiles <= 0 && t.w_stages <= 0 && r_max >= 2 &&
        p.XB == 10 && !(t.box_y & 16)) {
      // fp16c wide head: the fp16 pass and the e4m3 pass run one after the other through the
      // same activation box (one operand tensor resident at a time) and a 4-deep weight ring
      const uint32_t box = 4u * 18u * 10u * 128u;
      const uint32_t boxs = (box + 1023u) & ~1023u;
      for (int as = 2; as >= 1 && !found; --as) {
        if (boxs * as + 4u * w_slab + fixed > kSmemLimit) continue;
        p.flat = 0; p.R = 2; p.YB = 18; p.ZB = 4; p.TS = 18; p.WS = 4; p.AS = as;
        p.box_bytes = box; p.box_stride = boxs;
        p.seq2 = 1;
        halves = 1;
        found = true;
      }
    }
    if (!found && p.split && kz == 3 && t.tiles <= 0 && t.w_stages <= 0) {
      // split operands: the weight taps are re-streamed per work item, so more tiles per item
      // beats a deeper weight ring (measured 64 -> 64 body conv: R = 2 / 2 stages 450 us,
      // R = 1 / 4 stages 547 us)
      for (int R = r_max; R >= 2 && !found; --R)
        for (int ws = 3; ws >= 2 && !found; --ws) {
          const uint32_t box = (uint32_t)(R + 2) * 18u * p.XB * 128u;
          const uint32_t boxs = (box + 1023u) & ~1023u;
          if (boxs * halves + (uint32_t)ws * w_slab * halves + fixed > kSmemLimit) continue;
          p.flat = 0; p.R = R; p.YB = 18; p.ZB = R + 2; p.TS = 18; p.WS = ws;
          p.box_bytes = box; p.box_stride = boxs;
          p.AS = 1;
          found = true;
        }
    }
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
  if (!zring) p.dbg_flags = t.ring_slots;   // tile kernel: ring_slots carries experiment flags
  p.tile_fast = (!zcat && kz == 3 && (!p.split || p.seq2) && !p.flat && p.R == 2 && p.XB == 10 &&
                 p.YB == 18 && p.WS == 4 && p.ntaps == 27 && !(t.box_y & 16)) ? 1 : 0;
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
  p.trace = reinterpret_cast<long long*>(t.trace);

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
  uint32_t smem = p.box_stride * halves * p.AS + (uint32_t)p.WS * w_slab * halves + fixed;
  if (zring) smem = ((p.box_stride * (uint32_t)p.AS + 1023u) & ~1023u) + (uint32_t)p.WS * w_slab + fixed +
                    (p.epi_v4 ? 32768u : 0u);
  int ctas = t.max_ctas > 0 ? t.max_ctas : sm_count();
  if (ctas > p.n_items) ctas = p.n_items;
  // epilogue specialisation: fast paths only when their preconditions hold for EVERY row
  int epi = EPI_GENERIC;
  const bool aligned = (g.cstride % 8 == 0) && (g.coff % 8 == 0);
  const bool plain = g.r == 1 && g.m == 1 && g.rep[0] == 1 && g.rep[1] == 1 && g.rep[2] <= 3;
  const bool roomy = g.fd[1] >= 4 && g.fd[2] >= 4 && (g.ndim == 2 || g.fd[0] >= 4);
  if (plain && aligned && roomy && g.cout % 16 == 0 && !post_scale) {
    epi = EPI_PLAIN;
  } else if ((g.r > 1 || g.m > 1) && g.rep[0] * g.rep[1] * g.rep[2] == 1 && (y || mapped16) &&
             (!y_hi || mapped16) &&
             (g.cmap == 4 || g.cmap == 8 || g.cmap % 16 == 0) &&
             g.cout % (g.cmap < 16 ? g.cmap : 16) == 0 && g.cbase % 16 == 0 &&
             g.cstride % 4 == 0 && g.coff % 4 == 0) {
    epi = EPI_D2S;
  }
  S3_REQUIRE(!(fp16c && (y_lo || res_lo)) ||
                 (epi == EPI_PLAIN && g.cout == 64 && g.cstride == 64 && g.coff == 0 && !post_scale),
             "s3_conv_fwd_umma: fp16c corr rows (y_lo / res_lo) need the plain 64-channel configuration");
  S3_REQUIRE(!g.res_pre || (!zring && !zcat && epi == EPI_PLAIN && g.cout == 64 && g.cstride == 64 &&
                            g.coff == 0 && !post_scale && residual),
             "s3_conv_fwd_umma: res_pre_act needs the plain 64-channel tile-kernel configuration "
             "(2-D or split precision) with an f32 residual");
  if (zring) {
    CUtensorMap em[6];
    memset(em, 0, sizeof(em));
    p.epi_row_tma = (p.epi_v4 && g.in[2] % 8 == 0 && !(t.box_y & 1024)) ? 1 : 0;
    if (p.epi_v4 && g.rep[2] > 1) {
      // 5-D views: channel | replica rx | conv x | y | plane, output x = x * rep + rx.  Interior
      // view (tiles) and a view whose y axis is the padded one (y-halo rows); both start one voxel
      // into the padded x axis.
      const int rep = g.rep[2];
      const uint64_t row_pitch = (uint64_t)(g.fd[2] + 2) * 128;
      const uint64_t estr[4] = {128, (uint64_t)rep * 128, row_pitch, (uint64_t)(g.fd[1] + 2) * row_pitch};
      const uint64_t edims[5] = {64, (uint64_t)rep, (uint64_t)g.in[2], (uint64_t)g.fd[1], total_planes};
      const uint64_t rdims[5] = {64, (uint64_t)rep, (uint64_t)g.in[2], (uint64_t)g.fd[1] + 2, total_planes};
      const uint32_t ebox[5] = {32, 1, 8, 4, 1}, rbox[5] = {32, 1, 8, 1, 1};
      const size_t shift = (size_t)row_pitch + 128, rshift = 128;
      void* outs[2] = {y_hi, y_lo};
      for (int i = 0; i < 2; ++i) {
        if (!outs[i]) continue;
        if ((rc = encode_map_strided(&em[2 + i], static_cast<uint8_t*>(outs[i]) + shift, t.fmt, 5,
                                     edims, estr, ebox, CU_TENSOR_MAP_SWIZZLE_64B)))
          return rc;
        if ((rc = encode_map_strided(&em[4 + i], static_cast<uint8_t*>(outs[i]) + rshift, t.fmt, 5,
                                     rdims, estr, rbox, CU_TENSOR_MAP_SWIZZLE_64B)))
          return rc;
      }
    } else if (p.epi_v4) {
      // interior views (x + 1, y + 1) of the padded tensors: tile coordinates are plain voxel
      // indices, ragged tiles are clipped by the map extents
      const uint64_t edims[4] = {64, (uint64_t)g.fd[2], (uint64_t)g.fd[1], total_planes};
      const uint64_t estr[3] = {128, (uint64_t)(g.fd[2] + 2) * 128,
                                (uint64_t)(g.fd[1] + 2) * (g.fd[2] + 2) * 128};
      const uint32_t ebox[4] = {32, 8, 4, 1};   // one warp's 32 rows x 32 channels, SWIZZLE_64B
      const size_t shift = ((size_t)(g.fd[2] + 2) + 1) * 128;
      const void* ptrs[4] = {res_hi, res_lo, y_hi, y_lo};
      for (int i = 0; i < 4; ++i) {
        if (!ptrs[i]) continue;
        if ((rc = encode_map_strided(&em[i], static_cast<const uint8_t*>(ptrs[i]) + shift, t.fmt, 4,
                                     edims, estr, ebox, CU_TENSOR_MAP_SWIZZLE_64B)))
          return rc;
      }
      if (p.epi_row_tma) {
        // padded view (halo included) with one-row boxes for the y mirrors
        const uint64_t rdims[4] = {64, (uint64_t)g.fd[2] + 2, (uint64_t)g.fd[1] + 2, total_planes};
        const uint32_t rbox[4] = {32, 8, 1, 1};
        const void* rptrs[2] = {y_hi, y_lo};
        for (int i = 0; i < 2; ++i) {
          if (!rptrs[i]) continue;
          if ((rc = encode_map_strided(&em[4 + i], rptrs[i], t.fmt, 4, rdims, estr, rbox,
                                       CU_TENSOR_MAP_SWIZZLE_64B)))
            return rc;
        }
      }
    }
    const CUtensorMap maps[4] = {tm_a_hi, tm_w_hi, tm_a_lo, tm_w_lo};
    rc = launch_umma_zring(p, maps, em, epi, ctas, smem, as_stream(stream));
  }
  else
    rc = launch_umma_tile(p, tm_a_hi, tm_a_lo, tm_w_hi, tm_w_lo, epi, ctas, smem,
                          as_stream(stream));
  if (rc) return rc;

  return S3_OK;
}
