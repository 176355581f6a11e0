// Narrow 3x3x3 stride-1 "same" convolution (cin <= 8, cout <= 8) on the warp-level tensor-core
// path (mma.sync m16n8k16, bf16 operands, fp32 accumulate): the output convolution of the
// generators at full high resolution (north-star model: 8 -> 4 channels over 80x80x288 voxels
// per chunk).  The layer is far too narrow for tcgen05 (N = 4) and, on CUDA cores, bound by the
// fp32 FMA pipe (216 x 4 FMAs per voxel; profiles/r01_small_full.md: 0.93 ms per 8 chunks, FMA
// pipe 43 %); as an implicit GEMM with K = 27 taps x 8 channels = 216 (padded to 224) it needs
// 14 MMAs per 16 voxels.
//   * a CTA stages the input halo tile as bf16 [z][y][x][8 ch] (16 B per voxel; fp32 -> bf16,
//     channel padding and REFLECT / SYMMETRIC / zero folding done while staging);
//   * im2col costs nothing: one ldmatrix.x4 gathers the A fragment of a k-step (16 voxels x
//     2 taps x 8 channels) straight from 32 per-lane voxel addresses -- 8 consecutive x voxels
//     are 128 contiguous bytes, so every 8x8 sub-matrix read is bank-conflict free;
//   * the B fragments (all 14 k-steps, 28 registers) stay in registers for the whole CTA;
//   * a warp owns one (z, 16-voxel x segment) column of the tile and walks the y rows two at a
//     time (two independent accumulator sets hide the ldmatrix latency).
// Used for precision "bf16" (bf16 operands like every other tensor-core layer of that mode) and
// "fp16c" (fp16 operands: the one layer of that mode without correction rows -- its rounding does
// not compound through later layers); fp32 / bf16x3 keep the fp32 kernel in conv_small.cu.
// Replaces FlexiblePadding -> Conv3D -> Cropping3D at the end of
// sup3r/configs/spatiotemporal/gen_*.json as executed by sup3r/models/abstract.py:1081-1092.
#include <cuda_bf16.h>

#include "common.cuh"

namespace s3 {

namespace {

constexpr int MT_X = 32, MT_Y = 16, MT_Z = 4;            // output tile
constexpr int MH_X = MT_X + 2, MH_Y = MT_Y + 2, MH_Z = MT_Z + 2;
constexpr int M_THREADS = 256;                           // 8 warps: (z 0..3) x (x half 0..1)
constexpr int M_KSTEPS = 14;                             // 28 taps (27 + 1 zero tap) / 2
constexpr int M_HALO_VOX = MH_X * MH_Y * MH_Z;           // 3672 voxels x 16 B = 58 752 B

__device__ __forceinline__ int fold_idx(int q, int n, int mode, bool* ok) {
  if (q >= 0 && q < n) return q;
  if (mode == S3_PAD_REFLECT) return q < 0 ? -q : 2 * n - 2 - q;
  if (mode == S3_PAD_SYMMETRIC) return q < 0 ? -q - 1 : 2 * n - 1 - q;
  *ok = false;
  return 0;
}

// kF16: fp16 operands (the fp16c mode's output layer), else bf16
template <bool kF16>
__device__ __forceinline__ uint32_t pack_16x2(float a, float b) {
  if (kF16) return f16x2_sat(a, b);
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

template <bool kF16>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                          uint32_t b1) {
  if (kF16)
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
        "{%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
        "{%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __noinline__ float act_slow(float v, int act, float alpha) { return apply_act(v, act, alpha); }

// kTwo (cout <= 4): one accumulator row holds TWO neighbouring output voxels (n = voxel * 4 +
// channel) computed from a 4-wide x window (K = 9 x 4 taps x 8 ch = 18 k-steps per voxel pair
// instead of 2 x 14): the kernel is bound by the legacy HMMA issue rate (~32 cycles per
// m16n8k16 and SM sub-partition on B200), so fewer MMAs per voxel is the lever.
template <bool kTwo, bool kF16>
__global__ void __launch_bounds__(M_THREADS, 3)
conv_small_mma_kernel(const ConvGeom g, const float* __restrict__ x,
                      const uint4* __restrict__ x16, const float* __restrict__ w,
                      const Epilogue ep, int tiles_x, int tiles_y, int tiles_z) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint4* sin = reinterpret_cast<uint4*>(smem);             // [MH_Z][MH_Y][MH_X] voxels of 8 bf16
  uint4* szero = sin + M_HALO_VOX;                         // 8 zero voxels for the padding tap
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int t = blockIdx.x;
  const int bx = t % tiles_x; t /= tiles_x;
  const int by = t % tiles_y; t /= tiles_y;
  const int bz = t % tiles_z;
  const int b = t / tiles_z;
  const int x0 = bx * MT_X, y0 = by * MT_Y, z0 = bz * MT_Z;
  const int Z = g.in[0], Y = g.in[1], X = g.in[2];
  const int cin = g.cin, cout = g.cout;

  // ---- stage the halo tile (fp32 -> bf16, channels padded to 8): a warp takes (z, y) rows, a
  // lane the voxels x = lane and lane + 32 of the row; four rows in flight per lane
  if (tid < 8) szero[tid] = make_uint4(0, 0, 0, 0);
  if (x16 != nullptr) {
    // bf16 input with exactly 8 channels (one 16-byte voxel): asynchronous 16-byte copies with
    // the folded source address, all in flight at once, no registers
    bool okx0 = true, okx1 = true;
    int gx0 = fold_idx(x0 + lane - 1, X, g.pad_mode, &okx0);
    int gx1 = fold_idx(x0 + lane + 32 - 1, X, g.pad_mode, &okx1);
    okx0 = okx0 && gx0 >= 0 && gx0 < X;
    okx1 = okx1 && gx1 >= 0 && gx1 < X && lane + 32 < MH_X;
    constexpr int kRows = MH_Z * MH_Y;
    const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(sin);
    for (int r = warp; r < kRows; r += M_THREADS / 32) {
      const int hz = r / MH_Y, hy = r - hz * MH_Y;
      bool ok = true;
      const int gz = fold_idx(z0 + hz - 1, Z, g.pad_mode, &ok);
      const int gy = fold_idx(y0 + hy - 1, Y, g.pad_mode, &ok);
      ok = ok && gz >= 0 && gz < Z && gy >= 0 && gy < Y;
      const uint4* row = x16 + (((size_t)b * Z + (ok ? gz : 0)) * Y + (ok ? gy : 0)) * (size_t)X;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int hx = lane + 32 * e;
        if (hx >= MH_X) continue;
        const bool okv = ok && (e == 0 ? okx0 : okx1);
        const uint32_t dst = sdst + (uint32_t)(r * MH_X + hx) * 16u;
        if (okv) {
          const uint4* src = row + (e == 0 ? gx0 : gx1);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        } else {
          sin[r * MH_X + hx] = make_uint4(0, 0, 0, 0);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    bool okx0 = true, okx1 = true;
    int gx0 = fold_idx(x0 + lane - 1, X, g.pad_mode, &okx0);
    int gx1 = fold_idx(x0 + lane + 32 - 1, X, g.pad_mode, &okx1);
    okx0 = okx0 && gx0 >= 0 && gx0 < X;
    okx1 = okx1 && gx1 >= 0 && gx1 < X && lane + 32 < MH_X;
    constexpr int kRows = MH_Z * MH_Y;          // 108
    constexpr int kBatch = 4;
    for (int r0 = warp * kBatch; r0 < kRows; r0 += (M_THREADS / 32) * kBatch) {
      float4 va[kBatch][2], vb[kBatch][2];
      bool okr[kBatch];
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const int r = r0 + q;
        const int hz = r / MH_Y, hy = r - hz * MH_Y;
        bool ok = r < kRows;
        const int gz = fold_idx(z0 + hz - 1, Z, g.pad_mode, &ok);
        const int gy = fold_idx(y0 + hy - 1, Y, g.pad_mode, &ok);
        ok = ok && gz >= 0 && gz < Z && gy >= 0 && gy < Y;
        okr[q] = ok;
        const float* row = x + (((size_t)b * Z + (ok ? gz : 0)) * Y + (ok ? gy : 0)) * (size_t)X * cin;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool okv = ok && (e == 0 ? okx0 : okx1);
          const float* src = row + (size_t)(e == 0 ? gx0 : gx1) * cin;
          va[q][e] = vb[q][e] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (okv) {
            if (cin == 8) {
              va[q][e] = __ldg(reinterpret_cast<const float4*>(src));
              vb[q][e] = __ldg(reinterpret_cast<const float4*>(src) + 1);
            } else {
              float f[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) f[c] = c < cin ? __ldg(src + c) : 0.f;
              va[q][e] = make_float4(f[0], f[1], f[2], f[3]);
              vb[q][e] = make_float4(f[4], f[5], f[6], f[7]);
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const int r = r0 + q;
        if (r >= kRows) break;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int hx = lane + 32 * e;
          if (hx >= MH_X) continue;
          uint4 v;
          v.x = pack_16x2<kF16>(va[q][e].x, va[q][e].y);
          v.y = pack_16x2<kF16>(va[q][e].z, va[q][e].w);
          v.z = pack_16x2<kF16>(vb[q][e].x, vb[q][e].y);
          v.w = pack_16x2<kF16>(vb[q][e].z, vb[q][e].w);
          sin[r * MH_X + hx] = v;
        }
      }
      (void)okr;
    }
  }
  __syncthreads();

  // ---- B fragments: k = tap * 8 + ci (tap 27 = zero), n = output channel (>= cout -> 0)
  // thread (grp = lane / 4, t4 = lane % 4) holds b0 = {B[2 t4][grp], B[2 t4 + 1][grp]},
  // b1 = {B[2 t4 + 8][grp], B[2 t4 + 9][grp]} of every k-step
  constexpr int kSteps = kTwo ? 18 : M_KSTEPS;
  uint32_t bf[kSteps][2];
  {
    const int grp = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int j = 0; j < kSteps; ++j) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int tap = 2 * j + hh;
        float v0 = 0.f, v1 = 0.f;
        const int c0 = 2 * t4, c1 = 2 * t4 + 1;
        if (kTwo) {
          // tap = (dz, dy, dxw) over a 4-wide x window; n = grp = voxel * 4 + channel
          const int dz = tap / 12, dy = (tap / 4) % 3, dxw = tap % 4;
          const int vsel = grp >> 2, co = grp & 3, dx = dxw - vsel;
          if (dx >= 0 && dx < 3 && co < cout) {
            const int t3 = (dz * 3 + dy) * 3 + dx;
            if (c0 < cin) v0 = __ldg(w + ((size_t)t3 * cin + c0) * cout + co);
            if (c1 < cin) v1 = __ldg(w + ((size_t)t3 * cin + c1) * cout + co);
          }
        } else if (tap < 27 && grp < cout) {
          if (c0 < cin) v0 = __ldg(w + ((size_t)tap * cin + c0) * cout + grp);
          if (c1 < cin) v1 = __ldg(w + ((size_t)tap * cin + c1) * cout + grp);
        }
        bf[j][hh] = pack_16x2<kF16>(v0, v1);
      }
    }
  }

  // ---- main loop.  !kTwo: warp = (z, 16-voxel x half), all 16 y rows; kTwo: warp = (z, y half),
  // the 32 x voxels of a row are 16 accumulator rows of two voxels each.  Two y rows per step.
  const int wz = warp >> 1;
  const int wx = kTwo ? 0 : (warp & 1) * 16;
  const int y_begin = kTwo ? (warp & 1) * (MT_Y / 2) : 0;
  const int y_end = kTwo ? y_begin + MT_Y / 2 : MT_Y;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sin);
  const uint32_t zero_addr = (uint32_t)__cvta_generic_to_shared(szero) + (uint32_t)(lane & 7) * 16u;
  // ldmatrix row address of this lane: matrix (lane / 8): rows (lane % 8) + 8 * (mat & 1) of
  // tap 2 j + (mat >> 1)
  const int mat = lane >> 3;
  const int lrow = (lane & 7) + 8 * (mat & 1);
  const int ltap = mat >> 1;
  const int grp = lane >> 2, t4 = lane & 3;
  const int oz = z0 + wz;

  // per-lane shared-memory offset of every k-step's row address at y row 0
  uint32_t toff[kSteps];
#pragma unroll
  for (int j = 0; j < kSteps; ++j) {
    const int tap = 2 * j + ltap;
    if (kTwo) {
      const int dz = tap / 12, dy = (tap / 4) % 3, dxw = tap % 4;
      toff[j] = sbase + (uint32_t)((((wz + dz) * MH_Y + dy) * MH_X + 2 * lrow + dxw) * 16);
    } else {
      const int dz = tap / 9, dy = (tap / 3) % 3, dx = tap % 3;
      toff[j] = sbase + (uint32_t)((((wz + dz) * MH_Y + dy) * MH_X + wx + lrow + dx) * 16);
    }
  }
  const bool zero_last = !kTwo && ltap == 1;   // tap 27 (zero) = last k-step, second half
  // per-thread epilogue constants
  const int act = g.act, cstride = g.cstride, coff = g.coff;
  const float alpha = g.alpha;
  const bool pair_ok = ((cstride | coff) & 1) == 0;
  const int ec0 = kTwo ? (2 * t4) & 3 : 2 * t4, ec1 = ec0 + 1;
  const int vsel = kTwo ? t4 >> 1 : 0;
  const float bias0 = (ep.bias && ec0 < cout) ? ep.bias[ec0] : 0.f;
  const float bias1 = (ep.bias && ec1 < cout) ? ep.bias[ec1] : 0.f;
  const float sc0 = (ep.post_scale && ec0 < cout) ? ep.post_scale[ec0] : 1.f;
  const float sc1 = (ep.post_scale && ec1 < cout) ? ep.post_scale[ec1] : 1.f;
  const float sh0 = (ep.post_scale && ep.post_shift && ec0 < cout) ? ep.post_shift[ec0] : 0.f;
  const float sh1 = (ep.post_scale && ep.post_shift && ec1 < cout) ? ep.post_shift[ec1] : 0.f;

#pragma unroll 1
  for (int yy = y_begin; yy < y_end; yy += 2) {
    float acc[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[u][q] = 0.f;
#pragma unroll
    for (int j = 0; j < kSteps; ++j) {
      uint32_t a[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t addr = (j == kSteps - 1 && zero_last)
                                  ? zero_addr
                                  : toff[j] + (uint32_t)((yy + u) * MH_X * 16);
        ldmatrix_x4(addr, a[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) mma_16816<kF16>(acc[u], a[u], bf[j][0], bf[j][1]);
    }
    // ---- epilogue: thread holds rows grp and grp + 8, accumulator columns 2 t4, 2 t4 + 1
    if (oz < Z && ec0 < cout) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int oy = y0 + yy + u;
        if (oy >= Y) continue;
        const size_t row0 = (((size_t)b * Z + oz) * Y + oy) * X;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int ox = kTwo ? x0 + 2 * (grp + 8 * hh) + vsel : x0 + wx + grp + 8 * hh;
          if (ox >= X) continue;
          const size_t vox = row0 + ox;
          float v0 = acc[u][2 * hh] + bias0, v1 = acc[u][2 * hh + 1] + bias1;
          if (act == S3_ACT_LEAKY) {
            v0 = v0 >= 0.f ? v0 : alpha * v0;
            v1 = v1 >= 0.f ? v1 : alpha * v1;
          } else if (act == S3_ACT_RELU) {
            v0 = fmaxf(v0, 0.f);
            v1 = fmaxf(v1, 0.f);
          } else if (act != S3_ACT_NONE) {
            v0 = act_slow(v0, act, alpha);
            v1 = act_slow(v1, act, alpha);
          }
          if (ep.residual) {
            v0 += ep.residual[vox * cout + ec0];
            if (ec1 < cout) v1 += ep.residual[vox * cout + ec1];
          }
          v0 = v0 * sc0 + sh0;
          v1 = v1 * sc1 + sh1;
          float* dst = ep.y + vox * cstride + coff + ec0;
          if (ec1 < cout && pair_ok) {
            *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
          } else {
            dst[0] = v0;
            if (ec1 < cout) dst[1] = v1;
          }
        }
      }
    }
  }
}

}  // namespace

// Returns 1 if handled, 0 if the shape is not covered, < 0 on error.
template <bool kTwo, bool kF16>
static int launch_small_mma(const ConvGeom& g, const float* x, const void* x16, const float* w,
                            const Epilogue& ep, unsigned blocks, size_t smem, int tx, int ty, int tz,
                            cudaStream_t st) {
  static bool set = false;
  if (!set) {
    S3_CUDA(cudaFuncSetAttribute(conv_small_mma_kernel<kTwo, kF16>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = true;
  }
  conv_small_mma_kernel<kTwo, kF16><<<blocks, M_THREADS, smem, st>>>(
      g, x, static_cast<const uint4*>(x16), w, ep, tx, ty, tz);
  S3_CUDA(cudaGetLastError());
  return 1;
}

// ep.fmt: 0 = bf16 operands, otherwise fp16
int try_conv_small_mma(const ConvGeom& g, const float* x, const void* x16, const float* w,
                       const Epilogue& ep, cudaStream_t st) {
  if ((x == nullptr) == (x16 == nullptr)) return 0;
  if (x16 && g.cin != 8) return 0;
  if (ep.y == nullptr || ep.y_hi != nullptr || ep.res_hi != nullptr) return 0;
  if (g.ndim != 3 || g.r != 1 || g.m != 1 || g.rep[0] * g.rep[1] * g.rep[2] != 1) return 0;
  if (g.cout > 8 || g.cin > 8) return 0;
  for (int i = 0; i < 3; ++i)
    if (g.k[i] != 3 || g.st[i] != 1 || g.pl[i] != 1 || g.ph[i] != 1) return 0;
  if (g.pad_mode == S3_PAD_REFLECT && (g.in[0] < 2 || g.in[1] < 2 || g.in[2] < 2)) return 0;
  const size_t smem = sizeof(uint4) * (M_HALO_VOX + 8);
  const int tx = (g.in[2] + MT_X - 1) / MT_X, ty = (g.in[1] + MT_Y - 1) / MT_Y;
  const int tz = (g.in[0] + MT_Z - 1) / MT_Z;
  const long long blocks = (long long)tx * ty * tz * g.n;
  if (blocks > 0x7fffffffLL) {
    set_error("conv_small_mma: grid too large");
    return S3_ERR_INVALID;
  }
  const bool f16 = ep.fmt != 0;
  if (g.cout <= 4)
    return f16 ? launch_small_mma<true, true>(g, x, x16, w, ep, (unsigned)blocks, smem, tx, ty, tz, st)
               : launch_small_mma<true, false>(g, x, x16, w, ep, (unsigned)blocks, smem, tx, ty, tz, st);
  return f16 ? launch_small_mma<false, true>(g, x, x16, w, ep, (unsigned)blocks, smem, tx, ty, tz, st)
             : launch_small_mma<false, false>(g, x, x16, w, ep, (unsigned)blocks, smem, tx, ty, tz, st);
}

}  // namespace s3
