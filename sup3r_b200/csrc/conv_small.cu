// Small-channel 3x3[x3] stride-1 "same" convolution on CUDA cores for the HBM-side layers of
// the generators: the output convolution at full high resolution (e.g. 8 -> 4 channels at
// 80x80x288 per chunk: 1.84 M voxels, 88 MB of traffic, 3.2 GFLOP) and similar narrow layers.
// The generic implicit-GEMM kernel wastes most of its tile on such shapes; here
//   * a CTA stages the input halo tile [cin/4][z][y][x] (float4 per voxel-quad of channels) in
//     shared memory once (coalesced global reads, REFLECT / zero folding done while staging),
//   * a warp spans 32 consecutive x voxels, each thread owns 4 consecutive y outputs of one x,
//     so every staged float4 is reused by 3 dy taps x 4 outputs from registers and all
//     shared-memory reads are conflict-free (consecutive float4 per lane) or broadcasts,
//   * weights [tap][cin][COP] sit in shared memory and are read as warp-wide broadcasts,
//   * stores are coalesced (32 lanes x COP floats).
// fp32 in / fp32 accumulate / fp32 out; bias, activation, residual and the un-normalisation
// affine are fused.  Replaces the last FlexiblePadding -> Conv3D -> Cropping3D run of every
// generator config (sup3r/configs/spatiotemporal/gen_*.json).
#include "common.cuh"

namespace s3 {

constexpr int SM_TX = 32, SM_TY = 8, SM_TZ = 4;      // output tile (x, y, z)
constexpr int SM_YPT = 4;                            // y outputs per thread
constexpr int SM_THREADS = SM_TX * (SM_TY / SM_YPT) * SM_TZ;   // 256

__device__ __forceinline__ int fold1(int q, int n, int mode, bool* ok) {
  if (q >= 0 && q < n) return q;
  if (mode == S3_PAD_REFLECT) return q < 0 ? -q : 2 * n - 2 - q;
  if (mode == S3_PAD_SYMMETRIC) return q < 0 ? -q - 1 : 2 * n - 1 - q;
  *ok = false;
  return 0;
}

template <int CIN4, int COP, int KZ>
__global__ void __launch_bounds__(SM_THREADS)
conv_small_kernel(const ConvGeom g, const float* __restrict__ x, const float* __restrict__ w,
                  const Epilogue ep, int tiles_x, int tiles_y, int tiles_z) {
  constexpr int HX = SM_TX + 2, HY = SM_TY + 2, HZ = SM_TZ + (KZ == 3 ? 2 : 0);
  constexpr int NTAP = KZ * 9;
  extern __shared__ __align__(16) float4 smem4[];
  float4* sin = smem4;                                      // [CIN4][HZ][HY][HX]
  float* sw = reinterpret_cast<float*>(smem4 + CIN4 * HZ * HY * HX);   // [NTAP][CIN4*4][COP]

  const int tid = threadIdx.x;
  int t = blockIdx.x;
  const int bx = t % tiles_x; t /= tiles_x;
  const int by = t % tiles_y; t /= tiles_y;
  const int bz = t % tiles_z;
  const int b = t / tiles_z;
  const int x0 = bx * SM_TX, y0 = by * SM_TY, z0 = bz * SM_TZ;
  const int Z = g.in[0], Y = g.in[1], X = g.in[2];
  const int cin = g.cin, cout = g.cout;

  // ---- stage weights (zero padded to CIN4*4 x COP) and the input halo
  for (int i = tid; i < NTAP * CIN4 * 4 * COP; i += SM_THREADS) {
    const int co = i % COP, ci = (i / COP) % (CIN4 * 4), tap = i / (COP * CIN4 * 4);
    sw[i] = (ci < cin && co < cout) ? __ldg(w + ((size_t)tap * cin + ci) * cout + co) : 0.f;
  }
  const int pz = KZ == 3 ? 1 : 0;
  for (int i = tid; i < CIN4 * HZ * HY * HX; i += SM_THREADS) {
    // global-friendly order: channel quad fastest, then x
    const int c4 = i % CIN4;
    int r = i / CIN4;
    const int hx = r % HX; r /= HX;
    const int hy = r % HY;
    const int hz = r / HY;
    bool ok = true;
    const int gz = KZ == 3 ? fold1(z0 + hz - pz, Z, g.pad_mode, &ok) : z0 + hz;
    const int gy = fold1(y0 + hy - 1, Y, g.pad_mode, &ok);
    const int gx = fold1(x0 + hx - 1, X, g.pad_mode, &ok);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok && gz >= 0 && gz < Z && gy >= 0 && gy < Y && gx >= 0 && gx < X) {
      const float* src = x + ((((size_t)b * Z + gz) * Y + gy) * X + gx) * cin + c4 * 4;
      if (c4 * 4 + 3 < cin && (cin & 3) == 0) {
        v = __ldg(reinterpret_cast<const float4*>(src));
      } else {
        if (c4 * 4 + 0 < cin) v.x = __ldg(src + 0);
        if (c4 * 4 + 1 < cin) v.y = __ldg(src + 1);
        if (c4 * 4 + 2 < cin) v.z = __ldg(src + 2);
        if (c4 * 4 + 3 < cin) v.w = __ldg(src + 3);
      }
    }
    sin[((c4 * HZ + hz) * HY + hy) * HX + hx] = v;
  }
  __syncthreads();

  const int lx = tid % SM_TX;
  const int ly = (tid / SM_TX) % (SM_TY / SM_YPT);
  const int lz = tid / (SM_TX * (SM_TY / SM_YPT));
  float acc[SM_YPT][COP];
#pragma unroll
  for (int j = 0; j < SM_YPT; ++j)
#pragma unroll
    for (int c = 0; c < COP; ++c) acc[j][c] = 0.f;

#pragma unroll 1
  for (int dz = 0; dz < KZ; ++dz) {
#pragma unroll 1
    for (int c4 = 0; c4 < CIN4; ++c4) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        float4 in[SM_YPT + 2];
        const float4* col = sin + ((c4 * HZ + lz + dz) * HY + ly * SM_YPT) * HX + lx + dx;
#pragma unroll
        for (int r = 0; r < SM_YPT + 2; ++r) in[r] = col[r * HX];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const float* wt = sw + (((dz * 3 + dy) * 3 + dx) * CIN4 * 4 + c4 * 4) * COP;
          float wv[4][COP];
#pragma unroll
          for (int ci = 0; ci < 4; ++ci)
#pragma unroll
            for (int c = 0; c < COP; ++c) wv[ci][c] = wt[ci * COP + c];
#pragma unroll
          for (int j = 0; j < SM_YPT; ++j) {
            const float4 v = in[j + dy];
#pragma unroll
            for (int c = 0; c < COP; ++c) {
              acc[j][c] = fmaf(v.x, wv[0][c], acc[j][c]);
              acc[j][c] = fmaf(v.y, wv[1][c], acc[j][c]);
              acc[j][c] = fmaf(v.z, wv[2][c], acc[j][c]);
              acc[j][c] = fmaf(v.w, wv[3][c], acc[j][c]);
            }
          }
        }
      }
    }
  }

  // ---- epilogue: bias / activation / residual / affine, coalesced stores
  const int ox = x0 + lx, oz = z0 + lz;
  if (ox >= X || oz >= Z) return;
#pragma unroll
  for (int j = 0; j < SM_YPT; ++j) {
    const int oy = y0 + ly * SM_YPT + j;
    if (oy >= Y) continue;
    const size_t vox = (((size_t)b * Z + oz) * Y + oy) * X + ox;
    float v[COP];
#pragma unroll
    for (int c = 0; c < COP; ++c) v[c] = c < cout ? finish(g, ep, acc[j][c], c, vox) : 0.f;
    float* dst = ep.y + vox * g.cstride + g.coff;
    if (COP == 4 && cout == 4 && g.cstride == 4 && g.coff == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int c = 0; c < COP; ++c)
        if (c < cout) dst[c] = v[c];
    }
  }
}

template <int CIN4, int COP, int KZ>
static int launch_small(const ConvGeom& g, const float* x, const float* w, const Epilogue& ep,
                        cudaStream_t st) {
  constexpr int HX = SM_TX + 2, HY = SM_TY + 2, HZ = SM_TZ + (KZ == 3 ? 2 : 0);
  const size_t smem = sizeof(float4) * CIN4 * HZ * HY * HX + sizeof(float) * KZ * 9 * CIN4 * 4 * COP;
  static bool set = false;
  if (!set) {
    S3_CUDA(cudaFuncSetAttribute(conv_small_kernel<CIN4, COP, KZ>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = true;
  }
  const int tx = (g.in[2] + SM_TX - 1) / SM_TX, ty = (g.in[1] + SM_TY - 1) / SM_TY;
  const int tz = KZ == 3 ? (g.in[0] + SM_TZ - 1) / SM_TZ : 1;
  // 2-D: the z tile dimension runs over the batch instead (KZ == 1 -> Z == 1)
  const long long blocks = (long long)tx * ty * tz * g.n;
  if (blocks > 0x7fffffffLL) {
    set_error("conv_small: grid too large");
    return S3_ERR_INVALID;
  }
  conv_small_kernel<CIN4, COP, KZ><<<(unsigned)blocks, SM_THREADS, smem, st>>>(g, x, w, ep, tx, ty,
                                                                               tz);
  S3_CUDA(cudaGetLastError());
  return S3_OK;
}

// Returns 1 if handled, 0 if the shape is not covered (caller falls back to the generic
// kernel), < 0 on error.
int try_conv_small(const ConvGeom& g, const float* x, const float* w, const Epilogue& ep,
                   cudaStream_t st) {
  if (ep.y == nullptr || ep.y_hi != nullptr) return 0;
  if (g.r != 1 || g.m != 1 || g.rep[0] * g.rep[1] * g.rep[2] != 1) return 0;
  if (g.cout > 8 || g.cin > 20) return 0;  // halo tile of 5 channel quads = 163 KB smem
  const int kz = g.ndim == 3 ? 3 : 1;
  for (int i = 0; i < 3; ++i) {
    const int k = (i == 0) ? kz : 3, pd = (i == 0 && kz == 1) ? 0 : 1;
    if (g.k[i] != k || g.st[i] != 1 || g.pl[i] != pd || g.ph[i] != pd) return 0;
  }
  if (kz == 1) return 0;  // 2-D narrow layers are tiny; keep them on the generic kernel
  const int cin4 = (g.cin + 3) / 4;
  const int cop = g.cout <= 4 ? 4 : 8;
  int rc = 0;
#define S3_SMALL_CASE(C4, CO)                                        \
  if (cin4 == C4 && cop == CO) {                                     \
    rc = launch_small<C4, CO, 3>(g, x, w, ep, st);                   \
    return rc ? rc : 1;                                              \
  }
  S3_SMALL_CASE(1, 4) S3_SMALL_CASE(2, 4) S3_SMALL_CASE(3, 4) S3_SMALL_CASE(4, 4)
  S3_SMALL_CASE(5, 4)
  S3_SMALL_CASE(1, 8) S3_SMALL_CASE(2, 8) S3_SMALL_CASE(3, 8) S3_SMALL_CASE(4, 8)
  S3_SMALL_CASE(5, 8)
#undef S3_SMALL_CASE
  return 0;
}

}  // namespace s3
