// Weight gradient of the 64-channel 3x3x3 stride-1 reflect-pad-1 convolutions on tcgen05:
//     dW[dz][dy][dx][ci][co] = sum over voxels v of  X_pad[v + (dz,dy,dx)][ci] * dZ[v][co]
// (tape.gradient w.r.t. the conv kernels, sup3r/models/abstract.py:1190-1238).
//
// GEMM view: the reduction dimension K is the VOXEL index, so both operands are "MN-major" UMMA
// operands: a voxel's 64 channels (128 B of fp16) are one row of a SWIZZLE_128B box exactly as
// TMA lands it, and 8 consecutive x voxels are one 8 x 64 swizzle atom:
//   A = dZ tile   (M = co;  16 y x 8 x voxels of one plane, box 16 KB)
//   B = X window  (N = 3 dx taps x 64 ci: the three dx-shifted windows of the padded input plane
//                  are N blocks 128 B apart -- LBO = one voxel row; the dy shift is the start row)
//   one tcgen05.mma (M = 128, N = 192, K = 16 voxels = two x-runs of adjacent y rows) per k-step.
// (M = 128 is issued with rows 64-127 reading the next atom -- finite garbage nobody loads back:
// the M = 128 accumulator layout, lane = row, is the one the other kernels of this library use.)
// TMEM holds two (dy) accumulators of 192 columns, so a CTA owns one dz and either dy in {0, 1}
// (type A) or dy = 2 (type B) and a contiguous slice of the voxel tiles (split K); type A gets
// twice as many slices as type B so that all 144 CTAs do equal work.  Partial sums go to a
// workspace [job][384][64] (coalesced on co) and a second kernel reduces them in a fixed order
// (deterministic) into the keras layout (kz, ky, kx, cin, cout).
// Pipeline: warp 0 = TMA producer (X plane box 18 x 10 voxels + dZ box per tile), warp 1 = MMA
// issuer, both warps drain the accumulators at the end.
#include <cstring>

#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace s3 {

constexpr uint32_t kWgG = 16384;          // dZ box: 16 y x 8 x voxels x 128 B
constexpr uint32_t kWgX = 23552;          // X box: 18 x 10 x 128 = 23040, padded to 1 KiB
constexpr uint32_t kWgXBytes = 23040;
constexpr uint32_t kWgStage = kWgG + kWgX;
constexpr int kWgStages = 5;
constexpr int kWgCols = 384;              // workspace columns per job (2 dy x 3 dx x 64 ci)

struct WgradParams {
  int n, Z, Y, X, nyb, nxb, n_tiles;
  int g_off, g_pitch, x_pitch;
  int n_a, n_b;
  float* ws;
};

// instruction descriptor: f32 accumulate, fp16 operands, both MN-major, M = 128
__host__ __device__ inline uint32_t wg_idesc(uint32_t n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// MN-major SWIZZLE_128B shared-memory descriptor: LBO = stride between 64-element MN blocks,
// SBO = stride between 8-row K atoms (cute mma_traits_sm100.hpp: ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)))
__device__ __forceinline__ uint64_t wg_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

__device__ __forceinline__ void wg_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  for (uint32_t n = 0; !mbar_try_wait(bar, parity); ++n)
    if (n > (1u << 26)) __trap();
}

__global__ void __launch_bounds__(64, 1)
conv_wgrad_umma_kernel(const __grid_constant__ CUtensorMap tm_x,
                       const __grid_constant__ CUtensorMap tm_g, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + kWgStages * kWgStage;
  auto bar = [&](int i) { return bar_base + 8u * i; };   // full[S] | empty[S] | accfull | tmemptr
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- job: dz, the dy combos, the tile slice
  int j = blockIdx.x, dz, ncomb, dy0, slice, nsl;
  if (j < 3 * p.n_a) { dz = j / p.n_a; slice = j - dz * p.n_a; nsl = p.n_a; ncomb = 2; dy0 = 0; }
  else { j -= 3 * p.n_a; dz = j / p.n_b; slice = j - dz * p.n_b; nsl = p.n_b; ncomb = 1; dy0 = 2; }
  const int t0 = (int)(((long long)p.n_tiles * slice) / nsl);
  const int t1 = (int)(((long long)p.n_tiles * (slice + 1)) / nsl);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * kWgStages; ++i) mbar_init(bar(i), 1);
    mbar_init(bar(2 * kWgStages), 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_g);
  }
  if (warp == 1) {
    tmem_alloc(bar(2 * kWgStages + 1), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bar(2 * kWgStages + 1)));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int s = 0, ph = 0;
    for (int t = t0; t < t1; ++t) {
      const int xb = t % p.nxb;
      int r = t / p.nxb;
      const int yb = r % p.nyb;
      r /= p.nyb;
      const int z = r % p.Z, b = r / p.Z;
      wg_wait(bar(kWgStages + s), (uint32_t)(ph ^ 1));
      if (elect_one()) {
        const uint32_t st = base + s * kWgStage;
        mbar_expect_tx(bar(s), kWgG + kWgXBytes);
        tma_load_4d(st, &tm_g, bar(s), 0, xb * 8 + p.g_off, yb * 16 + p.g_off,
                    b * p.g_pitch + z + p.g_off);
        tma_load_4d(st + kWgG, &tm_x, bar(s), 0, xb * 8, yb * 16, b * p.x_pitch + z + dz);
      }
      __syncwarp();
      if (++s == kWgStages) { s = 0; ph ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = wg_idesc(192u);
    int s = 0, ph = 0;
    for (int t = t0; t < t1; ++t) {
      wg_wait(bar(s), (uint32_t)ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = base + s * kWgStage;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c < ncomb) {
            const int dy = dy0 + c;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t a = wg_desc(st + kk * 2048u, 1024u, 1024u);
              const uint64_t bd = wg_desc(st + kWgG + (uint32_t)((2 * kk + dy) * 10) * 128u, 128u,
                                          1280u);
              umma_f16(tmem_base + 192u * c, a, bd, idesc, (t > t0 || kk > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(bar(kWgStages + s));
        if (t == t1 - 1) umma_commit(bar(2 * kWgStages));
      }
      __syncwarp();
      if (++s == kWgStages) { s = 0; ph ^= 1; }
    }
  }

  // ---------------------------------------------------------------------- drain (lanes 0-63)
  float* out = p.ws + (size_t)blockIdx.x * kWgCols * 64;
  const int co = warp * 32 + lane;
  if (t1 > t0) {
    wg_wait(bar(2 * kWgStages), 0u);
    tc_fence_after();
    const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int col = 0; col < 192 * ncomb; col += 16) {
      uint32_t v[16];
      tmem_ld16(ta + col, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) out[(size_t)(col + i) * 64 + co] = __uint_as_float(v[i]);
    }
  } else {
    for (int col = 0; col < 192 * ncomb; ++col) out[(size_t)col * 64 + co] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// dW[(dz,dy,dx)][ci][co] = sum over the K slices of the job that owns (dz, dy), fixed order
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int cin,
                                    int n_a, int n_b, float scale) {
  const int total = 27 * cin * 64;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int co = idx & 63;
    int r = idx >> 6;
    const int ci = r % cin;
    r /= cin;
    const int dx = r % 3, dy = (r / 3) % 3, dz = r / 9;
    const bool a = dy < 2;
    const int job0 = a ? dz * n_a : 3 * n_a + dz * n_b;
    const int nsl = a ? n_a : n_b;
    const int col = (a ? dy : 0) * 192 + dx * 64 + ci;
    float acc = 0.f;
    for (int s = 0; s < nsl; ++s) acc += ws[((size_t)(job0 + s) * kWgCols + col) * 64 + co];
    dw[idx] = acc * scale;
  }
}

typedef CUresult (*WgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                               const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int wg_encode(CUtensorMap* tm, const void* ptr, const uint64_t* dims, const uint32_t* box) {
  static WgEncodeFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) ==
            cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WgEncodeFn>(q);
  }
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return S3_ERR_CUDA;
  }
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4];
  uint64_t stride = 2;
  for (int i = 0; i < 4; ++i) {
    gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1;
    stride *= dims[i];
    if (i < 3) gstr[i] = stride;
  }
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d (wgrad)", (int)r);
    return S3_ERR_CUDA;
  }
  return S3_OK;
}

static void wg_slices(int n_tiles, int* n_a, int* n_b) {
  int nb = sm_count() / 9;
  if (nb < 1) nb = 1;
  int na = 2 * nb;
  if (na > n_tiles) na = n_tiles;
  if (nb > n_tiles) nb = n_tiles;
  *n_a = na; *n_b = nb;
}

}  // namespace s3

using namespace s3;

extern "C" size_t s3_conv_wgrad_umma_ws_bytes(int n, int z, int y, int x) {
  int na, nb;
  const long long tiles = (long long)n * z * ((y + 15) / 16) * ((x + 7) / 8);
  wg_slices((int)(tiles > (1 << 30) ? (1 << 30) : tiles), &na, &nb);
  return (size_t)(3 * (na + nb)) * kWgCols * 64 * sizeof(float);
}

extern "C" int s3_conv_wgrad_umma(const void* x_hi, const void* g_hi, int g_halo, int n, int z,
                                  int y, int x, int cin, float scale, float* dw, void* ws,
                                  size_t ws_bytes, s3_stream stream) {
  S3_REQUIRE(x_hi && g_hi && dw && ws, "s3_conv_wgrad_umma: null argument");
  S3_REQUIRE(n >= 1 && z >= 2 && y >= 2 && x >= 2, "s3_conv_wgrad_umma: extents must be >= 2");
  S3_REQUIRE(g_halo == 1 || g_halo == 2, "s3_conv_wgrad_umma: dZ halo must be 1 or 2");
  S3_REQUIRE(cin >= 1 && cin <= 64, "s3_conv_wgrad_umma: cin must be <= 64");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.n = n; p.Z = z; p.Y = y; p.X = x;
  p.nyb = (y + 15) / 16; p.nxb = (x + 7) / 8;
  const long long tiles = (long long)n * z * p.nyb * p.nxb;
  S3_REQUIRE(tiles < (1 << 30), "s3_conv_wgrad_umma: too many tiles");
  p.n_tiles = (int)tiles;
  p.g_off = g_halo; p.g_pitch = z + 2 * g_halo; p.x_pitch = z + 2;
  wg_slices(p.n_tiles, &p.n_a, &p.n_b);
  const size_t need = (size_t)(3 * (p.n_a + p.n_b)) * kWgCols * 64 * sizeof(float);
  S3_REQUIRE(ws_bytes >= need, "s3_conv_wgrad_umma: workspace of %zu bytes needed, got %zu", need,
             ws_bytes);
  p.ws = static_cast<float*>(ws);
  CUtensorMap tm_x, tm_g;
  const uint64_t xd[4] = {64, (uint64_t)x + 2, (uint64_t)y + 2, (uint64_t)n * (z + 2)};
  const uint32_t xbox[4] = {64, 10, 18, 1};
  const uint64_t gd[4] = {64, (uint64_t)x + 2 * g_halo, (uint64_t)y + 2 * g_halo,
                          (uint64_t)n * (z + 2 * g_halo)};
  const uint32_t gbox[4] = {64, 8, 16, 1};
  int rc;
  if ((rc = wg_encode(&tm_x, x_hi, xd, xbox))) return rc;
  if ((rc = wg_encode(&tm_g, g_hi, gd, gbox))) return rc;
  const uint32_t smem = kWgStages * kWgStage + 1024u + 1024u;
  static bool attr = false;
  if (!attr) {
    S3_CUDA(cudaFuncSetAttribute(conv_wgrad_umma_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int ctas = 3 * (p.n_a + p.n_b);
  conv_wgrad_umma_kernel<<<ctas, 64, smem, as_stream(stream)>>>(tm_x, tm_g, p);
  S3_CUDA(cudaGetLastError());
  const int total = 27 * cin * 64;
  wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(p.ws, dw, cin, p.n_a,
                                                                          p.n_b, scale);
  S3_LAUNCH_CHECK("conv_wgrad_umma");
  return S3_OK;
}
