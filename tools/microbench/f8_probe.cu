// Probe for the fp16 + fp8-correction scheme ("fp16c"):
//  (1) numerics: D = A16 * B16^T (kind::f16, fp16 operands, K = 64) followed by
//      D += A8 * B8^T (kind::f8f6f4, e4m3 operands, K = 128) into the SAME TMEM accumulator,
//      both operand pairs K-major SWIZZLE_128B tiles with 128-byte rows; checked against the host;
//  (2) rate: the ring kernel's per-slab issue pattern (N = 64 / 128 / 192 groups of 4 k-steps)
//      with kind::f16 (K = 16) and kind::f8f6f4 (K = 32) MMAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 f8_probe.cu -o f8_probe.bin
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include "../../sup3r_b200/csrc/ptx.cuh"
using namespace s3;

__device__ __forceinline__ void umma_f8_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.eq.u32 p, 1, 1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}

// tiles in global memory, already in the swizzled shared-memory image (128-byte rows)
__global__ void __launch_bounds__(128, 1) numerics(const uint8_t* a16, const uint8_t* b16,
                                                   const uint8_t* a8, const uint8_t* b8, float* d,
                                                   int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  uint8_t* s = smem + (base - smem_u32(smem));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A16 @0 (16 KB), B16 @16K (8 KB), A8 @24K (16 KB), B8 @40K (8 KB)
  for (int i = threadIdx.x; i < 16384 / 16; i += 128) {
    reinterpret_cast<uint4*>(s)[i] = reinterpret_cast<const uint4*>(a16)[i];
    reinterpret_cast<uint4*>(s + 24576)[i] = reinterpret_cast<const uint4*>(a8)[i];
  }
  for (int i = threadIdx.x; i < 8192 / 16; i += 128) {
    reinterpret_cast<uint4*>(s + 16384)[i] = reinterpret_cast<const uint4*>(b16)[i];
    reinterpret_cast<uint4*>(s + 40960)[i] = reinterpret_cast<const uint4*>(b8)[i];
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tptr), 64); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tptr;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t hi = sdesc_hi_sw128(1024);
      const uint32_t id = make_idesc_f16(64, 0);   // fp16 x fp16 -> f32 / e4m3 x e4m3 -> f32
      if (mode & 1) {
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t ad = mk_desc(sdesc_lo(base) + 2 * kk, hi);
          const uint64_t bd = mk_desc(sdesc_lo(base + 16384) + 2 * kk, hi);
          if (kk == 0) umma_f16_new(tm, ad, bd, id); else umma_f16_acc(tm, ad, bd, id);
        }
      }
      if (mode & 2) {
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t ad = mk_desc(sdesc_lo(base + 24576) + 2 * kk, hi);
          const uint64_t bd = mk_desc(sdesc_lo(base + 40960) + 2 * kk, hi);
          if (kk == 0 && !(mode & 1)) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\t"
                "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
                ::"r"(tm), "l"(ad), "l"(bd), "r"(id) : "memory");
          } else {
            umma_f8_acc(tm, ad, bd, id);
          }
        }
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0, nullptr, 0, 0, 0);
  tc_fence_after();
  for (int c = 0; c < 4; ++c) {
    uint32_t v[16];
    tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + 16 * c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d[(warp * 32 + lane) * 64 + 16 * c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

__global__ void __launch_bounds__(128, 1) rate(int kind, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)(smem))[i] = 0;
  if (warp == 0) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tptr;
  const long long t0 = clock64();
  if (warp == 1) {
    const uint32_t id1 = make_idesc_f16(64, 0), id2 = make_idesc_f16(128, 0), id3 = make_idesc_f16(192, 0);
    const uint32_t hi_a = sdesc_hi_sw128(1280), hi_b = sdesc_hi_sw128(1024);
    const uint32_t a0 = sdesc_lo(base), b0 = sdesc_lo(base + 168 * 1024);
    for (int it = 0; it < iters; ++it) {
      const uint32_t tap = (((it % 9) / 3) * 1280u + ((it % 9) % 3) * 128u) >> 4;
      const bool use8 = kind == 1 || (kind == 2 && ((it / 9) & 1)) || (kind == 3 && (it & 1));
      if (elect_one()) {
        if (use8) {
#pragma unroll
          for (int ip = 0; ip < 6; ++ip) {
            const int jlo = ip - 3 > 0 ? ip - 3 : 0;
            const int jhi = ip < 2 ? ip : 2;
            const int nblk = jhi - jlo + 1;
            const uint32_t dcol = tm + 64 * (3 - (ip - jlo));
            const uint32_t idn = nblk == 3 ? id3 : (nblk == 2 ? id2 : id1);
            const uint32_t al = a0 + tap + ((ip * 23040u) >> 4);
            const uint32_t bl = b0 + ((jlo * 8192u) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f8_acc(dcol, mk_desc(al + 2 * kk, hi_a), mk_desc(bl + 2 * kk, hi_b), idn);
          }
        } else {
#pragma unroll
          for (int ip = 0; ip < 6; ++ip) {
            const int jlo = ip - 3 > 0 ? ip - 3 : 0;
            const int jhi = ip < 2 ? ip : 2;
            const int nblk = jhi - jlo + 1;
            const uint32_t dcol = tm + 64 * (3 - (ip - jlo));
            const uint32_t idn = nblk == 3 ? id3 : (nblk == 2 ? id2 : id1);
            const uint32_t al = a0 + tap + ((ip * 23040u) >> 4);
            const uint32_t bl = b0 + ((jlo * 8192u) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16_acc(dcol, mk_desc(al + 2 * kk, hi_a), mk_desc(bl + 2 * kk, hi_b), idn);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, nullptr, 0, 0, 0);
    const long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// host-side swizzled image: row r (128 B), 16-byte chunk k -> chunk (k ^ (r & 7))
static void put(std::vector<uint8_t>& img, int r, int byte, uint8_t v) {
  const int k = byte >> 4, o = byte & 15;
  img[(size_t)r * 128 + ((k ^ (r & 7)) << 4) + o] = v;
}

int main() {
  const int M = 128, N = 64;
  std::vector<float> A16(M * 64), B16(N * 64), A8(M * 128), B8(N * 128);
  std::vector<uint8_t> ia16(M * 128), ib16(N * 128), ia8(M * 128), ib8(N * 128);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (int r = 0; r < M; ++r)
    for (int k = 0; k < 64; ++k) {
      __half h = __float2half(rnd());
      A16[r * 64 + k] = __half2float(h);
      uint16_t u; memcpy(&u, &h, 2);
      put(ia16, r, 2 * k, u & 0xff); put(ia16, r, 2 * k + 1, u >> 8);
    }
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < 64; ++k) {
      __half h = __float2half(rnd() * 1000.f);   // scaled weights (large fp16 values)
      B16[r * 64 + k] = __half2float(h);
      uint16_t u; memcpy(&u, &h, 2);
      put(ib16, r, 2 * k, u & 0xff); put(ib16, r, 2 * k + 1, u >> 8);
    }
  for (int r = 0; r < M; ++r)
    for (int k = 0; k < 128; ++k) {
      __nv_fp8_e4m3 q(rnd() * 4.f);
      A8[r * 128 + k] = (float)q;
      put(ia8, r, k, q.__x);
    }
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < 128; ++k) {
      __nv_fp8_e4m3 q(rnd() * 8.f);
      B8[r * 128 + k] = (float)q;
      put(ib8, r, k, q.__x);
    }
  uint8_t *da16, *db16, *da8, *db8; float* dd;
  cudaMalloc(&da16, ia16.size()); cudaMalloc(&db16, ib16.size());
  cudaMalloc(&da8, ia8.size()); cudaMalloc(&db8, ib8.size()); cudaMalloc(&dd, M * N * 4);
  cudaMemcpy(da16, ia16.data(), ia16.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db16, ib16.data(), ib16.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(da8, ia8.data(), ia8.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db8, ib8.data(), ib8.size(), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(numerics, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 1; mode <= 3; ++mode) {
    numerics<<<1, 128, 64 * 1024>>>(da16, db16, da8, db8, dd, mode);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(M * N);
    cudaMemcpy(D.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int r = 0; r < M; ++r)
      for (int c = 0; c < N; ++c) {
        double ref = 0;
        if (mode & 1) for (int k = 0; k < 64; ++k) ref += (double)A16[r * 64 + k] * B16[c * 64 + k];
        if (mode & 2) for (int k = 0; k < 128; ++k) ref += (double)A8[r * 128 + k] * B8[c * 128 + k];
        maxerr = fmax(maxerr, fabs(ref - D[r * N + c]));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("numerics mode %d (1 = f16, 2 = f8, 3 = f16 then f8 into one accumulator): max err %.3e, max |ref| %.3e (%s)\n",
           mode, maxerr, maxref, cudaGetErrorString(e));
  }
  long long* dc; cudaMalloc(&dc, 148 * 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const char* names[] = {"kind::f16 K=16", "kind::f8f6f4 K=32", "alternating every 9 slabs", "alternating every slab"};
  for (int kind = 0; kind < 4; ++kind) {
    rate<<<148, 128, 220 * 1024>>>(kind, 1800, dc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("rate %-28s: %7.1f cycles/slab (%s)\n", names[kind], avg / 1800, cudaGetErrorString(e));
  }
  return 0;
}
