// Microbenchmark 3: what bounds a stream of SS-mode tcgen05.mma (K = 16, bf16)?
//   single issuer vs two issuer warps (different accumulators), M = 128 vs 64, N sweep.
// cycles are per MMA instruction of the whole CTA (aggregate over issuers).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../sup3r_b200/csrc/ptx.cuh"
using namespace s3;

__device__ __forceinline__ uint32_t idesc_mn(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// issuers: 1 or 2 warps; each issues `iters` groups of 4 MMAs (kk = 0..3)
__global__ void __launch_bounds__(128, 1) k(int M, int N, int iters, int issuers, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), issuers); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)(smem))[i] = 0;
  if (warp == 0) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = tptr;
  long long t0 = clock64();
  if (warp >= 1 && warp <= issuers) {
    const uint32_t idesc = idesc_mn(M, N);
    const uint32_t hi_a = sdesc_hi_sw128(1280), hi_b = sdesc_hi_sw128(1024);
    const uint32_t a0 = sdesc_lo(base), b0 = sdesc_lo(base + 144 * 1024);
    const uint32_t dcol = tm + (warp - 1) * 256;
    for (int it = 0; it < iters; ++it) {
      const uint32_t al = a0 + (((it % 6) * 23040u + ((it / 6) % 9) * 128u) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_acc(dcol, mk_desc(al + 2 * kk, hi_a), mk_desc(b0 + 2 * kk, hi_b), idesc);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
  }
  if (warp == 1) {
    mbar_wait(smem_u32(&bar), 0, nullptr, 0, 0, 0);
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int iters = 2000;
  for (int issuers = 1; issuers <= 2; ++issuers)
    for (int M : {128, 64})
      for (int N : {16, 64, 128, 192, 256}) {
        k<<<148, 128, 220 * 1024>>>(M, N, iters, issuers, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        double per = avg / (iters * 4.0 * issuers);
        printf("issuers %d M %3d N %3d: %6.1f cycles/MMA -> %5.0f FLOP/cycle/SM (%s)\n", issuers, M,
               N, per, 2.0 * M * N * 16 / per, cudaGetErrorString(e));
      }
  return 0;
}
