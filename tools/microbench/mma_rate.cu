// Microbenchmark: tcgen05.mma issue rate (SS mode, M=128, K=16, bf16) vs N and operand layout.
// One CTA per SM, operands resident in smem (garbage data), accumulators in TMEM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../sup3r_b200/csrc/ptx.cuh"
using namespace s3;

__global__ void __launch_bounds__(128, 1) mma_rate(int N, int iters, int a_rows_stride, int mode,
                                                   long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)(smem))[i] = 0;
  if (warp == 0) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = tptr;
  if (warp == 1 && lane == 0) {
    uint32_t idesc = make_idesc_f16(N, 1);
    uint32_t a_base = base, b_base = base + 96 * 1024;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // mode 0: SW128 K-major, 4 k-slices of 32B; A window shifts by (it % 27) rows*stride
      uint32_t shift = (mode & 1) ? (uint32_t)((it % 9) * 128) : 0u;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t da = make_sdesc_sw128(a_base + shift + kk * 32, a_rows_stride, 0);
        uint64_t db = make_sdesc_sw128(b_base + kk * 32, 1024, 0);
        umma_f16(tm, da, db, idesc, 1);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 0, 0, 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int stride : {1024, 1280})
      for (int N : {16, 32, 64, 128, 256}) {
        mma_rate<<<148, 128, 200 * 1024>>>(N, iters, stride, mode, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("mode %d sbo %d N %3d: %.1f cycles/MMA(K=16)  -> %.0f FLOP/cycle/SM  (%s)\n", mode,
               stride, N, avg / (iters * 4.0), 2.0 * 128 * N * 16 / (avg / (iters * 4.0)),
               cudaGetErrorString(e));
      }
  return 0;
}
