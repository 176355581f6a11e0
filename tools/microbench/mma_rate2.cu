// Microbenchmark 2: tcgen05.mma (SS, M=128, K=16, bf16) cycles/MMA under interference:
//   mode 0: MMA alone;  mode 1: 4 warps spin on tcgen05.ld of other TMEM columns;
//   mode 2: 4 warps stream 16-byte st.shared into a scratch region;  mode 3: both.
// A window rotates over 6 "planes" and 9 shifts like the conv kernel; SBO = 1280.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../sup3r_b200/csrc/ptx.cuh"
using namespace s3;

__global__ void __launch_bounds__(192, 1) k(int N, int iters, int mode, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  __shared__ volatile int stop;
  uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); stop = 0; }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)(smem))[i] = 0;
  if (warp == 0) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = tptr;
  if (warp == 1) {
    uint32_t idesc = make_idesc_f16(N, 1);
    uint32_t hi_a = sdesc_hi_sw128(1280), hi_b = sdesc_hi_sw128(1024);
    uint32_t a0 = sdesc_lo(base), b0 = sdesc_lo(base + 144 * 1024);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t al = a0 + (((it % 6) * 23040u + ((it / 6) % 9) * 128u) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_acc(tm + (it & 1) * 64, mk_desc(al + 2 * kk, hi_a), mk_desc(b0 + 2 * kk, hi_b), idesc);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, nullptr, 0, 0, 0);
    long long t1 = clock64();
    if (lane == 0) { cycles[blockIdx.x] = t1 - t0; stop = 1; }
  } else if (warp >= 2) {
    int q = warp & 3;
    uint32_t v[16];
    float acc = 0.f;
    uint4* scratch = reinterpret_cast<uint4*>(smem + 176 * 1024) + threadIdx.x;
    while (!stop) {
      if (mode & 1) {
        tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + 256, v);
        tmem_ld_wait();
        acc += __uint_as_float(v[3]);
      }
      if (mode & 2) {
        for (int r = 0; r < 8; ++r) scratch[r * 192] = make_uint4(r, r, r, r);
      }
      if (mode == 0) __nanosleep(100);
    }
    if (acc == 123.f) cycles[200] = 1;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 256 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  int iters = 3000;
  for (int mode = 0; mode < 4; ++mode)
    for (int N : {64, 128, 192, 208, 256}) {
      k<<<148, 192, 220 * 1024>>>(N, iters, mode, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      printf("mode %d N %3d: %.1f cycles/MMA  (%s)\n", mode, N, avg / (iters * 4.0), cudaGetErrorString(e));
    }
  return 0;
}
