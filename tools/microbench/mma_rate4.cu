// Microbenchmark 4: does a stream of SS-mode tcgen05.mma (M=128, K=16, bf16) lose cycles when
// consecutive MMAs change N (instruction descriptor), accumulator columns, or operand address?
// Patterns mimic the conv kernel's per-slab issue order: 6 groups of 4 k-step MMAs.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../sup3r_b200/csrc/ptx.cuh"
using namespace s3;

__global__ void __launch_bounds__(384, 1) k(int pattern, int iters, long long* cycles, int nspin = 0, int backoff = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[8];
  __shared__ uint32_t tptr;
  uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar2[i]), 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)(smem))[i] = 0;
  if (warp == 0) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = tptr;
  long long t0 = clock64();
  if (warp == 1) {
    const uint32_t id1 = make_idesc_f16(64, 1), id2 = make_idesc_f16(128, 1), id3 = make_idesc_f16(192, 1);
    const uint32_t hi_a = sdesc_hi_sw128(1280), hi_b = sdesc_hi_sw128(1024);
    const uint32_t a0 = sdesc_lo(base), b0 = sdesc_lo(base + 168 * 1024);
    for (int it = 0; it < iters; ++it) {
      const uint32_t tap = (((it % 9) / 3) * 1280u + ((it % 9) % 3) * 128u) >> 4;
      if (elect_one()) {
#pragma unroll
        for (int ip = 0; ip < 6; ++ip) {
          // kernel geometry: R = 4
          const int jlo = ip - 3 > 0 ? ip - 3 : 0;
          const int jhi = ip < 2 ? ip : 2;
          const int nblk = jhi - jlo + 1;
          uint32_t dcol = tm + 64 * (3 - (ip - jlo));
          uint32_t idn = nblk == 3 ? id3 : (nblk == 2 ? id2 : id1);
          uint32_t al = a0 + tap + ((ip * 23040u) >> 4);
          uint32_t bl = b0 + ((jlo * 8192u) >> 4);
          if (pattern == 1) { idn = id3; dcol = tm + 64 * (ip % 3); }          // same N, dcol varies
          if (pattern == 2) { dcol = tm; }                                     // N varies, same dcol
          if (pattern == 3) { idn = id3; dcol = tm; }                          // only A/B address vary
          if (pattern == 4) { idn = id3; dcol = tm; bl = b0; }                 // only A varies
          if (pattern == 5) { idn = id3; dcol = tm; bl = b0; al = a0 + tap; }  // nothing varies in slab
          if (pattern == 6) { idn = id2; dcol = tm + 128 * (ip & 1); }         // N=128 alternating 2 accs
          if (pattern == 7) { idn = id2; dcol = tm; }
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16_acc(dcol, mk_desc(al + 2 * kk, hi_a), mk_desc(bl + 2 * kk, hi_b), idn);
          if (pattern == 8) umma_commit(smem_u32(&bar2[ip]));                  // kernel + commit per group
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, nullptr, 0, 0, 0);
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
  }
  if (warp >= 2 && warp < 2 + nspin) {
    // waiting warps as in the conv kernel: poll the completion barrier
    while (!mbar_try_wait(smem_u32(&bar), 0)) { if (backoff) __nanosleep(backoff); }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int iters = 1000;
  const char* names[] = {"kernel order (N 64/128/192, 4 accs)", "N=192 always, dcol rotates", "N varies, same dcol",
                         "N=192 same dcol, A+B addr vary", "N=192 same dcol, A varies", "N=192 nothing varies",
                         "N=128, 2 accs alternate", "N=128 same acc", "kernel order + commit per group"};
  const double ideal[] = {4 * (2 * 48.7 + 2 * 64 + 2 * 96), 24 * 96, 4 * (2 * 48.7 + 2 * 64 + 2 * 96), 24 * 96, 24 * 96, 24 * 96, 24 * 64, 24 * 64,
                          4 * (2 * 48.7 + 2 * 64 + 2 * 96)};
  for (int backoff : {0, 32, 128, 512})
  for (int it_ : {0, 1, 2, 4, 9}) { int it = it_;
    k<<<148, 384, 220 * 1024>>>(0, 1000, d, it, backoff); const int itx = it; it = 1000;
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("pattern 0 spinning warps %d backoff %3d ns: %7.1f cycles/slab\n", itx, backoff, avg / it);
  }
  for (int p = 0; p < 1; ++p) {
    k<<<148, 384, 220 * 1024>>>(p, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("pattern %d %-40s: %7.1f cycles/slab (isolated-rate model %7.1f) (%s)\n", p, names[p], avg / iters, ideal[p],
           cudaGetErrorString(e));
  }
  return 0;
}
