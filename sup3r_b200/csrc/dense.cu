// Dense layers of the discriminator tail (keras Dense: y = act(x @ W + b)), their adjoints,
// and a column-sum reduction shared with the conv bias gradient.
// M (= batch) is tiny and K is large (flattened conv features), so the op is bound by one
// streaming read of W: the grid is split over N tiles x K slices to cover all SMs.
#include "common.cuh"

namespace s3 {

// C(m, n) (+)= sum_k A(m, k) * B(k, n) with A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
             int M, int N, int K, long long sam, long long sak, long long sbk, long long sbn,
             int k_per_split, size_t split_stride) {
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    for (int e = tid; e < BM * BK; e += NT) {
      int mm, kk;
      if (sak == 1) { mm = e / BK; kk = e % BK; } else { kk = e / BM; mm = e % BM; }
      float v = 0.f;
      if (m0 + mm < M && k0 + kk < k_end) v = __ldg(A + (m0 + mm) * sam + (k0 + kk) * sak);
      As[kk][mm] = v;
    }
    for (int e = tid; e < BK * BN; e += NT) {
      int kk, nn;
      if (sbn == 1) { kk = e / BN; nn = e % BN; } else { nn = e / BK; kk = e % BK; }
      float v = 0.f;
      if (n0 + nn < N && k0 + kk < k_end) v = __ldg(B + (k0 + kk) * sbk + (n0 + nn) * sbn);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int m = m0 + ty * TM + i, n = n0 + tx * TN + j;
      if (m < M && n < N) {
        // (split K: every split writes its own partial C, reduced in a fixed order afterwards)
        C[(size_t)blockIdx.z * split_stride + (size_t)m * N + n] = acc[i][j];
      }
    }
}

__global__ void bias_act_kernel(float* __restrict__ y, const float* __restrict__ b, size_t total,
                                int n, int act, float alpha) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    float v = y[i] + (b ? b[i % n] : 0.f);
    y[i] = apply_act(v, act, alpha);
  }
}

// Column sums out[c] = sum_r x[r, c], DETERMINISTIC: every block writes its partial sums to
// part[block][c] (out itself when there is one block), a second kernel adds the blocks in a
// fixed order.  32 channels x 8 row lanes per block.
__global__ void colsum_kernel(const float* __restrict__ x, long long rows, int cols,
                              float* __restrict__ part) {
  const int c = blockIdx.y * 32 + (threadIdx.x & 31);
  const int lane = threadIdx.x >> 5, lanes = blockDim.x >> 5;
  float s = 0.f;
  if (c < cols)
    for (long long r = (long long)blockIdx.x * lanes + lane; r < rows;
         r += (long long)gridDim.x * lanes)
      s += x[r * cols + c];
  __shared__ float sh[8][33];
  sh[lane][threadIdx.x & 31] = s;
  __syncthreads();
  if (lane == 0 && c < cols) {
    for (int i = 1; i < lanes; ++i) s += sh[i][threadIdx.x & 31];
    part[(size_t)blockIdx.x * cols + c] = s;
  }
}

// 32 columns x 8 lanes per block: lane l adds blocks l, l + 8, ... in order, the 8 lane sums are
// added in lane order -- a fixed summation tree, whatever the grid
__global__ void colsum_reduce_kernel(const float* __restrict__ part, int nblk, int cols,
                                     float* __restrict__ out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int lane = threadIdx.x >> 5;
  float s = 0.f;
  if (c < cols)
    for (int b = lane; b < nblk; b += 8) s += part[(size_t)b * cols + c];
  __shared__ float sh[8][33];
  sh[lane][threadIdx.x & 31] = s;
  __syncthreads();
  if (lane == 0 && c < cols) {
    for (int i = 1; i < 8; ++i) s += sh[i][threadIdx.x & 31];
    out[c] = s;
  }
}

// partial-sum scratch of the column sums: grown on demand, one per device (the only allocation
// the library makes after s3_init; stream-ordered use, freed with the context)
static size_t gemm_scratch_offset() { return (size_t)1 << 18; }   // floats reserved for colsum

static float* colsum_scratch(size_t floats) {
  static float* buf[16] = {nullptr};
  static size_t cap[16] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  if (cap[dev] < floats) {
    if (buf[dev]) {
      cudaDeviceSynchronize();
      cudaFree(buf[dev]);
    }
    size_t want = floats < (1u << 18) ? (1u << 18) : floats;
    if (cudaMalloc(&buf[dev], want * sizeof(float)) != cudaSuccess) {
      buf[dev] = nullptr; cap[dev] = 0;
      return nullptr;
    }
    cap[dev] = want;
  }
  return buf[dev];
}

int launch_colsum(const float* x, long long rows, int cols, float* out, cudaStream_t st) {
  if (rows == 0) {
    S3_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
    return S3_OK;
  }
  long long bx = (rows + 7) / 8;
  long long cap = (long long)sm_count() * 8 / ((cols + 31) / 32) + 1;
  if (bx > cap) bx = cap;
  if (rows <= 4096) bx = 1;     // small reductions (dense layers): one block, no second stage
  dim3 grid((unsigned)bx, (cols + 31) / 32);
  if (bx == 1) {
    colsum_kernel<<<grid, 256, 0, st>>>(x, rows, cols, out);
  } else {
    float* part = colsum_scratch((size_t)bx * cols);
    S3_REQUIRE(part != nullptr, "colsum: cannot allocate %lld x %d partial sums", bx, cols);
    colsum_kernel<<<grid, 256, 0, st>>>(x, rows, cols, part);
    colsum_reduce_kernel<<<(cols + 31) / 32, 256, 0, st>>>(part, (int)bx, cols, out);
  }
  S3_CUDA(cudaGetLastError());
  return S3_OK;
}

__global__ void split_reduce_kernel(const float* __restrict__ part, int splits, size_t n,
                                    float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(size_t)k * n + i];
    out[i] = s;
  }
}

static int launch_gemm(const float* A, const float* B, float* C, int M, int N, int K,
                       long long sam, long long sak, long long sbk, long long sbn,
                       cudaStream_t st) {
  const int tiles = ((N + 63) / 64) * ((M + 63) / 64);
  int splits = (2 * sm_count() + tiles - 1) / tiles;
  int max_splits = (K + 63) / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int kps = (K + splits - 1) / splits;
  kps = (kps + 15) / 16 * 16;
  splits = (K + kps - 1) / kps;
  const size_t n = (size_t)M * N;
  float* dst = C;
  if (splits > 1) {   // deterministic split K: partial products, then a fixed-order reduction
    dst = colsum_scratch((size_t)splits * n + gemm_scratch_offset());
    S3_REQUIRE(dst != nullptr, "gemm: cannot allocate %d x %zu partial sums", splits, n);
    dst += gemm_scratch_offset();
  }
  dim3 grid((N + 63) / 64, (M + 63) / 64, splits);
  sgemm_kernel<64, 64, 16, 4, 4><<<grid, 256, 0, st>>>(A, B, dst, M, N, K, sam, sak, sbk, sbn, kps,
                                                      splits > 1 ? n : 0);
  S3_CUDA(cudaGetLastError());
  if (splits > 1) {
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (blocks > 2048u) blocks = 2048u;
    split_reduce_kernel<<<blocks, 256, 0, st>>>(dst, splits, n, C);
    S3_CUDA(cudaGetLastError());
  }
  return S3_OK;
}

}  // namespace s3

using namespace s3;

extern "C" int s3_dense_fwd(const float* x, const float* w, const float* b, float* y, int m, int k,
                            int n, int act, float alpha, s3_stream stream) {
  S3_REQUIRE(x && w && y && m > 0 && k > 0 && n > 0, "s3_dense_fwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  int rc = launch_gemm(x, w, y, m, n, k, k, 1, n, 1, st);
  if (rc) return rc;
  if (b || act != S3_ACT_NONE) {
    size_t total = (size_t)m * n;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (blocks > 2048) blocks = 2048;
    bias_act_kernel<<<blocks, 256, 0, st>>>(y, b, total, n, act, alpha);
    S3_LAUNCH_CHECK("bias_act");
  }
  return S3_OK;
}

extern "C" int s3_dense_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw,
                            float* db, int m, int k, int n, s3_stream stream) {
  S3_REQUIRE(dy && m > 0 && k > 0 && n > 0, "s3_dense_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  int rc;
  if (dx) {  // dx (m x k) = dy (m x n) @ W^T
    S3_REQUIRE(w, "s3_dense_bwd: dx needs w");
    rc = launch_gemm(dy, w, dx, m, k, n, n, 1, 1, n, st);
    if (rc) return rc;
  }
  if (dw) {  // dw (k x n) = x^T (k x m) @ dy (m x n)
    S3_REQUIRE(x, "s3_dense_bwd: dw needs x");
    rc = launch_gemm(x, dy, dw, k, n, m, 1, k, n, 1, st);
    if (rc) return rc;
  }
  if (db) {
    rc = launch_colsum(dy, m, n, db, st);
    if (rc) return rc;
  }
  return S3_OK;
}
