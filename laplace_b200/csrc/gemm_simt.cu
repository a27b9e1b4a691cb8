// Exact-fp32 SIMT contraction  D[M,N] += alpha * A[M,K] * B[N,K]^T  on K-major operands.
//
// This is the full-precision path of the factor contractions (K1/K2/K3 in SURVEY 2.3): used
// for operands too small / too irregular for the tcgen05 path and wherever the caller asks for
// fp32 products (predictive variances at 1e-5).  Split-K over gridDim.z, fp32 atomics into the
// accumulated factor buffer; SYM computes only tiles with tn >= tm and mirrors them.
#include "common.cuh"

namespace lpb {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool SYM>
__global__ void __launch_bounds__(256) gemm_nt_f32_kernel(const float* __restrict__ A, int64_t lda,
                                                           const float* __restrict__ B, int64_t ldb, int M, int N,
                                                           int64_t K, float alpha, float* __restrict__ D, int64_t ldd,
                                                           int64_t k_per_split) {
  const int tn = blockIdx.x, tm = blockIdx.y;
  if (SYM && tn < tm) return;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = tm * BM, n0 = tn * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = min(K, kbeg + k_per_split);
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const bool vecA = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool vecB = (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  float acc[4][4] = {};
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int m = m0 + lrow;
      const int64_t k = k0 + lk;
      if (m < M) {
        const float* p = A + (int64_t)m * lda + k;
        if (vecA && k + 3 < kend) {
          float4 q = *reinterpret_cast<const float4*>(p);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (k + i < kend) v[i] = p[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[lk + i][lrow] = v[i];
    }
    {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int n = n0 + lrow;
      const int64_t k = k0 + lk;
      if (n < N) {
        const float* p = B + (int64_t)n * ldb + k;
        if (vecB && k + 3 < kend) {
          float4 q = *reinterpret_cast<const float4*>(p);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (k + i < kend) v[i] = p[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[lk + i][lrow] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const float v = alpha * acc[i][j];
      atomicAdd(&D[(int64_t)m * ldd + n], v);
      if (SYM && tn != tm) atomicAdd(&D[(int64_t)n * ldd + m], v);
    }
  }
}

int gemm_nt_f32(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t N, int64_t K, float alpha,
                int accumulate, float* D, int64_t ldd, int symmetric, cudaStream_t st) {
  LPB_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm_nt_f32: negative extent");
  LPB_REQUIRE(!symmetric || M == N, "gemm_nt_f32: symmetric needs M == N");
  if (M == 0 || N == 0) return 0;
  if (!accumulate) {
    if (check_cuda(cudaMemset2DAsync(D, ldd * sizeof(float), 0, N * sizeof(float), M, st), "gemm_nt_f32 memset"))
      return 1;
  }
  if (K == 0) return 0;
  const int tiles_m = (int)ceil_div(M, BM), tiles_n = (int)ceil_div(N, BN);
  LPB_REQUIRE(tiles_m <= 65535, "gemm_nt_f32: M too large");
  const int64_t tiles = symmetric ? (int64_t)tiles_m * (tiles_m + 1) / 2 : (int64_t)tiles_m * tiles_n;
  // split K so that ~4 waves of CTAs exist, each with at least 8 k-steps
  const int64_t ksteps = ceil_div(K, BK);
  int64_t splits = ceil_div((int64_t)sm_count() * 8, tiles);
  splits = imax(1, imin(splits, ceil_div(ksteps, 8)));
  splits = imin(splits, 65535);
  const int64_t k_per_split = ceil_div(ksteps, splits) * BK;
  splits = ceil_div(K, k_per_split);
  dim3 grid(tiles_n, tiles_m, (unsigned)splits);
  if (symmetric)
    gemm_nt_f32_kernel<true><<<grid, 256, 0, st>>>(A, lda, B, ldb, (int)M, (int)N, K, alpha, D, ldd, k_per_split);
  else
    gemm_nt_f32_kernel<false><<<grid, 256, 0, st>>>(A, lda, B, ldb, (int)M, (int)N, K, alpha, D, ldd, k_per_split);
  LPB_CHECK_LAUNCH("gemm_nt_f32");
  return 0;
}

}  // namespace lpb
