// Batched symmetric eigendecomposition (K5) for small Kronecker factors: parallel cyclic two-sided
// Jacobi, one CTA per matrix, matrix and eigenvector accumulators resident in shared memory.
// Semantics of reference symeig (utils/utils.py:193-228): upper triangle is read, eigenvalues
// ascending, clamped at 0, NaN -> 0.  fp32 throughout; Jacobi is backward stable to ~n*eps.
#include "common.cuh"

namespace lpb {

__global__ void __launch_bounds__(512) eigh_jacobi_kernel(const float* __restrict__ Ain, int n, float* __restrict__ evals,
                                                          float* __restrict__ Qout, int max_sweeps) {
  extern __shared__ float sm[];
  const int np = (n + 1) & ~1;  // even player count for the round-robin schedule
  const int ld = np + 1;
  float* A = sm;                 // np x ld
  float* V = A + np * ld;        // np x ld
  float* cs = V + np * ld;       // 2 * (np/2)
  int* pq = reinterpret_cast<int*>(cs + np);  // 2 * (np/2)
  __shared__ int rotated;
  __shared__ float abs_floor;
  __shared__ int order[256];
  const float* Ab = Ain + (int64_t)blockIdx.x * n * n;
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int e = tid; e < np * np; e += nt) {
    const int i = e / np, j = e - i * np;
    float a = 0.f;
    if (i < n && j < n) a = (i <= j) ? Ab[(int64_t)i * n + j] : Ab[(int64_t)j * n + i];  // UPLO = 'U'
    A[i * ld + j] = a;
    V[i * ld + j] = (i == j) ? 1.f : 0.f;
  }
  __syncthreads();
  if (tid == 0) {
    // off-diagonal entries below 2e-8 * max|diag| are at the rounding level of the rotations that touch their rows
    // (6e-8 of the large entries): chasing them with the relative criterion alone never terminates for the
    // (large, tiny) eigenvalue pairs of a rank-deficient PSD factor and burns all max_sweeps
    float dmax = 0.f;
    for (int i = 0; i < n; ++i) dmax = fmaxf(dmax, fabsf(A[i * ld + i]));
    abs_floor = 2e-8f * dmax;
  }
  __syncthreads();

  const int half = np / 2;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int step = 0; step < np - 1; ++step) {
      // phase 1: rotation angles for the np/2 disjoint pairs of this round
      if (tid < half) {
        int p, q;
        if (tid == 0) { p = np - 1; q = step; }
        else { p = (step + tid) % (np - 1); q = (step - tid + (np - 1)) % (np - 1); }
        if (p > q) { int t = p; p = q; q = t; }
        const float app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
        float c = 1.f, s = 0.f;
        if (fabsf(apq) > 1e-30f && fabsf(apq) > abs_floor && fabsf(apq) > 6e-8f * sqrtf(fabsf(app * aqq))) {
          // angles in double: fp32 angles (even correctly rounded division / square root) leave c^2 + s^2 off by ~1e-7
          // per rotation, which compounds over the ~2n rotations that touch every element per sweep (r02 measured:
          // reconstruction error 1.6e-5 instead of < 1e-5 at n = 63, and no speed-up -- the step time is the shared-memory
          // sweep over A and V, not this arithmetic)
          const double tau = ((double)aqq - (double)app) / (2.0 * (double)apq);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          const double cd = 1.0 / sqrt(1.0 + t * t);
          c = (float)cd;
          s = (float)(t * cd);
          rotated = 1;
        }
        cs[2 * tid] = c; cs[2 * tid + 1] = s;
        pq[2 * tid] = p; pq[2 * tid + 1] = q;
      }
      __syncthreads();
      // phase 2: columns  A <- A J,  V <- V J
      for (int e = tid; e < half * np; e += nt) {
        const int k = e / np, i = e - k * np;
        const float c = cs[2 * k], s = cs[2 * k + 1];
        if (s == 0.f) continue;
        const int p = pq[2 * k], q = pq[2 * k + 1];
        const float aip = A[i * ld + p], aiq = A[i * ld + q];
        A[i * ld + p] = c * aip - s * aiq;
        A[i * ld + q] = s * aip + c * aiq;
        const float vip = V[i * ld + p], viq = V[i * ld + q];
        V[i * ld + p] = c * vip - s * viq;
        V[i * ld + q] = s * vip + c * viq;
      }
      __syncthreads();
      // phase 3: rows  A <- J^T A
      for (int e = tid; e < half * np; e += nt) {
        const int k = e / np, j = e - k * np;
        const float c = cs[2 * k], s = cs[2 * k + 1];
        if (s == 0.f) continue;
        const int p = pq[2 * k], q = pq[2 * k + 1];
        const float apj = A[p * ld + j], aqj = A[q * ld + j];
        A[p * ld + j] = c * apj - s * aqj;
        A[q * ld + j] = s * apj + c * aqj;
      }
      __syncthreads();
    }
    if (!rotated) break;
    __syncthreads();
  }

  // rank eigenvalues ascending (n <= 256: O(n^2) ranking by one thread per entry)
  if (tid < n) {
    const float v = A[tid * ld + tid];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const float w = A[j * ld + j];
      r += (w < v) || (w == v && j < tid);
    }
    order[r] = tid;
  }
  __syncthreads();
  float* ev = evals + (int64_t)blockIdx.x * n;
  float* Qb = Qout + (int64_t)blockIdx.x * n * n;
  for (int r = tid; r < n; r += nt) {
    float v = A[order[r] * ld + order[r]];
    v = (v != v) ? 0.f : fmaxf(v, 0.f);
    ev[r] = v;
  }
  for (int e = tid; e < n * n; e += nt) {
    const int i = e / n, r = e - i * n;
    float v = V[i * ld + order[r]];
    Qb[e] = (v != v) ? 0.f : v;
  }
}

int eigh_jacobi(const float* A, int batch, int n, float* evals, float* Q, int max_sweeps, cudaStream_t st) {
  if (batch == 0 || n == 0) return 0;
  LPB_REQUIRE(n <= 128, "eigh_jacobi: n=%d exceeds the shared-memory Jacobi limit (128)", n);
  const int np = (n + 1) & ~1;
  const size_t smem = (size_t)(2 * np * (np + 1) + 2 * np) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    if (check_cuda(cudaFuncSetAttribute(eigh_jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                   "eigh_jacobi attr"))
      return 1;
    attr_set = true;
  }
  eigh_jacobi_kernel<<<batch, np >= 96 ? 512 : 256, smem, st>>>(A, n, evals, Q, max_sweeps > 0 ? max_sweeps : 30);
  LPB_CHECK_LAUNCH("eigh_jacobi");
  return 0;
}

}  // namespace lpb
