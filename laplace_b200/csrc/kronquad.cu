// Kron GLM-predictive quadratic form of a weight-sharing (convolution / token-shared linear) layer WITHOUT the dense
// Jacobian (SURVEY App. A "Conv / token-shared layer: J_{n,c} = G_{n,c}^T A_n, then the same eigenbasis reduction";
// reference utils/matrix.py:406-461 does four dense (d_out x d_in) GEMMs per (n,c) on a materialised (B, C, P) tensor).
//
//   out[n, c, k] += sum_{i < d_out, j < d_in}  w(i,j) * Z_c[i,j] * Z_k[i,j]
//   Z_c[i,j] = sum_t Gt[i, (c, n, t)] * At[j, (n, t)]            (the layer Jacobian of (n,c) in the Kron eigenbasis)
//   w(i,j)   = 1 / (l1[i] * l2[j] + delta)      or, damped,     1 / ((l1[i] + sqrt(delta)) * (l2[j] + sqrt(delta)))
//
// Gt = Q1^T G (output-gradient rows rotated into the eigenbasis of the B factor), At = Q2^T A (unfolded input rows
// rotated into the eigenbasis of the A factor), both K-major fp32 (feature-major, sample rows contiguous), produced by
// the GEMM kernels.  Z_c -- d_out x d_in per (n,c), 94 MB per sample for ResNet-18's last blocks -- never exists in
// memory: a CTA forms the 64x64 tiles of Z_c for all C outputs of ONE sample in shared memory (C x 16.6 KB), reduces
// the C(C+1)/2 weighted pair products of the tile in registers, walks on to its next tile, and finally adds
// C(C+1)/2 numbers to out[n].  fp32 FMA throughout (the 1e-5 predictive gate).
#include "common.cuh"

namespace lpb {

template <int CMAX>
__global__ void __launch_bounds__(256) kron_conv_quadform_kernel(
    const float* __restrict__ G, int64_t ldg, int64_t g_stride_c, const float* __restrict__ A, int64_t lda, int d_out, int d_in,
    int T, int C, const float* __restrict__ l1, const float* __restrict__ l2, float delta, int damping, int tiles_j,
    int jt_per_block, float* __restrict__ out) {
  constexpr int TM = 64, TN = 64, TK = 16, PLD = TN + 1;
  constexpr int NACC = CMAX * (CMAX + 1) / 2;
  extern __shared__ float smem[];
  float* Ps = smem;                                  // [CMAX][TM][PLD]
  float* Gs = Ps + CMAX * TM * PLD;                  // [TK][TM + 4]
  float* As = Gs + TK * (TM + 4);                    // [TK][TN + 4]
  float* red = As + TK * (TN + 4);                   // [NACC]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n = blockIdx.z, i0 = blockIdx.y * TM;
  const int jt_beg = blockIdx.x * jt_per_block, jt_end = min(tiles_j, jt_beg + jt_per_block);
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const float sd = damping ? sqrtf(delta) : 0.f;
  float acc[NACC];
#pragma unroll
  for (int a = 0; a < NACC; ++a) acc[a] = 0.f;

  for (int jt = jt_beg; jt < jt_end; ++jt) {
    const int j0 = jt * TN;
    // ---- phase 1: Z_c tile for every output c ----
    for (int c = 0; c < C; ++c) {
      float p[4][4] = {};
      const float* Gc = G + (int64_t)c * g_stride_c + (int64_t)n * T;
      const float* An = A + (int64_t)n * T;
      for (int t0 = 0; t0 < T; t0 += TK) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int t = t0 + lk + u;
          const int i = i0 + lrow, j = j0 + lrow;
          Gs[(lk + u) * (TM + 4) + lrow] = (i < d_out && t < T) ? Gc[(int64_t)i * ldg + t] : 0.f;
          As[(lk + u) * (TN + 4) + lrow] = (j < d_in && t < T) ? An[(int64_t)j * lda + t] : 0.f;
        }
        __syncthreads();
        const int kmax = min(TK, T - t0);
        for (int kk = 0; kk < kmax; ++kk) {
          const float4 a = *reinterpret_cast<const float4*>(&Gs[kk * (TM + 4) + ty * 4]);
          const float4 b = *reinterpret_cast<const float4*>(&As[kk * (TN + 4) + tx * 4]);
          const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) p[u][v] = fmaf(av[u], bv[v], p[u][v]);
        }
        __syncthreads();
      }
      float* Pc = Ps + c * TM * PLD;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) Pc[(ty * 4 + u) * PLD + tx * 4 + v] = p[u][v];
    }
    __syncthreads();
    // ---- phase 2: weighted pair products of the tile ----
    for (int e = tid; e < TM * TN; e += 256) {
      const int il = e >> 6, jl = e & 63;
      const int i = i0 + il, j = j0 + jl;
      if (i >= d_out || j >= d_in) continue;
      const float w = damping ? 1.f / ((l1[i] + sd) * (l2[j] + sd)) : 1.f / fmaf(l1[i], l2[j], delta);
      float pv[CMAX];
#pragma unroll
      for (int c = 0; c < CMAX; ++c) pv[c] = c < C ? Ps[c * TM * PLD + il * PLD + jl] : 0.f;
      int a = 0;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        const float wc = w * pv[c];
#pragma unroll
        for (int k = 0; k <= c; ++k) { acc[a] = fmaf(wc, pv[k], acc[a]); ++a; }
      }
    }
    __syncthreads();
  }
  // ---- block reduction of the C(C+1)/2 accumulators, one atomic per pair ----
  for (int a = tid; a < NACC; a += 256) red[a] = 0.f;
  __syncthreads();
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    float s = acc[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) atomicAdd(&red[a], s);
  }
  __syncthreads();
  for (int a = tid; a < NACC; a += 256) {
    // a = c (c + 1) / 2 + k, k <= c
    int c = 0;
    while ((c + 1) * (c + 2) / 2 <= a) ++c;
    const int k = a - c * (c + 1) / 2;
    if (c < C) {
      float* o = out + (int64_t)n * C * C;
      atomicAdd(&o[c * C + k], red[a]);
      if (k != c) atomicAdd(&o[k * C + c], red[a]);
    }
  }
}

int kron_conv_quadform(const float* Gt, int64_t ldg, int64_t g_stride_c, const float* At, int64_t lda, int d_out, int d_in,
                       int T, int Nn, int C, const float* l1, const float* l2, float delta, int damping, float* out,
                       cudaStream_t st) {
  if (Nn == 0 || C == 0 || d_out == 0 || d_in == 0) return 0;
  LPB_REQUIRE(C <= 12, "kron_conv_quadform: at most 12 outputs per launch (C=%d); use the dense route", C);
  LPB_REQUIRE(T > 0 && Nn <= 65535, "kron_conv_quadform: bad extents (T=%d, Nn=%d)", T, Nn);
  const int tiles_i = (int)ceil_div(d_out, 64), tiles_j = (int)ceil_div(d_in, 64);
  LPB_REQUIRE(tiles_i <= 65535, "kron_conv_quadform: d_out too large");
  // enough CTAs to fill the machine twice; every extra split of the j tiles costs one more reduction per (n, i-tile)
  int64_t zs = ceil_div((int64_t)sm_count() * 2, (int64_t)Nn * tiles_i);
  zs = imax(1, imin(zs, tiles_j));
  const int jt_per_block = (int)ceil_div(tiles_j, zs);
  zs = ceil_div(tiles_j, jt_per_block);
  dim3 grid((unsigned)zs, (unsigned)tiles_i, (unsigned)Nn);
  auto smem_bytes = [](int cmax) {
    return (size_t)(cmax * 64 * 65 + 2 * 16 * 68 + cmax * (cmax + 1) / 2) * sizeof(float);
  };
#define LPB_KQ(CM)                                                                                                        \
  do {                                                                                                                    \
    static bool attr = false;                                                                                             \
    if (!attr) {                                                                                                          \
      if (check_cuda(cudaFuncSetAttribute(kron_conv_quadform_kernel<CM>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                          (int)smem_bytes(CM)), "kron_conv_quadform attr"))                               \
        return 1;                                                                                                         \
      attr = true;                                                                                                        \
    }                                                                                                                     \
    kron_conv_quadform_kernel<CM><<<grid, 256, smem_bytes(CM), st>>>(Gt, ldg, g_stride_c, At, lda, d_out, d_in, T, C, l1, \
                                                                      l2, delta, damping, tiles_j, jt_per_block, out);    \
  } while (0)
  if (C <= 4) LPB_KQ(4);
  else if (C <= 10) LPB_KQ(10);
  else LPB_KQ(12);
#undef LPB_KQ
  LPB_CHECK_LAUNCH("kron_conv_quadform");
  return 0;
}

}  // namespace lpb
