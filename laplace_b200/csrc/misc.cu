// Small HBM-/latency-bound kernels of the curvature path: per-sample contractions for layers
// with weight sharing, Jacobian writers, batched quadratic-form reductions, last-layer GGN
// block expansion.  All fp32.
#include "common.cuh"

namespace lpb {

// ---------------------------------------------------------------------------------------
// Per-sample layer Jacobian for weight-sharing layers (conv / token-shared linear):
//     P_q[i, j] = sum_t G[i, q*T + t] * A[j, n(q)*T + t],   q = c*Nn + n
// MODE 0 (diag GGN/EF, K4):   out[i, j] += scale * sum_q P_q[i,j]^2
// MODE 1 (Jacobian writer, K8): Js[n, c, off + i*d_in + j] = P_q[i,j]
// G: [d_out, Q*T] K-major fp32, A: [d_in, Nn*T] K-major fp32.
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) shared_weight_contract_kernel(
    const float* __restrict__ G, int64_t ldg, const float* __restrict__ A, int64_t lda, int d_out, int d_in, int T,
    int Nn, int Q, int q_per_block, float scale, float* __restrict__ out, int64_t out_ld, int64_t js_stride_n,
    int64_t js_stride_c) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float Gs[TK][TM + 4];
  __shared__ float As[TK][TN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  const int qbeg = blockIdx.z * q_per_block, qend = min(Q, qbeg + q_per_block);
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  float acc[4][4] = {};
  for (int q = qbeg; q < qend; ++q) {
    const int n = q % Nn;
    float p[4][4] = {};
    for (int t0 = 0; t0 < T; t0 += TK) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = t0 + lk + u;
        const int i = i0 + lrow, j = j0 + lrow;
        Gs[lk + u][lrow] = (i < d_out && t < T) ? G[(int64_t)i * ldg + (int64_t)q * T + t] : 0.f;
        As[lk + u][lrow] = (j < d_in && t < T) ? A[(int64_t)j * lda + (int64_t)n * T + t] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&Gs[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&As[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) p[u][v] = fmaf(av[u], bv[v], p[u][v]);
      }
      __syncthreads();
    }
    if (MODE == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(p[u][v], p[u][v], acc[u][v]);
    } else {
      const int c = q / Nn;
      float* dst = out + (int64_t)n * js_stride_n + (int64_t)c * js_stride_c;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + ty * 4 + u;
        if (i >= d_out) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int j = j0 + tx * 4 + v;
          if (j < d_in) dst[(int64_t)i * d_in + j] = p[u][v];
        }
      }
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + ty * 4 + u;
      if (i >= d_out) continue;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int j = j0 + tx * 4 + v;
        if (j < d_in) atomicAdd(&out[(int64_t)i * out_ld + j], scale * acc[u][v]);
      }
    }
  }
}

int shared_weight_contract(int mode, const float* G, int64_t ldg, const float* A, int64_t lda, int d_out, int d_in, int T,
                           int Nn, int ncols, float scale, float* out, int64_t out_ld, int64_t js_stride_n,
                           int64_t js_stride_c, cudaStream_t st) {
  const int Q = Nn * ncols;
  if (Q == 0 || d_out == 0 || d_in == 0) return 0;
  const int tiles_i = (int)ceil_div(d_out, 64), tiles_j = (int)ceil_div(d_in, 64);
  LPB_REQUIRE(tiles_i <= 65535, "shared_weight_contract: d_out too large");
  int64_t zsplit = 1;
  if (mode == 0) {
    zsplit = ceil_div((int64_t)sm_count() * 4, (int64_t)tiles_i * tiles_j);
    zsplit = imax(1, imin(zsplit, Q));
  } else {
    zsplit = imin(Q, 65535);
  }
  const int q_per_block = (int)ceil_div(Q, zsplit);
  zsplit = ceil_div(Q, q_per_block);
  LPB_REQUIRE(zsplit <= 65535, "shared_weight_contract: batch too large for one launch (Q=%d)", Q);
  dim3 grid(tiles_j, tiles_i, (unsigned)zsplit);
  if (mode == 0)
    shared_weight_contract_kernel<0><<<grid, 256, 0, st>>>(G, ldg, A, lda, d_out, d_in, T, Nn, Q, q_per_block, scale,
                                                           out, out_ld, 0, 0);
  else
    shared_weight_contract_kernel<1><<<grid, 256, 0, st>>>(G, ldg, A, lda, d_out, d_in, T, Nn, Q, q_per_block, scale,
                                                           out, out_ld, js_stride_n, js_stride_c);
  LPB_CHECK_LAUNCH("shared_weight_contract");
  return 0;
}

// ---------------------------------------------------------------------------------------
// Jacobian writer for layers without weight sharing (K8):
//   Js[n, c, off_w + i*d_in + j] = g[c, n, i] * a[n, j];   Js[n, c, off_b + i] = g[c, n, i]
// g: [C, Nn, d_out] fp32, a: [Nn, d_in] fp32.  Coalesced fp32 writes along j.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) jac_linear_write_kernel(const float* __restrict__ g, const float* __restrict__ a,
                                                                int Nn, int C, int d_out, int d_in,
                                                                float* __restrict__ Js, int64_t stride_n,
                                                                int64_t stride_c, int64_t off_w, int64_t off_b) {
  const int n = blockIdx.x, c = blockIdx.y;
  const float* gr = g + ((int64_t)c * Nn + n) * d_out;
  const float* ar = a + (int64_t)n * d_in;
  float* dst = Js + (int64_t)n * stride_n + (int64_t)c * stride_c;
  if (off_w >= 0) {
    const int64_t total = (int64_t)d_out * d_in;
    for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
      const int i = e / d_in, j = e - (int64_t)i * d_in;
      dst[off_w + e] = gr[i] * ar[j];
    }
  }
  if (off_b >= 0)
    for (int i = threadIdx.x; i < d_out; i += blockDim.x) dst[off_b + i] = gr[i];
}

int jac_linear_write(const float* g, const float* a, int Nn, int C, int d_out, int d_in, float* Js, int64_t stride_n,
                     int64_t stride_c, int64_t off_w, int64_t off_b, cudaStream_t st) {
  if (Nn == 0 || C == 0) return 0;
  LPB_REQUIRE(C <= 65535, "jac_linear_write: too many outputs");
  dim3 grid(Nn, C);
  jac_linear_write_kernel<<<grid, 256, 0, st>>>(g, a, Nn, C, d_out, d_in, Js, stride_n, stride_c, off_w, off_b);
  LPB_CHECK_LAUNCH("jac_linear_write");
  return 0;
}

// Last-layer Jacobian  J_n = [I_C (x) phi_n^T , I_C]  (reference curvature.py:157-165)
__global__ void __launch_bounds__(256) ll_jacobian_write_kernel(const float* __restrict__ phi, int Nn, int C, int D,
                                                                 int has_bias, float* __restrict__ Js) {
  const int n = blockIdx.x, c = blockIdx.y;
  const int64_t P = (int64_t)C * D + (has_bias ? C : 0);
  float* dst = Js + ((int64_t)n * C + c) * P;
  const float* pr = phi + (int64_t)n * D;
  for (int64_t e = threadIdx.x; e < P; e += blockDim.x) {
    float v = 0.f;
    if (e < (int64_t)C * D) {
      const int cc = e / D;
      if (cc == c) v = pr[e - (int64_t)cc * D];
    } else if (e - (int64_t)C * D == c) {
      v = 1.f;
    }
    dst[e] = v;
  }
}

int ll_jacobian_write(const float* phi, int Nn, int C, int D, int has_bias, float* Js, cudaStream_t st) {
  if (Nn == 0 || C == 0) return 0;
  dim3 grid(Nn, C);
  ll_jacobian_write_kernel<<<grid, 256, 0, st>>>(phi, Nn, C, D, has_bias, Js);
  LPB_CHECK_LAUNCH("ll_jacobian_write");
  return 0;
}

// ---------------------------------------------------------------------------------------
// Batched weighted pair reduction (K6/K7 epilogues):
//   out[n, c, k] (+)= sum_i X[n, c, i] * Z[n, k, i] * (m ? m[n*m_stride + i] : 1)
// one CTA per n, one warp per (c,k) pair round-robin, warp-shuffle reduction.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) batched_pair_dot_kernel(const float* __restrict__ X, const float* __restrict__ Z,
                                                                const float* __restrict__ m, int64_t m_stride, int CX,
                                                                int CZ, int d, int64_t x_stride_n, int64_t x_stride_c,
                                                                int64_t z_stride_n, int64_t z_stride_c,
                                                                int accumulate, float* __restrict__ out) {
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* Xn = X + (int64_t)n * x_stride_n;
  const float* Zn = Z + (int64_t)n * z_stride_n;
  const float* mn = m ? m + (int64_t)n * m_stride : nullptr;
  for (int p = warp; p < CX * CZ; p += nwarps) {
    const int c = p / CZ, k = p - c * CZ;
    const float* xr = Xn + (int64_t)c * x_stride_c;
    const float* zr = Zn + (int64_t)k * z_stride_c;
    float s = 0.f;
    for (int i = lane; i < d; i += 32) s = fmaf(xr[i] * zr[i], mn ? mn[i] : 1.f, s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      float* o = out + ((int64_t)n * CX + c) * CZ + k;
      *o = accumulate ? *o + s : s;
    }
  }
}

int batched_pair_dot(const float* X, const float* Z, const float* m, int64_t m_stride, int Nn, int CX, int CZ, int d,
                     int64_t x_stride_n, int64_t x_stride_c, int64_t z_stride_n, int64_t z_stride_c, int accumulate,
                     float* out, cudaStream_t st) {
  if (Nn == 0 || CX == 0 || CZ == 0) return 0;
  batched_pair_dot_kernel<<<Nn, 256, 0, st>>>(X, Z, m, m_stride, CX, CZ, d, x_stride_n, x_stride_c, z_stride_n,
                                              z_stride_c, accumulate, out);
  LPB_CHECK_LAUNCH("batched_pair_dot");
  return 0;
}

// ---------------------------------------------------------------------------------------
// Last-layer full GGN block expansion (K3, structured):
//   G[dt, pair(c<=k), et] = sum_n L_n[c,k] phit_n[dt] phit_n[et]   (phit = [phi; 1] if bias)
//   H[idx(c,dt), idx(k,et)] (+)= G   with idx(c,d<D) = c*D + d, idx(c,D) = C*D + c
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ll_ggn_expand_kernel(const float* __restrict__ G, int C, int D, int has_bias,
                                                             int accumulate, float* __restrict__ H) {
  const int Dt = D + (has_bias ? 1 : 0);
  const int64_t P = (int64_t)C * D + (has_bias ? C : 0);
  const int npairs = C * (C + 1) / 2;
  const int64_t total = P * P;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / P, s = e - r * P;
    int c, dt, k, et;
    if (r < (int64_t)C * D) { c = r / D; dt = r - (int64_t)c * D; } else { c = r - (int64_t)C * D; dt = D; }
    if (s < (int64_t)C * D) { k = s / D; et = s - (int64_t)k * D; } else { k = s - (int64_t)C * D; et = D; }
    int cc = c, kk = k, a = dt, b = et;
    if (cc > kk) { cc = k; kk = c; a = et; b = dt; }
    const int pair = cc * C - cc * (cc - 1) / 2 + (kk - cc);
    const float v = G[((int64_t)a * npairs + pair) * Dt + b];
    H[e] = accumulate ? H[e] + v : v;
  }
}

int ll_ggn_expand(const float* G, int C, int D, int has_bias, int accumulate, float* H, cudaStream_t st) {
  const int64_t P = (int64_t)C * D + (has_bias ? C : 0);
  if (P == 0) return 0;
  const int64_t total = P * P;
  const int blocks = (int)imin(ceil_div(total, 256), (int64_t)sm_count() * 16);
  ll_ggn_expand_kernel<<<blocks, 256, 0, st>>>(G, C, D, has_bias, accumulate, H);
  LPB_CHECK_LAUNCH("ll_ggn_expand");
  return 0;
}

// Sigma (P x P, last-layer ordering) -> Sg[(pair(c,k) all C*C ordered), et, dt] so that the
// predictive GEMM  Y[n, (c,k,et)] = sum_dt phit[n,dt] * Sigma[idx(c,dt), idx(k,et)]  is an NT GEMM.
__global__ void __launch_bounds__(256) ll_sigma_gather_kernel(const float* __restrict__ S, int C, int D, int has_bias,
                                                               float* __restrict__ Sg) {
  const int Dt = D + (has_bias ? 1 : 0);
  const int64_t P = (int64_t)C * D + (has_bias ? C : 0);
  const int64_t total = (int64_t)C * C * Dt * Dt;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int dt = e % Dt;
    const int et = (e / Dt) % Dt;
    const int ck = e / ((int64_t)Dt * Dt);
    const int c = ck / C, k = ck - c * C;
    const int64_t r = dt < D ? (int64_t)c * D + dt : (int64_t)C * D + c;
    const int64_t s = et < D ? (int64_t)k * D + et : (int64_t)C * D + k;
    Sg[e] = S[r * P + s];
  }
}

int ll_sigma_gather(const float* S, int C, int D, int has_bias, float* Sg, cudaStream_t st) {
  const int Dt = D + (has_bias ? 1 : 0);
  const int64_t total = (int64_t)C * C * Dt * Dt;
  if (total == 0) return 0;
  const int blocks = (int)imin(ceil_div(total, 256), (int64_t)sm_count() * 16);
  ll_sigma_gather_kernel<<<blocks, 256, 0, st>>>(S, C, D, has_bias, Sg);
  LPB_CHECK_LAUNCH("ll_sigma_gather");
  return 0;
}

}  // namespace lpb
