// Pack kernels: bring layer inputs / output-gradients into the K-major staging layout the
// GEMM kernels consume:  dst[j * ldk + k0 + k]  (feature j, sample-row k; k contiguous).
//
// These are the HBM-bound front end of every factor contraction:
// algorithmic bytes per element = 4 (fp32 read) + 2 / 4 (bf16 / bf16 hi+lo write) or + 4 (fp32).
#include "common.cuh"

namespace lpb {

// ---------------------------------------------------------------------------------------
// rows:   src [rows=K, cols=d] fp32 row-major  ->  dst[j, k0 + k]      (tiled transpose)
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) pack_rows_t_kernel(const float* __restrict__ src, int64_t rows, int64_t cols,
                                                           int64_t ld_src, const float* __restrict__ row_scale,
                                                           float scale, int flags, void* hi, void* lo, int64_t ldk,
                                                           int64_t k0) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t kb = (int64_t)blockIdx.x * 32, jb = (int64_t)blockIdx.y * 32;
  // blockIdx.z = replica: replica z uses row_scale[z*rows + k] and writes rows [z*cols, (z+1)*cols)
  const int64_t rep = blockIdx.z;
  if (row_scale) row_scale += rep * rows;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t k = kb + ty + 8 * i, j = jb + tx;
    float v = 0.f;
    if (k < rows && j < cols) {
      v = src[k * ld_src + j] * scale;
      if (flags & PACK_SQUARE) v = v * v;
      if (row_scale) v *= row_scale[k];
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t j = jb + ty + 8 * i, k = kb + tx;
    if (j < cols && k < rows) store_packed<KIND>(hi, lo, (rep * cols + j) * ldk + k0 + k, tile[tx][ty + 8 * i]);
  }
}

// ---------------------------------------------------------------------------------------
// conv2d patches: x [N, C, H, W] -> dst[(ci, kh, kw), k0 + (n, oh, ow)]   (im2col, K-major)
// mirrors the row order of F.unfold / einconv patches used by KFAC-expand
// ---------------------------------------------------------------------------------------

template <int KIND>
__global__ void __launch_bounds__(256) pack_conv2d_t_kernel(const float* __restrict__ x, ConvGeom g, float scale,
                                                             int flags, void* hi, void* lo, int64_t ldk, int64_t k0) {
  const int64_t K = (int64_t)g.N * g.OH * g.OW;
  const int r = blockIdx.y;  // (ci, kh, kw)
  const int kw = r % g.KW, kh = (r / g.KW) % g.KH, ci = r / (g.KW * g.KH);
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < K; k += (int64_t)gridDim.x * blockDim.x) {
    const int ow = k % g.OW;
    const int oh = (k / g.OW) % g.OH;
    const int n = k / ((int64_t)g.OW * g.OH);
    const int ih = oh * g.SH - g.PH + kh * g.DH, iw = ow * g.SW - g.PW + kw * g.DW;
    float v = 0.f;
    if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) {
      v = x[(((int64_t)n * g.C + ci) * g.H + ih) * g.W + iw] * scale;
      if (flags & PACK_SQUARE) v = v * v;
    }
    store_packed<KIND>(hi, lo, (int64_t)r * ldk + k0 + k, v);
  }
}

// KFAC-reduce input rows for conv: mean over output positions of the unfolded patches
//   dst[(ci,kh,kw), k0 + n] = 1/(OH*OW) sum_{oh,ow} x[n, ci, ih, iw]
template <int KIND>
__global__ void __launch_bounds__(256) pack_conv2d_mean_t_kernel(const float* __restrict__ x, ConvGeom g, float scale,
                                                                  void* hi, void* lo, int64_t ldk, int64_t k0) {
  const int r = blockIdx.y;
  const int kw = r % g.KW, kh = (r / g.KW) % g.KH, ci = r / (g.KW * g.KH);
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= g.N) return;
  float s = 0.f;
  for (int oh = 0; oh < g.OH; ++oh) {
    const int ih = oh * g.SH - g.PH + kh * g.DH;
    if (ih < 0 || ih >= g.H) continue;
    for (int ow = 0; ow < g.OW; ++ow) {
      const int iw = ow * g.SW - g.PW + kw * g.DW;
      if (iw >= 0 && iw < g.W) s += x[(((int64_t)n * g.C + ci) * g.H + ih) * g.W + iw];
    }
  }
  store_packed<KIND>(hi, lo, (int64_t)r * ldk + k0 + n, s * scale / (float)(g.OH * g.OW));
}

// ---------------------------------------------------------------------------------------
// channel-major gradients: g [Nn, Cc, HW] -> dst[ch, k0 + n*HW + hw]
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) pack_nchw_t_kernel(const float* __restrict__ g, int64_t Nn, int Cc, int HW,
                                                           float scale, int flags, void* hi, void* lo, int64_t ldk,
                                                           int64_t k0) {
  const int64_t K = Nn * HW;
  const int ch = blockIdx.y;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < K; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = k / HW;
    const int hw = k - n * HW;
    float v = g[(n * Cc + ch) * HW + hw] * scale;
    if (flags & PACK_SQUARE) v = v * v;
    store_packed<KIND>(hi, lo, (int64_t)ch * ldk + k0 + k, v);
  }
}

// sum over HW (KFAC-reduce output rows): dst[ch, k0 + n] = sum_hw g[n, ch, hw]
template <int KIND>
__global__ void __launch_bounds__(256) pack_nchw_sum_t_kernel(const float* __restrict__ g, int64_t Nn, int Cc, int HW,
                                                               float scale, void* hi, void* lo, int64_t ldk,
                                                               int64_t k0) {
  const int ch = blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Nn) return;
  float s = 0.f;
  for (int hw = 0; hw < HW; ++hw) s += g[(n * Cc + ch) * HW + hw];
  store_packed<KIND>(hi, lo, (int64_t)ch * ldk + k0 + n, s * scale);
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
#define DISPATCH_KIND(kind, CALL)                           \
  switch (kind) {                                           \
    case OUT_F32: { constexpr int KIND = OUT_F32; CALL; } break;           \
    case OUT_BF16: { constexpr int KIND = OUT_BF16; CALL; } break;         \
    case OUT_BF16_HILO: { constexpr int KIND = OUT_BF16_HILO; CALL; } break; \
    default: set_error("bad out_kind %d", kind); return 1;  \
  }

int pack_rows_t(const float* src, int64_t rows, int64_t cols, int64_t ld_src, const float* row_scale, int nrep,
                float scale, int flags, void* hi, void* lo, int kind, int64_t ldk, int64_t k0, cudaStream_t st) {
  if (rows == 0 || cols == 0 || nrep <= 0) return 0;
  LPB_REQUIRE(nrep == 1 || row_scale != nullptr, "pack_rows_t: replicas need per-replica row scales");
  LPB_REQUIRE(nrep <= 65535, "pack_rows_t: too many replicas");
  LPB_REQUIRE(kind != OUT_BF16_HILO || lo != nullptr, "pack_rows_t: hi+lo output needs a lo buffer");
  dim3 grid((unsigned)ceil_div(rows, 32), (unsigned)ceil_div(cols, 32), (unsigned)nrep);
  LPB_REQUIRE(grid.y <= 65535, "pack_rows_t: too many columns (%lld)", (long long)cols);
  DISPATCH_KIND(kind, (pack_rows_t_kernel<KIND><<<grid, 256, 0, st>>>(src, rows, cols, ld_src, row_scale, scale, flags,
                                                                      hi, lo, ldk, k0)));
  LPB_CHECK_LAUNCH("pack_rows_t");
  return 0;
}

int pack_conv2d_t(const float* x, const ConvGeom& g, float scale, int flags, int reduce_mean, void* hi, void* lo,
                  int kind, int64_t ldk, int64_t k0, cudaStream_t st) {
  const int rows = g.C * g.KH * g.KW;
  LPB_REQUIRE(rows <= 65535, "pack_conv2d_t: too many patch rows (%d)", rows);
  LPB_REQUIRE(kind != OUT_BF16_HILO || lo != nullptr, "pack_conv2d_t: hi+lo output needs a lo buffer");
  if (g.N == 0) return 0;
  if (reduce_mean) {
    dim3 grid((unsigned)ceil_div(g.N, 256), rows);
    DISPATCH_KIND(kind, (pack_conv2d_mean_t_kernel<KIND><<<grid, 256, 0, st>>>(x, g, scale, hi, lo, ldk, k0)));
  } else {
    const int64_t K = (int64_t)g.N * g.OH * g.OW;
    dim3 grid((unsigned)imin(ceil_div(K, 256), 4096), rows);
    DISPATCH_KIND(kind, (pack_conv2d_t_kernel<KIND><<<grid, 256, 0, st>>>(x, g, scale, flags, hi, lo, ldk, k0)));
  }
  LPB_CHECK_LAUNCH("pack_conv2d_t");
  return 0;
}

int pack_nchw_t(const float* gp, int64_t Nn, int Cc, int HW, float scale, int flags, int reduce_sum, void* hi, void* lo,
                int kind, int64_t ldk, int64_t k0, cudaStream_t st) {
  LPB_REQUIRE(Cc <= 65535, "pack_nchw_t: too many channels (%d)", Cc);
  LPB_REQUIRE(kind != OUT_BF16_HILO || lo != nullptr, "pack_nchw_t: hi+lo output needs a lo buffer");
  if (Nn == 0) return 0;
  if (reduce_sum) {
    dim3 grid((unsigned)ceil_div(Nn, 256), Cc);
    DISPATCH_KIND(kind, (pack_nchw_sum_t_kernel<KIND><<<grid, 256, 0, st>>>(gp, Nn, Cc, HW, scale, hi, lo, ldk, k0)));
  } else {
    const int64_t K = Nn * HW;
    dim3 grid((unsigned)imin(ceil_div(K, 256), 4096), Cc);
    DISPATCH_KIND(kind, (pack_nchw_t_kernel<KIND><<<grid, 256, 0, st>>>(gp, Nn, Cc, HW, scale, flags, hi, lo, ldk, k0)));
  }
  LPB_CHECK_LAUNCH("pack_nchw_t");
  return 0;
}

}  // namespace lpb
