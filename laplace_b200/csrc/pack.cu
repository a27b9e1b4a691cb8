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
    case OUT_F16_HILO: { constexpr int KIND = OUT_F16_HILO; CALL; } break;   \
    default: set_error("bad out_kind %d", kind); return 1;  \
  }

int pack_rows_t(const float* src, int64_t rows, int64_t cols, int64_t ld_src, const float* row_scale, int nrep,
                float scale, int flags, void* hi, void* lo, int kind, int64_t ldk, int64_t k0, cudaStream_t st) {
  if (rows == 0 || cols == 0 || nrep <= 0) return 0;
  LPB_REQUIRE(nrep == 1 || row_scale != nullptr, "pack_rows_t: replicas need per-replica row scales");
  LPB_REQUIRE(nrep <= 65535, "pack_rows_t: too many replicas");
  LPB_REQUIRE(kind < OUT_BF16_HILO || lo != nullptr, "pack_rows_t: hi+lo output needs a lo buffer");
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
  LPB_REQUIRE(kind < OUT_BF16_HILO || lo != nullptr, "pack_conv2d_t: hi+lo output needs a lo buffer");
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
  LPB_REQUIRE(kind < OUT_BF16_HILO || lo != nullptr, "pack_nchw_t: hi+lo output needs a lo buffer");
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

// =========================================================================================
// Operand packers of the convolution engine (forward / backward-data of nn.Conv2d as GEMMs on the
// tcgen05 kernel, fp32-accurate through the bf16 hi/lo split -- see DESIGN.md section 3b).
// =========================================================================================
namespace lpb {

// patch-major im2col: x [N, C, H, W] -> dst[(n, oh, ow), (ci, kh, kw)]   (row length ld >= C*KH*KW)
template <int KIND>
__global__ void __launch_bounds__(256) pack_conv2d_rows_kernel(const float* __restrict__ x, ConvGeom g, void* hi, void* lo,
                                                                int64_t ld) {
  const int d_in = g.C * g.KH * g.KW;
  const int64_t rows = (int64_t)g.N * g.OH * g.OW;
  // one warp per output row chunk: threads run along the patch index (coalesced writes)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
    const int ow = r % g.OW;
    const int oh = (r / g.OW) % g.OH;
    const int n = r / ((int64_t)g.OW * g.OH);
    const float* xn = x + (int64_t)n * g.C * g.H * g.W;
    for (int j = lane; j < d_in; j += 32) {
      const int kw = j % g.KW, kh = (j / g.KW) % g.KH, ci = j / (g.KW * g.KH);
      const int ih = oh * g.SH - g.PH + kh * g.DH, iw = ow * g.SW - g.PW + kw * g.DW;
      float v = 0.f;
      if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) v = xn[((int64_t)ci * g.H + ih) * g.W + iw];
      store_packed<KIND>(hi, lo, r * ld + j, v);
    }
  }
}

// channel-minor rows: g [Q, Cc, HW] -> dst[(q, hw), ch]   (row length ld >= Cc); tiled transpose per q
template <int KIND>
__global__ void __launch_bounds__(256) pack_nchw_rows_kernel(const float* __restrict__ g, int Cc, int HW, void* hi, void* lo,
                                                              int64_t ld) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t q = blockIdx.z;
  const int hb = blockIdx.x * 32, cb = blockIdx.y * 32;
  const float* gq = g + q * (int64_t)Cc * HW;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ch = cb + ty + 8 * i, hw = hb + tx;
    tile[ty + 8 * i][tx] = (ch < Cc && hw < HW) ? gq[(int64_t)ch * HW + hw] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int hw = hb + ty + 8 * i, ch = cb + tx;
    if (hw < HW && ch < Cc) store_packed<KIND>(hi, lo, (q * HW + hw) * ld + ch, tile[tx][ty + 8 * i]);
  }
}

// plain cast of a [rows, cols] fp32 matrix into a K-major operand (no transpose)
template <int KIND>
__global__ void __launch_bounds__(256) pack_cast_kernel(const float* __restrict__ src, int64_t rows, int64_t cols,
                                                         int64_t ld_src, void* hi, void* lo, int64_t ld) {
  const int64_t total = rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / cols, c = e - r * cols;
    store_packed<KIND>(hi, lo, r * ld + c, src[r * ld_src + c]);
  }
}

// col2im gather: Dc [(ci,kh,kw), ldd] with columns (q, oh, ow)  ->  grad_in [Q, C, H, W]
// grid: x over (q, ih, iw) flattened (32-bit), y = ci; consecutive threads -> consecutive iw (coalesced taps)
__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ Dc, int64_t ldd, ConvGeom g,
                                                      float* __restrict__ out) {
  const unsigned HW = g.H * g.W, total = (unsigned)g.N * HW;
  const int ci = blockIdx.y;
  const float* base = Dc + (int64_t)ci * g.KH * g.KW * ldd;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const unsigned q = e / HW, r = e - q * HW;
    const int ih = r / g.W, iw = r - ih * g.W;
    const unsigned qoff = q * g.OH * g.OW;
    float acc = 0.f;
#pragma unroll 1
    for (int kh = 0; kh < g.KH; ++kh) {
      const int th = ih + g.PH - kh * g.DH;
      if (th < 0 || th % g.SH) continue;
      const int oh = th / g.SH;
      if (oh >= g.OH) continue;
      const float* row = base + (int64_t)kh * g.KW * ldd + qoff + oh * g.OW;
      for (int kw = 0; kw < g.KW; ++kw) {
        const int tw = iw + g.PW - kw * g.DW;
        if (tw < 0 || tw % g.SW) continue;
        const int ow = tw / g.SW;
        if (ow < g.OW) acc += __ldg(row + (int64_t)kw * ldd + ow);
      }
    }
    out[((int64_t)q * g.C + ci) * HW + r] = acc;
  }
}

// vectorised variant for 16-bit hi/lo outputs: one thread = 8 consecutive patch entries of one row (16 B stores)
template <int KIND>
__global__ void __launch_bounds__(256) pack_conv2d_rows_vec8_kernel(const float* __restrict__ x, ConvGeom g, void* hi,
                                                                     void* lo, int64_t ld) {
  const int d_in = g.C * g.KH * g.KW;
  const int chunks = (d_in + 7) / 8;
  const int64_t total = (int64_t)g.N * g.OH * g.OW * chunks;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / chunks;
    const int j0 = (int)(e - r * chunks) * 8;
    const int ow = r % g.OW;
    const int oh = (r / g.OW) % g.OH;
    const int64_t n = r / ((int64_t)g.OW * g.OH);
    const float* xn = x + n * g.C * g.H * g.W;
    alignas(16) unsigned short h[8], l[8];
    // decode (ci, kh, kw) of the first entry once, then step it (integer divisions dominated the first version)
    int kw = j0 % g.KW, kh = (j0 / g.KW) % g.KH, ci = j0 / (g.KW * g.KH);
    const int ih0 = oh * g.SH - g.PH, iw0 = ow * g.SW - g.PW;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = 0.f;
      if (j0 + i < d_in) {
        const int ih = ih0 + kh * g.DH, iw = iw0 + kw * g.DW;
        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) v = __ldg(xn + ((int64_t)ci * g.H + ih) * g.W + iw);
      }
      if (++kw == g.KW) { kw = 0; if (++kh == g.KH) { kh = 0; ++ci; } }
      if constexpr (KIND == OUT_F16_HILO) {
        const __half hh = __float2half_rn(v);
        h[i] = __half_as_ushort(hh);
        l[i] = __half_as_ushort(__float2half_rn(v - __half2float(hh)));
      } else {
        const __nv_bfloat16 hh = __float2bfloat16_rn(v);
        h[i] = __bfloat16_as_ushort(hh);
        l[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(hh)));
      }
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(hi) + r * ld + j0) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(lo) + r * ld + j0) = *reinterpret_cast<const uint4*>(l);
  }
}

int pack_conv2d_rows(const float* x, const ConvGeom& g, void* hi, void* lo, int kind, int64_t ld, cudaStream_t st) {
  const int64_t rows = (int64_t)g.N * g.OH * g.OW;
  if (rows == 0) return 0;
  LPB_REQUIRE(kind < OUT_BF16_HILO || lo != nullptr, "pack_conv2d_rows: hi+lo output needs a lo buffer");
  const int d_in = g.C * g.KH * g.KW;
  if ((kind == OUT_BF16_HILO || kind == OUT_F16_HILO) && ld % 8 == 0 && ld >= (int64_t)((d_in + 7) / 8) * 8 &&
      ((uintptr_t)hi % 16) == 0 && ((uintptr_t)lo % 16) == 0) {
    const int64_t total = rows * ((d_in + 7) / 8);
    const int blocks = (int)imin(ceil_div(total, 256), (int64_t)sm_count() * 64);
    if (kind == OUT_F16_HILO)
      pack_conv2d_rows_vec8_kernel<OUT_F16_HILO><<<blocks, 256, 0, st>>>(x, g, hi, lo, ld);
    else
      pack_conv2d_rows_vec8_kernel<OUT_BF16_HILO><<<blocks, 256, 0, st>>>(x, g, hi, lo, ld);
    LPB_CHECK_LAUNCH("pack_conv2d_rows_vec8");
    return 0;
  }
  const int blocks = (int)imin(ceil_div(rows, 8), (int64_t)sm_count() * 32);
  DISPATCH_KIND(kind, (pack_conv2d_rows_kernel<KIND><<<blocks, 256, 0, st>>>(x, g, hi, lo, ld)));
  LPB_CHECK_LAUNCH("pack_conv2d_rows");
  return 0;
}

int pack_nchw_rows(const float* gp, int64_t Q, int Cc, int HW, void* hi, void* lo, int kind, int64_t ld, cudaStream_t st) {
  if (Q == 0) return 0;
  LPB_REQUIRE(kind < OUT_BF16_HILO || lo != nullptr, "pack_nchw_rows: hi+lo output needs a lo buffer");
  if (HW == 1) {
    const int64_t total = Q * Cc;
    const int blocks = (int)imin(ceil_div(total, 256), (int64_t)sm_count() * 32);
    DISPATCH_KIND(kind, (pack_cast_kernel<KIND><<<blocks, 256, 0, st>>>(gp, Q, Cc, Cc, hi, lo, ld)));
  } else {
    // gridDim.z is limited to 65535: fold larger Q into several launches
    for (int64_t q0 = 0; q0 < Q; q0 += 65535) {
      const int64_t qn = imin(65535, Q - q0);
      dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)ceil_div(Cc, 32), (unsigned)qn);
      const float* src = gp + q0 * (int64_t)Cc * HW;
      void* h2 = (kind == OUT_F32) ? (void*)((float*)hi + q0 * HW * ld) : (void*)((__nv_bfloat16*)hi + q0 * HW * ld);
      void* l2 = lo ? (void*)((__nv_bfloat16*)lo + q0 * HW * ld) : nullptr;
      DISPATCH_KIND(kind, (pack_nchw_rows_kernel<KIND><<<grid, 256, 0, st>>>(src, Cc, HW, h2, l2, ld)));
    }
  }
  LPB_CHECK_LAUNCH("pack_nchw_rows");
  return 0;
}

// vectorised cast for 16-bit hi/lo outputs: 8 columns per thread (2 x 16 B loads, 2 x 16 B stores)
template <int KIND>
__global__ void __launch_bounds__(256) pack_cast_vec8_kernel(const float* __restrict__ src, int64_t rows, int64_t cols8,
                                                              int64_t ld_src, void* hi, void* lo, int64_t ld) {
  const int64_t total = rows * cols8;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / cols8, c = (e - r * cols8) * 8;
    const float4 a = *reinterpret_cast<const float4*>(src + r * ld_src + c);
    const float4 b = *reinterpret_cast<const float4*>(src + r * ld_src + c + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    alignas(16) unsigned short h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if constexpr (KIND == OUT_F16_HILO) {
        const __half hh = __float2half_rn(v[i]);
        h[i] = __half_as_ushort(hh);
        l[i] = __half_as_ushort(__float2half_rn(v[i] - __half2float(hh)));
      } else {
        const __nv_bfloat16 hh = __float2bfloat16_rn(v[i]);
        h[i] = __bfloat16_as_ushort(hh);
        l[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v[i] - __bfloat162float(hh)));
      }
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(hi) + r * ld + c) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(lo) + r * ld + c) = *reinterpret_cast<const uint4*>(l);
  }
}

// pack_cast with the two element-wise maps that precede a convolution's reverse pass fused in:
//     dst[r, c] = split( src[r, c] * scale[c] * (y[r % rows_y, c] > 0) )
// scale (frozen BatchNorm as a per-channel affine map) and y (forward output of the ReLU, shared by the rows_y-periodic
// curvature columns) are optional.  One pass over the gradient instead of three (relu_bwd, scale_channels, pack_cast),
// and the fp32 intermediates are never written.
template <int KIND, bool VEC>
__global__ void __launch_bounds__(256) pack_cast_fused_kernel(const float* __restrict__ src, int64_t rows, int64_t colsv,
                                                               int64_t ld_src, const float* __restrict__ scale,
                                                               const float* __restrict__ y, int64_t rows_y, int64_t ld_y,
                                                               void* hi, void* lo, int64_t ld) {
  constexpr int V = VEC ? 8 : 1;
  const int64_t total = rows * colsv;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / colsv, c = (e - r * colsv) * V;
    float v[V];
    if (VEC) {
      const float4 a = *reinterpret_cast<const float4*>(src + r * ld_src + c);
      const float4 b = *reinterpret_cast<const float4*>(src + r * ld_src + c + 4);
      v[0] = a.x; v[1 % V] = a.y; v[2 % V] = a.z; v[3 % V] = a.w; v[4 % V] = b.x; v[5 % V] = b.y; v[6 % V] = b.z; v[7 % V] = b.w;
    } else {
      v[0] = src[r * ld_src + c];
    }
    if (scale) {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] *= __ldg(scale + c + i);
    }
    if (y) {
      const float* yp = y + (r % rows_y) * ld_y + c;
      if (VEC) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(yp));
        const float4 b = __ldg(reinterpret_cast<const float4*>(yp + 4));
        const float m[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = m[i] > 0.f ? v[i] : 0.f;
      } else {
        v[0] = __ldg(yp) > 0.f ? v[0] : 0.f;
      }
    }
    if (VEC) {
      alignas(16) unsigned short h[8], l[8];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        if constexpr (KIND == OUT_F16_HILO) {
          const __half hh = __float2half_rn(v[i]);
          h[i] = __half_as_ushort(hh);
          l[i] = __half_as_ushort(__float2half_rn(v[i] - __half2float(hh)));
        } else {
          const __nv_bfloat16 hh = __float2bfloat16_rn(v[i]);
          h[i] = __bfloat16_as_ushort(hh);
          l[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v[i] - __bfloat162float(hh)));
        }
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(hi) + r * ld + c) = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(lo) + r * ld + c) = *reinterpret_cast<const uint4*>(l);
    } else {
      store_packed<KIND>(hi, lo, r * ld + c, v[0]);
    }
  }
}

// Same map, mask-major: one thread owns an 8-wide piece of a MASK row and walks the rows / rows_y gradient rows that share it
// (the folded curvature columns), so y is read once instead of once per column.  The row-major kernel above streams the
// columns one after the other; with a mask larger than a fraction of L2 (the stem of a ResNet-18 at B = 4096: 268 MB) every
// column re-reads it from DRAM -- r02 ncu: 594 MB read / 362 MB written per launch, i.e. y cost 0.64 of the gradient's own
// bytes instead of 0.1.  Arithmetic and rounding are those of the row-major kernel, element by element.
template <int KIND>
__global__ void __launch_bounds__(256) pack_cast_fused_maskmajor_kernel(const float* __restrict__ src, int64_t rows_y, int64_t colsv,
                                                                         int reps, int64_t ld_src, const float* __restrict__ scale,
                                                                         const float* __restrict__ y, int64_t ld_y, void* hi, void* lo,
                                                                         int64_t ld) {
  const int64_t total = rows_y * colsv;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ry = e / colsv, c = (e - ry * colsv) * 8;
    float mul[8];
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(y + ry * ld_y + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(y + ry * ld_y + c + 4));
      const float m[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) mul[i] = m[i] > 0.f ? 1.f : 0.f;
    }
    float sc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sc[i] = scale ? __ldg(scale + c + i) : 1.f;
#pragma unroll 4
    for (int rep = 0; rep < reps; ++rep) {
      const int64_t r = (int64_t)rep * rows_y + ry;
      const float4 a = *reinterpret_cast<const float4*>(src + r * ld_src + c);
      const float4 b = *reinterpret_cast<const float4*>(src + r * ld_src + c + 4);
      float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      alignas(16) unsigned short h[8], l[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (scale) v[i] *= sc[i];                      // same operation order as the row-major kernel: scale, then mask
        v[i] = mul[i] != 0.f ? v[i] : 0.f;
        if constexpr (KIND == OUT_F16_HILO) {
          const __half hh = __float2half_rn(v[i]);
          h[i] = __half_as_ushort(hh);
          l[i] = __half_as_ushort(__float2half_rn(v[i] - __half2float(hh)));
        } else {
          const __nv_bfloat16 hh = __float2bfloat16_rn(v[i]);
          h[i] = __bfloat16_as_ushort(hh);
          l[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v[i] - __bfloat162float(hh)));
        }
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(hi) + r * ld + c) = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(lo) + r * ld + c) = *reinterpret_cast<const uint4*>(l);
    }
  }
}

// Mask size (elements) from which the mask-major kernels take over: below it the mask survives in L2 between the columns
// (a pass streams 3x its size past it) and the row-major kernels expose more parallelism.  <0: never (tests, A/B timing).
static int64_t g_mask_major_min = 4 << 20;
void set_mask_major_min(int64_t n) { g_mask_major_min = n; }
int64_t mask_major_min() { return g_mask_major_min; }

int pack_cast_fused(const float* src, int64_t rows, int64_t cols, int64_t ld_src, const float* scale, const float* y,
                    int64_t rows_y, int64_t ld_y, void* hi, void* lo, int kind, int64_t ld, cudaStream_t st) {
  if (rows == 0 || cols == 0) return 0;
  LPB_REQUIRE(kind < OUT_BF16_HILO || lo != nullptr, "pack_cast_fused: hi+lo output needs a lo buffer");
  LPB_REQUIRE(y == nullptr || (rows_y > 0 && rows % rows_y == 0 && ld_y >= cols),
              "pack_cast_fused: mask rows must divide the gradient rows");
  if (y == nullptr) rows_y = 1;
  const bool vec = (kind == OUT_BF16_HILO || kind == OUT_F16_HILO) && cols % 8 == 0 && ld_src % 4 == 0 && ld % 8 == 0 &&
                   ((uintptr_t)src % 16) == 0 && ((uintptr_t)hi % 16) == 0 && ((uintptr_t)lo % 16) == 0 &&
                   (y == nullptr || (ld_y % 4 == 0 && ((uintptr_t)y % 16) == 0));
  if (vec && y != nullptr && rows / rows_y > 1 && g_mask_major_min >= 0 && rows_y * cols >= g_mask_major_min) {
    const int64_t total = rows_y * (cols / 8);
    const int blocks = (int)imin(ceil_div(total, 256), (int64_t)sm_count() * 32);
    const int reps = (int)(rows / rows_y);
    if (kind == OUT_F16_HILO)
      pack_cast_fused_maskmajor_kernel<OUT_F16_HILO><<<blocks, 256, 0, st>>>(src, rows_y, cols / 8, reps, ld_src, scale, y, ld_y, hi, lo, ld);
    else
      pack_cast_fused_maskmajor_kernel<OUT_BF16_HILO><<<blocks, 256, 0, st>>>(src, rows_y, cols / 8, reps, ld_src, scale, y, ld_y, hi, lo, ld);
  } else if (vec) {
    const int64_t total = rows * (cols / 8);
    const int blocks = (int)imin(ceil_div(total, 256), (int64_t)sm_count() * 32);
    if (kind == OUT_F16_HILO)
      pack_cast_fused_kernel<OUT_F16_HILO, true><<<blocks, 256, 0, st>>>(src, rows, cols / 8, ld_src, scale, y, rows_y, ld_y, hi, lo, ld);
    else
      pack_cast_fused_kernel<OUT_BF16_HILO, true><<<blocks, 256, 0, st>>>(src, rows, cols / 8, ld_src, scale, y, rows_y, ld_y, hi, lo, ld);
  } else {
    const int blocks = (int)imin(ceil_div(rows * cols, 256), (int64_t)sm_count() * 32);
    DISPATCH_KIND(kind, (pack_cast_fused_kernel<KIND, false><<<blocks, 256, 0, st>>>(src, rows, cols, ld_src, scale, y, rows_y,
                                                                                   ld_y, hi, lo, ld)));
  }
  LPB_CHECK_LAUNCH("pack_cast_fused");
  return 0;
}

int pack_cast(const float* src, int64_t rows, int64_t cols, int64_t ld_src, void* hi, void* lo, int kind, int64_t ld,
              cudaStream_t st) {
  if (rows == 0 || cols == 0) return 0;
  LPB_REQUIRE(kind < OUT_BF16_HILO || lo != nullptr, "pack_cast: hi+lo output needs a lo buffer");
  if ((kind == OUT_BF16_HILO || kind == OUT_F16_HILO) && cols % 8 == 0 && ld_src % 4 == 0 && ld % 8 == 0 &&
      ((uintptr_t)src % 16) == 0 && ((uintptr_t)hi % 16) == 0 && ((uintptr_t)lo % 16) == 0) {
    const int64_t total = rows * (cols / 8);
    const int blocks = (int)imin(ceil_div(total, 256), (int64_t)sm_count() * 32);
    if (kind == OUT_F16_HILO)
      pack_cast_vec8_kernel<OUT_F16_HILO><<<blocks, 256, 0, st>>>(src, rows, cols / 8, ld_src, hi, lo, ld);
    else
      pack_cast_vec8_kernel<OUT_BF16_HILO><<<blocks, 256, 0, st>>>(src, rows, cols / 8, ld_src, hi, lo, ld);
    LPB_CHECK_LAUNCH("pack_cast_vec8");
    return 0;
  }
  const int blocks = (int)imin(ceil_div(rows * cols, 256), (int64_t)sm_count() * 32);
  DISPATCH_KIND(kind, (pack_cast_kernel<KIND><<<blocks, 256, 0, st>>>(src, rows, cols, ld_src, hi, lo, ld)));
  LPB_CHECK_LAUNCH("pack_cast");
  return 0;
}

// small images (H*W <= 64): one thread per (q, ci) accumulates its whole input image in shared memory.  All reads
// of a warp are contiguous (columns (q, oh, ow) of one Dc row), no modulo / parity tests per element -- the
// scatter form, which is exact for any stride / dilation.
__global__ void __launch_bounds__(128) col2im_small_kernel(const float* __restrict__ Dc, int64_t ldd, ConvGeom g,
                                                            float* __restrict__ out) {
  extern __shared__ float acc[];  // [H*W][129] (padded: conflict-free in both phases)
  const int HW = g.H * g.W, T = g.OH * g.OW;
  const int tid = threadIdx.x, ci = blockIdx.y;
  const int64_t q0 = (int64_t)blockIdx.x * 128, q = q0 + tid;
  for (int i = 0; i < HW; ++i) acc[i * 129 + tid] = 0.f;
  if (q < g.N) {
    for (int kh = 0; kh < g.KH; ++kh)
      for (int kw = 0; kw < g.KW; ++kw) {
        const float* row = Dc + (int64_t)((ci * g.KH + kh) * g.KW + kw) * ldd + q * T;
        for (int oh = 0; oh < g.OH; ++oh) {
          const int ih = oh * g.SH - g.PH + kh * g.DH;
          if (ih < 0 || ih >= g.H) continue;
          for (int ow = 0; ow < g.OW; ++ow) {
            const int iw = ow * g.SW - g.PW + kw * g.DW;
            if (iw >= 0 && iw < g.W) acc[(ih * g.W + iw) * 129 + tid] += __ldg(row + oh * g.OW + ow);
          }
        }
      }
  }
  __syncthreads();
  // coalesced write-out: the block owns rows q0..q0+127 of channel ci
  const int nq = (int)imin(128, g.N - q0);
  for (int e = tid; e < nq * HW; e += 128) {
    const int ql = e / HW, idx = e - ql * HW;
    out[((q0 + ql) * g.C + ci) * HW + idx] = acc[idx * 129 + ql];
  }
}

// col2im gather, channels-last form: Dc [(q,oh,ow), (kh,kw,ci)] (row stride ldd)  ->  grad_in [Q, H, W, C].
// One thread per (q, h, w, 4 channels): the <= ceil(KH/SH) * ceil(KW/SW) taps that reach the pixel are read as float4
// runs that are contiguous in ci, the store is contiguous in ci -- both sides coalesced, no shared memory.
template <bool VEC>
__global__ void __launch_bounds__(256) col2im_nhwc_kernel(const float* __restrict__ Dc, int64_t ldd, ConvGeom g,
                                                          float* __restrict__ out) {
  constexpr int V = VEC ? 4 : 1;
  // 32-bit index arithmetic (the host checks N*H*W*C < 2^31); only the final addresses are 64-bit
  const uint32_t cv = (uint32_t)g.C / V;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = (uint32_t)g.N * g.H * g.W * cv;
  if (i >= total) return;
  const int c = (int)(i % cv) * V;
  uint32_t pix = i / cv;
  const int w = (int)(pix % (uint32_t)g.W);
  pix /= (uint32_t)g.W;
  const int h = (int)(pix % (uint32_t)g.H);
  const uint32_t q = pix / (uint32_t)g.H;
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  for (int kh = 0; kh < g.KH; ++kh) {
    const int hn = h + g.PH - kh * g.DH;
    if (hn < 0 || hn % g.SH) continue;
    const int oh = hn / g.SH;
    if (oh >= g.OH) continue;
    for (int kw = 0; kw < g.KW; ++kw) {
      const int wn = w + g.PW - kw * g.DW;
      if (wn < 0 || wn % g.SW) continue;
      const int ow = wn / g.SW;
      if (ow >= g.OW) continue;
      const float* src = Dc + (int64_t)((q * g.OH + oh) * g.OW + ow) * ldd + (kh * g.KW + kw) * g.C + c;
      if (VEC) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src));
        acc[0] += t.x; acc[1 % V] += t.y; acc[2 % V] += t.z; acc[3 % V] += t.w;
      } else {
        acc[0] += __ldg(src);
      }
    }
  }
  float* dst = out + (int64_t)i * V;   // i enumerates (q, h, w, c / V) in memory order
  if (VEC) *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]);
  else *dst = acc[0];
}

int col2im_nhwc(const float* Dc, int64_t ldd, const ConvGeom& g, float* out, cudaStream_t st) {
  if ((int64_t)g.N * g.C * g.H * g.W == 0) return 0;
  const bool vec = (g.C % 4 == 0) && (ldd % 4 == 0) && ((uintptr_t)Dc % 16 == 0) && ((uintptr_t)out % 16 == 0);
  const int64_t total = (int64_t)g.N * g.H * g.W * (vec ? g.C / 4 : g.C);
  const int64_t blocks = ceil_div(total, 256);
  LPB_REQUIRE((int64_t)g.N * g.H * g.W * g.C < (1LL << 31) && (int64_t)g.N * g.OH * g.OW < (1LL << 31),
              "col2im_nhwc: batch too large for 32-bit indexing");
  if (vec) col2im_nhwc_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(Dc, ldd, g, out);
  else col2im_nhwc_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(Dc, ldd, g, out);
  LPB_CHECK_LAUNCH("col2im_nhwc");
  return 0;
}

int col2im(const float* Dc, int64_t ldd, const ConvGeom& g, float* out, cudaStream_t st) {
  const int64_t per_c = (int64_t)g.N * g.H * g.W;
  if (per_c == 0) return 0;
  if (g.H * g.W <= 64 && g.C <= 65535) {
    dim3 grid((unsigned)ceil_div(g.N, 128), (unsigned)g.C);
    col2im_small_kernel<<<grid, 128, (size_t)g.H * g.W * 129 * sizeof(float), st>>>(Dc, ldd, g, out);
    LPB_CHECK_LAUNCH("col2im_small");
    return 0;
  }
  LPB_REQUIRE(per_c < (1LL << 31) && (int64_t)g.N * g.OH * g.OW < (1LL << 31), "col2im: batch too large for 32-bit indexing");
  LPB_REQUIRE(g.C <= 65535, "col2im: too many channels");
  dim3 grid((unsigned)imin(ceil_div(per_c, 256), 8192), (unsigned)g.C);
  col2im_kernel<<<grid, 256, 0, st>>>(Dc, ldd, g, out);
  LPB_CHECK_LAUNCH("col2im");
  return 0;
}

}  // namespace lpb
