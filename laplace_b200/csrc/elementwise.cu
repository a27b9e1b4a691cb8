// Reverse-pass element-wise kernels of the convolution engine: with the C curvature columns folded into the batch,
// PyTorch's vmapped backward formulas for ReLU / frozen BatchNorm / MaxPool run through generic broadcasting kernels
// (19 % + 10 % of a KFAC step in profiles/r01_launches_final_b2048.md).  These are the same maps as HBM-bound,
// vectorised kernels: every gradient element is read once and written once.
#include "common.cuh"

namespace lpb {

// out[i] = g[i] * scale[(i / inner) % C]     (inner = H*W for NCHW, 1 for NHWC rows)
__global__ void __launch_bounds__(256) scale_channels_kernel(const float* __restrict__ g, const float* __restrict__ scale,
                                                              float* __restrict__ out, int64_t n4, int C, int64_t inner,
                                                              int mode) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[e];
    float4 o;
    if (mode == 0) {  // 4 consecutive elements share one channel (inner % 4 == 0)
      const float s = __ldg(scale + ((e * 4) / inner) % C);
      o = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
    } else {          // inner == 1, C % 4 == 0: 4 consecutive channels
      const float4 s = *reinterpret_cast<const float4*>(scale + (e * 4) % C);
      o = make_float4(v.x * s.x, v.y * s.y, v.z * s.z, v.w * s.w);
    }
    reinterpret_cast<float4*>(out)[e] = o;
  }
}

__global__ void __launch_bounds__(256) scale_channels_scalar_kernel(const float* __restrict__ g,
                                                                     const float* __restrict__ scale,
                                                                     float* __restrict__ out, int64_t n, int C,
                                                                     int64_t inner) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = g[e] * __ldg(scale + (e / inner) % C);
}

int scale_channels(const float* g, const float* scale, float* out, int64_t n, int C, int64_t inner, cudaStream_t st) {
  if (n == 0) return 0;
  LPB_REQUIRE(C > 0 && inner > 0, "scale_channels: bad extents");
  const bool aligned = ((uintptr_t)g % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)scale % 16) == 0 && n % 4 == 0;
  const int blocks_cap = sm_count() * 16;
  if (aligned && inner % 4 == 0) {
    scale_channels_kernel<<<(int)imin(ceil_div(n / 4, 256), blocks_cap), 256, 0, st>>>(g, scale, out, n / 4, C, inner, 0);
  } else if (aligned && inner == 1 && C % 4 == 0) {
    scale_channels_kernel<<<(int)imin(ceil_div(n / 4, 256), blocks_cap), 256, 0, st>>>(g, scale, out, n / 4, C, inner, 1);
  } else {
    scale_channels_scalar_kernel<<<(int)imin(ceil_div(n, 256), blocks_cap), 256, 0, st>>>(g, scale, out, n, C, inner);
  }
  LPB_CHECK_LAUNCH("scale_channels");
  return 0;
}

// out[r*n + i] = y[i] > 0 ? g[r*n + i] : 0     (y: forward ReLU output of the B images, shared by all reps)
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                        float* __restrict__ out, int64_t n4, int64_t total4) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[e];
    const float4 m = reinterpret_cast<const float4*>(y)[e % n4];
    reinterpret_cast<float4*>(out)[e] =
        make_float4(m.x > 0.f ? v.x : 0.f, m.y > 0.f ? v.y : 0.f, m.z > 0.f ? v.z : 0.f, m.w > 0.f ? v.w : 0.f);
  }
}

__global__ void __launch_bounds__(256) relu_bwd_scalar_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                               float* __restrict__ out, int64_t n, int64_t total) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = y[e % n] > 0.f ? g[e] : 0.f;
}

// Mask-major variant (see pack_cast_fused_maskmajor_kernel, pack.cu): a thread keeps its 4 mask values and walks the reps
// gradient blocks that share them, so a mask larger than L2 is read once instead of once per folded column.
__global__ void __launch_bounds__(256) relu_bwd_maskmajor_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                                  float* __restrict__ out, int64_t n4, int reps) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
    const float4 m = __ldg(reinterpret_cast<const float4*>(y) + e);
#pragma unroll 4
    for (int r = 0; r < reps; ++r) {
      const float4 v = reinterpret_cast<const float4*>(g)[(int64_t)r * n4 + e];
      reinterpret_cast<float4*>(out)[(int64_t)r * n4 + e] =
          make_float4(m.x > 0.f ? v.x : 0.f, m.y > 0.f ? v.y : 0.f, m.z > 0.f ? v.z : 0.f, m.w > 0.f ? v.w : 0.f);
    }
  }
}

int64_t mask_major_min();   // pack.cu

int relu_bwd(const float* g, const float* y, float* out, int64_t n, int reps, cudaStream_t st) {
  if (n == 0 || reps == 0) return 0;
  const int64_t total = n * reps;
  const int blocks_cap = sm_count() * 16;
  const bool vec = n % 4 == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)out % 16) == 0;
  if (vec && reps > 1 && mask_major_min() >= 0 && n >= mask_major_min()) {
    relu_bwd_maskmajor_kernel<<<(int)imin(ceil_div(n / 4, 256), sm_count() * 32), 256, 0, st>>>(g, y, out, n / 4, reps);
    LPB_CHECK_LAUNCH("relu_bwd");
    return 0;
  }
  if (vec)
    relu_bwd_kernel<<<(int)imin(ceil_div(total / 4, 256), blocks_cap), 256, 0, st>>>(g, y, out, n / 4, total / 4);
  else
    relu_bwd_scalar_kernel<<<(int)imin(ceil_div(total, 256), blocks_cap), 256, 0, st>>>(g, y, out, n, total);
  LPB_CHECK_LAUNCH("relu_bwd");
  return 0;
}

// max-pool backward, gather form (no atomics): one CTA per (q, c) plane; the plane's OH*OW gradients and argmax
// indices are staged in shared memory, every input pixel then sums the (at most ceil(k/s)^2) windows whose argmax it is.
// g [Q, C, OH, OW], idx [Nb, C, OH, OW] (flattened h*W + w of the argmax, forward of the Nb images; q -> q % Nb),
// out [Q, C, H, W]; all NCHW-contiguous.
// MaxPool backward for `cols` stacked gradient columns that share one argmax map (the vmapped reverse pass:
// q = col * Nb + n).  One CTA per (n, c) plane, one thread per input pixel (block = W x H).  The pixel's candidate
// windows -- at most NC x NC, fixed by (h, w) -- and whether each window's argmax IS this pixel are evaluated once
// from the int64 argmax map; the column loop then only issues independent gradient loads (served by L1: each plane of
// gradients is read by the <= NC^2 pixels sharing a window) and one coalesced store per column.  The argmax map is
// read once instead of once per column, and no load depends on another (the previous version chained
// idx -> compare -> gradient per plane and ran at ~1 TB/s).
template <int NC>
__global__ void maxpool2d_bwd_kernel(const float* __restrict__ g, const int64_t* __restrict__ idx, float* __restrict__ out,
                                     int cols, int planes_per_col, int H, int W, int OH, int OW, int k, int s, int p) {
  const int w = threadIdx.x, h = threadIdx.y;
  const int T = OH * OW, HW = H * W, me = h * W + w;
  const int plane = blockIdx.x;
  const int oh_lo = max(0, (h + p - k + s) / s), oh_hi = min(OH - 1, (h + p) / s);
  const int ow_lo = max(0, (w + p - k + s) / s), ow_hi = min(OW - 1, (w + p) / s);
  const int64_t* ip = idx + (int64_t)plane * T;
  int tpos[NC * NC];
  float sel[NC * NC];
#pragma unroll
  for (int a = 0; a < NC; ++a)
#pragma unroll
    for (int b = 0; b < NC; ++b) {
      const int oh = oh_lo + a, ow = ow_lo + b;
      const bool in = oh <= oh_hi && ow <= ow_hi;
      const int t = in ? oh * OW + ow : 0;
      tpos[a * NC + b] = t;
      sel[a * NC + b] = (in && (int)__ldg(ip + t) == me) ? 1.f : 0.f;
    }
  const int64_t gstride = (int64_t)planes_per_col * T, ostride = (int64_t)planes_per_col * HW;
  const float* gp = g + (int64_t)plane * T;
  float* op = out + (int64_t)plane * HW + me;
#pragma unroll 4
  for (int col = 0; col < cols; ++col) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NC * NC; ++i) acc = fmaf(sel[i], __ldg(gp + tpos[i]), acc);
    *op = acc;
    gp += gstride;
    op += ostride;
  }
}

// Channels-last form of the same reverse pass: g [Q, OH, OW, C], argmax map [Nb, OH, OW, C] (values h * W + w),
// out [Q, H, W, C].  One thread per (n, h, w, V channels), V = 4 when C % 4 == 0: consecutive threads walk the
// channels, so the argmax loads, the gradient loads (float4) and the stores (float4) are all coalesced; window
// membership is evaluated once and reused by every column.  32-bit index arithmetic (64-bit runtime divisions made an
// earlier version ALU-bound at 1.5 TB/s).
template <int NC, int V>
__global__ void __launch_bounds__(256) maxpool2d_bwd_nhwc_kernel(const float* __restrict__ g, const int64_t* __restrict__ idx,
                                                                 float* __restrict__ out, int cols, int Nb, int C, int H,
                                                                 int W, int OH, int OW, int k, int s, int p) {
  const uint32_t cv = (uint32_t)C / V;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = (uint32_t)Nb * H * W * cv;
  if (i >= total) return;
  const uint32_t c = (i % cv) * V;
  uint32_t pix = i / cv;
  const int w = (int)(pix % (uint32_t)W);
  pix /= (uint32_t)W;
  const int h = (int)(pix % (uint32_t)H);
  const uint32_t n = pix / (uint32_t)H;
  const int me = h * W + w;
  const int oh_lo = max(0, (h + p - k + s) / s), oh_hi = min(OH - 1, (h + p) / s);
  const int ow_lo = max(0, (w + p - k + s) / s), ow_hi = min(OW - 1, (w + p) / s);
  uint32_t tpos[NC * NC];
  float sel[NC * NC][V];
#pragma unroll
  for (int a = 0; a < NC; ++a)
#pragma unroll
    for (int b = 0; b < NC; ++b) {
      const int oh = oh_lo + a, ow = ow_lo + b;
      const bool in = oh <= oh_hi && ow <= ow_hi;
      const uint32_t t = ((n * OH + (in ? oh : 0)) * OW + (in ? ow : 0)) * C + c;   // < Nb*OH*OW*C <= total
      tpos[a * NC + b] = t;
#pragma unroll
      for (int v = 0; v < V; ++v) sel[a * NC + b][v] = (in && (int)__ldg(idx + t + v) == me) ? 1.f : 0.f;
    }
  const int64_t gstride = (int64_t)Nb * OH * OW * C, per_col = (int64_t)Nb * H * W * C;
  float* op = out + (int64_t)i * V;
#pragma unroll 2
  for (int col = 0; col < cols; ++col) {
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
#pragma unroll
    for (int j = 0; j < NC * NC; ++j) {
      if (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(g + tpos[j]));
        acc[0] = fmaf(sel[j][0], t.x, acc[0]); acc[1 % V] = fmaf(sel[j][1 % V], t.y, acc[1 % V]);
        acc[2 % V] = fmaf(sel[j][2 % V], t.z, acc[2 % V]); acc[3 % V] = fmaf(sel[j][3 % V], t.w, acc[3 % V]);
      } else {
        acc[0] = fmaf(sel[j][0], __ldg(g + tpos[j]), acc[0]);
      }
    }
    if (V == 4) *reinterpret_cast<float4*>(op) = make_float4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]);
    else *op = acc[0];
    g += gstride;
    op += per_col;
  }
}

int maxpool2d_bwd_nhwc(const float* g, const int64_t* idx, float* out, int64_t Q, int Nb, int C, int H, int W, int OH,
                       int OW, int k, int s, int p, cudaStream_t st) {
  if (Q * C == 0) return 0;
  LPB_REQUIRE(Nb > 0 && k > 0 && s > 0 && p >= 0, "maxpool2d_bwd_nhwc: bad geometry");
  LPB_REQUIRE(Q % Nb == 0, "maxpool2d_bwd_nhwc: Q must be a multiple of the argmax batch Nb");
  const int nc = (k + s - 1) / s;
  LPB_REQUIRE(nc <= 3, "maxpool2d_bwd_nhwc: kernel_size > 3 * stride is not supported");
  LPB_REQUIRE(OH <= H && OW <= W, "maxpool2d_bwd_nhwc: output larger than input");
  const bool vec = (C % 4 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)out % 16 == 0);
  const int64_t per_col = (int64_t)Nb * H * W * C;
  LPB_REQUIRE(per_col < (1LL << 31), "maxpool2d_bwd_nhwc: argmax batch too large for 32-bit indexing");
  const int64_t blocks = ceil_div(vec ? per_col / 4 : per_col, 256);
  const int cols = (int)(Q / Nb);
#define LPB_POOL(NCV, VV) \
  maxpool2d_bwd_nhwc_kernel<NCV, VV><<<(unsigned)blocks, 256, 0, st>>>(g, idx, out, cols, Nb, C, H, W, OH, OW, k, s, p)
  if (vec) {
    if (nc <= 1) LPB_POOL(1, 4); else if (nc == 2) LPB_POOL(2, 4); else LPB_POOL(3, 4);
  } else {
    if (nc <= 1) LPB_POOL(1, 1); else if (nc == 2) LPB_POOL(2, 1); else LPB_POOL(3, 1);
  }
#undef LPB_POOL
  LPB_CHECK_LAUNCH("maxpool2d_bwd_nhwc");
  return 0;
}

// Fused form for a max-pool that directly follows a fused conv -> (frozen BN) -> ReLU chain (the ResNet stem): the un-pooled
// gradient is never written as fp32 -- it is multiplied by the ReLU mask (y > 0, y = the pool's input) and the BN scale and
// emitted straight as the bf16 hi/lo operand rows [(col, n, h, w), C] of the chain's backward-data convolution and B-factor
// SYRK.  Saves one fp32 write and one fp32 read of the largest gradient tensor of the network (2 x 2.7 GB per step for
// ResNet-18 at batch 4096).  One thread per (n, h, w, 4 channels) like maxpool2d_bwd_nhwc_kernel<., 4>.
template <int NC>
__global__ void __launch_bounds__(256) maxpool2d_bwd_pack_nhwc_kernel(const float* __restrict__ g, const int64_t* __restrict__ idx,
                                                                      const float* __restrict__ scale, const float* __restrict__ y,
                                                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                                      int64_t ld, int cols, int Nb, int C, int H, int W, int OH, int OW,
                                                                      int k, int s, int p) {
  constexpr int V = 4;
  const uint32_t cv = (uint32_t)C / V;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = (uint32_t)Nb * H * W * cv;
  if (i >= total) return;
  const uint32_t c = (i % cv) * V;
  uint32_t pix = i / cv;
  const uint32_t row0 = pix;                    // (n, h, w) row of column 0
  const int w = (int)(pix % (uint32_t)W);
  pix /= (uint32_t)W;
  const int h = (int)(pix % (uint32_t)H);
  const uint32_t n = pix / (uint32_t)H;
  const int me = h * W + w;
  const int oh_lo = max(0, (h + p - k + s) / s), oh_hi = min(OH - 1, (h + p) / s);
  const int ow_lo = max(0, (w + p - k + s) / s), ow_hi = min(OW - 1, (w + p) / s);
  float mul[V];
#pragma unroll
  for (int v = 0; v < V; ++v) mul[v] = scale ? __ldg(scale + c + v) : 1.f;
  if (y) {
    const float4 yy = __ldg(reinterpret_cast<const float4*>(y + (int64_t)row0 * C + c));
    mul[0] = yy.x > 0.f ? mul[0] : 0.f; mul[1] = yy.y > 0.f ? mul[1] : 0.f;
    mul[2] = yy.z > 0.f ? mul[2] : 0.f; mul[3] = yy.w > 0.f ? mul[3] : 0.f;
  }
  uint32_t tpos[NC * NC];
  float sel[NC * NC][V];
#pragma unroll
  for (int a = 0; a < NC; ++a)
#pragma unroll
    for (int b = 0; b < NC; ++b) {
      const int oh = oh_lo + a, ow = ow_lo + b;
      const bool in = oh <= oh_hi && ow <= ow_hi;
      const uint32_t t = ((n * OH + (in ? oh : 0)) * OW + (in ? ow : 0)) * C + c;
      tpos[a * NC + b] = t;
#pragma unroll
      for (int v = 0; v < V; ++v) sel[a * NC + b][v] = (in && (int)__ldg(idx + t + v) == me) ? mul[v] : 0.f;
    }
  const int64_t gstride = (int64_t)Nb * OH * OW * C, rows_per_col = (int64_t)Nb * H * W;
  int64_t row = row0;
#pragma unroll 2
  for (int col = 0; col < cols; ++col) {
    float acc[V] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NC * NC; ++j) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(g + tpos[j]));
      acc[0] = fmaf(sel[j][0], t.x, acc[0]); acc[1] = fmaf(sel[j][1], t.y, acc[1]);
      acc[2] = fmaf(sel[j][2], t.z, acc[2]); acc[3] = fmaf(sel[j][3], t.w, acc[3]);
    }
    alignas(8) __nv_bfloat16 hh[V], ll[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      hh[v] = __float2bfloat16_rn(acc[v]);
      ll[v] = __float2bfloat16_rn(acc[v] - __bfloat162float(hh[v]));
    }
    *reinterpret_cast<uint2*>(hi + row * ld + c) = *reinterpret_cast<const uint2*>(hh);
    *reinterpret_cast<uint2*>(lo + row * ld + c) = *reinterpret_cast<const uint2*>(ll);
    g += gstride;
    row += rows_per_col;
  }
}

int maxpool2d_bwd_pack_nhwc(const float* g, const int64_t* idx, const float* scale, const float* y, void* hi, void* lo, int64_t ld,
                            int64_t Q, int Nb, int C, int H, int W, int OH, int OW, int k, int s, int p, cudaStream_t st) {
  if (Q * C == 0) return 0;
  LPB_REQUIRE(Nb > 0 && k > 0 && s > 0 && p >= 0 && Q % Nb == 0, "maxpool2d_bwd_pack_nhwc: bad geometry");
  const int nc = (k + s - 1) / s;
  LPB_REQUIRE(nc <= 3, "maxpool2d_bwd_pack_nhwc: kernel_size > 3 * stride is not supported");
  LPB_REQUIRE(OH <= H && OW <= W, "maxpool2d_bwd_pack_nhwc: output larger than input");
  LPB_REQUIRE(C % 4 == 0 && ld % 4 == 0 && ld >= C && ((uintptr_t)g % 16) == 0 && ((uintptr_t)hi % 8) == 0 && ((uintptr_t)lo % 8) == 0 &&
                  (y == nullptr || ((uintptr_t)y % 16) == 0),
              "maxpool2d_bwd_pack_nhwc: needs C % 4 == 0 and 16-byte aligned operands");
  const int64_t per_col = (int64_t)Nb * H * W * C;
  LPB_REQUIRE(per_col < (1LL << 31), "maxpool2d_bwd_pack_nhwc: argmax batch too large for 32-bit indexing");
  const int64_t blocks = ceil_div(per_col / 4, 256);
  const int cols = (int)(Q / Nb);
  __nv_bfloat16 *h = reinterpret_cast<__nv_bfloat16*>(hi), *l = reinterpret_cast<__nv_bfloat16*>(lo);
  if (nc <= 1) maxpool2d_bwd_pack_nhwc_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(g, idx, scale, y, h, l, ld, cols, Nb, C, H, W, OH, OW, k, s, p);
  else if (nc == 2) maxpool2d_bwd_pack_nhwc_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(g, idx, scale, y, h, l, ld, cols, Nb, C, H, W, OH, OW, k, s, p);
  else maxpool2d_bwd_pack_nhwc_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(g, idx, scale, y, h, l, ld, cols, Nb, C, H, W, OH, OW, k, s, p);
  LPB_CHECK_LAUNCH("maxpool2d_bwd_pack_nhwc");
  return 0;
}

int maxpool2d_bwd(const float* g, const int64_t* idx, float* out, int64_t Q, int Nb, int C, int H, int W, int OH, int OW,
                  int k, int s, int p, cudaStream_t st) {
  if (Q * C == 0) return 0;
  LPB_REQUIRE(Nb > 0 && k > 0 && s > 0 && p >= 0, "maxpool2d_bwd: bad geometry");
  LPB_REQUIRE(Q % Nb == 0, "maxpool2d_bwd: Q must be a multiple of the argmax batch Nb");
  LPB_REQUIRE((int64_t)Nb * C < (1LL << 31), "maxpool2d_bwd: too many planes");
  LPB_REQUIRE((int64_t)H * W <= 1024, "maxpool2d_bwd: plane larger than one thread block (H*W <= 1024)");
  const int nc = (k + s - 1) / s;   // windows covering one pixel, per dimension
  LPB_REQUIRE(nc <= 3, "maxpool2d_bwd: kernel_size > 3 * stride is not supported");
  dim3 block(W, H);
  const unsigned blocks = (unsigned)(Nb * C);
  const int cols = (int)(Q / Nb);
  if (nc <= 1) maxpool2d_bwd_kernel<1><<<blocks, block, 0, st>>>(g, idx, out, cols, Nb * C, H, W, OH, OW, k, s, p);
  else if (nc == 2) maxpool2d_bwd_kernel<2><<<blocks, block, 0, st>>>(g, idx, out, cols, Nb * C, H, W, OH, OW, k, s, p);
  else maxpool2d_bwd_kernel<3><<<blocks, block, 0, st>>>(g, idx, out, cols, Nb * C, H, W, OH, OW, k, s, p);
  LPB_CHECK_LAUNCH("maxpool2d_bwd");
  return 0;
}

}  // namespace lpb
