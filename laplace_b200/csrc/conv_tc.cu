// Implicit-GEMM convolution on the tcgen05 tensor cores (stride 1, "same" padding, NHWC operands):
//
//     D[(q,h,w), n] = alpha * sum_{kh,kw} sum_k  X[q, h + bh + s*kh, w + bw + s*kw, k] * Wt[(kh,kw), n, k]
//
// forward:        X = layer input,    Wt[(kh,kw), co, ci] = W[co,ci,kh,kw],   (bh,bw,s) = (-p,-p,+1)
// backward-data:  X = output gradient, Wt[(kh,kw), ci, co] = W[co,ci,kh,kw],   (bh,bw,s) = (+p,+p,-1)
//
// No im2col / col2im ever touches HBM: the A operand of every tap is a *shifted 4-D TMA box* of the NHWC
// tensor ({64 channels, W, H, 128/(H*W) images}, 128B swizzle, out-of-bounds rows zero-filled by the TMA
// unit = the zero padding of the convolution).  One CTA owns 128 consecutive NHWC output rows x 128 output
// channels, loops over taps x channel chunks through the same full/empty mbarrier ring as the GEMM kernel,
// accumulates in TMEM and stores the fp32 tile once.  bf16 hi/lo operands, three products per tile
// (NPROD = 3) give fp32-level accuracy.
#include <stdlib.h>

#include "tc_common.cuh"

namespace lpb {
namespace tc {

template <int NPROD>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_nhwc_tc_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                    const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, int64_t Mrows,
                    int N, float alpha, float* __restrict__ D, int64_t ldd, int q_per_tile, int KH, int KW, int base_h,
                    int base_w, int sgn, int kchunks, int num_stages, int fp16_operands, int bn) {
  // bn = MMA N extent (64 or 128): layers with <= 64 output channels would waste half of a 128-wide tile
  const int B_BYTES = bn * BK * 2;
  const int STAGE_BYTES = (NPROD == 3 ? 2 : 1) * (TILE_BYTES + B_BYTES);
  const int OFF_B_HI = TILE_BYTES, OFF_A_LO = TILE_BYTES + B_BYTES, OFF_B_LO = 2 * TILE_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)num_stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + num_stages;
  uint64_t* tmem_full_bar = empty_bar + num_stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int tm = blockIdx.x, tn = blockIdx.y;
  const int q0 = tm * q_per_tile;
  const int total = KH * KW * kchunks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < total; ++it) {
        const int tap = it / kchunks, kc = it - tap * kchunks;
        const int kh = tap / KW, kw = tap - kh * KW;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        const int ch = base_h + sgn * kh, cw = base_w + sgn * kw;
        tma_load_4d(&tmX_hi, &full_bar[stage], st, kc * BK, cw, ch, q0);
        tma_load_2d(&tmW_hi, &full_bar[stage], st + OFF_B_HI, kc * BK, tap * N + tn * bn);
        if (NPROD == 3) {
          tma_load_4d(&tmX_lo, &full_bar[stage], st + OFF_A_LO, kc * BK, cw, ch, q0);
          tma_load_2d(&tmW_lo, &full_bar[stage], st + OFF_B_LO, kc * BK, tap * N + tn * bn);
        }
        if (++stage == num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, bn, fp16_operands);
      int stage = 0; uint32_t phase = 0; uint32_t acc = 0;
      for (int it = 0; it < total; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint64_t a_hi = make_smem_desc(sbase), b_hi = make_smem_desc(sbase + OFF_B_HI);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
          umma_f16(tmem_base, a_hi + koff, b_hi + koff, idesc, acc);
          acc = 1;
          if (NPROD == 3) {
            const uint64_t a_lo = make_smem_desc(sbase + OFF_A_LO), b_lo = make_smem_desc(sbase + OFF_B_LO);
            umma_f16(tmem_base, a_hi + koff, b_lo + koff, idesc, 1);
            umma_f16(tmem_base, a_lo + koff, b_hi + koff, idesc, 1);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == num_stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int64_t row = (int64_t)tm * BM + q * 32 + lane;
    const bool vec_ok = ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
#pragma unroll 1
    for (int chunk = 0; chunk < bn / 32; ++chunk) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(chunk * 32), v);
      const int col0 = tn * bn + chunk * 32;
      if (row < Mrows && col0 < N) {
        float* drow = D + row * ldd + col0;
        if (vec_ok && col0 + 32 <= N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(drow + j) = make_float4(alpha * v[j], alpha * v[j + 1], alpha * v[j + 2], alpha * v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < N) drow[j] = alpha * v[j];
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// Persistent form of the same convolution: one CTA per SM walks the (row tile, channel tile) list; taps x channel
// chunks of a tile are cut into groups of <= 16 k-chunks that alternate between two TMEM accumulators, the epilogue
// warps sum the groups in registers and store the tile while the tensor pipe already works on the next one.  Hides the
// per-tile prologue (barrier init, TMEM allocation, first TMA round trip) and epilogue that made the one-tile-per-CTA
// kernel spend ~1/3 of its time outside the MMA loop on 64-channel layers (9 k-chunks per tile).
constexpr int CONV_GROUP = 16;
constexpr int MAX_TAPS = 64;

// taps of one launch: weight row block `w` (rows w*N .. of the tap-major weight matrix) and the shift of the operand box
struct ConvTaps {
  int n;
  short w[MAX_TAPS], dh[MAX_TAPS], dw[MAX_TAPS];
};
// where a tile's rows go: mode 0 = the row itself (dense NHWC output); mode 1 = pixel (i, j) of the operand grid
// (gh x gw per image) lands at (i*sh + ph, j*sw + pw) of an H x W output image (parity classes of a strided reverse pass)
struct ConvOutMap {
  int mode, gh, gw, sh, sw, ph, pw, H, W;
};

template <int NPROD>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_nhwc_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                               const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                               int64_t Mrows, int N, float alpha, float* __restrict__ D, int64_t ldd, int q_per_tile,
                               const __grid_constant__ ConvTaps taps, const __grid_constant__ ConvOutMap om, int kchunks,
                               int num_stages, int fp16_operands, int bn, int tiles_n, int num_items, int tiles_per_img,
                               int rows_per_tile) {
  // NPROD 1: A_hi*B_hi; 2: + A_hi*B_lo (rounded activations / gradients, exact weights); 3: + A_lo*B_hi
  const int B_BYTES = bn * BK * 2;
  const int STAGE_BYTES = NPROD == 3 ? 2 * (TILE_BYTES + B_BYTES) : (NPROD == 2 ? TILE_BYTES + 2 * B_BYTES : TILE_BYTES + B_BYTES);
  const int OFF_B_HI = TILE_BYTES, OFF_A_LO = TILE_BYTES + B_BYTES;
  const int OFF_B_LO = NPROD == 2 ? TILE_BYTES + B_BYTES : 2 * TILE_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)num_stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + num_stages;
  uint64_t* tfull_bar = empty_bar + num_stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int total = taps.n * kchunks;
  const int ngroups = (total + CONV_GROUP - 1) / CONV_GROUP;
  const int glen = (total + ngroups - 1) / ngroups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int tm = item / tiles_n, tn = item - tm * tiles_n;
        // small images: a tile is q_per_tile whole images; large images: rows_per_tile image rows of one image
        int q0 = tm * q_per_tile, h0 = 0;
        if (tiles_per_img > 0) { q0 = tm / tiles_per_img; h0 = (tm - q0 * tiles_per_img) * rows_per_tile; }
        // (tap, channel chunk) stepped, not divided out of `it`: the producer thread's per-iteration arithmetic is on the
        // critical path of the ring (cf. the implicit-patch SYRK loader, gemm_tc3.cu)
        int ti = 0, kc = 0;
        int ch = h0 + taps.dh[0], cw = taps.dw[0], wrow = taps.w[0] * N + tn * bn;
        for (int it = 0; it < total; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_4d(&tmX_hi, &full_bar[stage], st, kc * BK, cw, ch, q0);
          tma_load_2d(&tmW_hi, &full_bar[stage], st + OFF_B_HI, kc * BK, wrow);
          if (NPROD == 3) tma_load_4d(&tmX_lo, &full_bar[stage], st + OFF_A_LO, kc * BK, cw, ch, q0);
          if (NPROD >= 2) tma_load_2d(&tmW_lo, &full_bar[stage], st + OFF_B_LO, kc * BK, wrow);
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
          if (++kc == kchunks) {
            kc = 0;
            if (++ti < taps.n) { ch = h0 + taps.dh[ti]; cw = taps.dw[ti]; wrow = taps.w[ti] * N + tn * bn; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, bn, fp16_operands);
      int stage = 0; uint32_t phase = 0; uint32_t grp = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        for (int g0 = 0; g0 < total; g0 += glen, ++grp) {
          const uint32_t buf = grp & 1;
          mbar_wait(&tempty_bar[buf], ((grp >> 1) & 1) ^ 1);
          tcgen05_fence_after();
          const uint32_t tacc = tmem_base + buf * 128;
          uint32_t acc = 0;
          const int gend = min(total, g0 + glen);
          for (int it = g0; it < gend; ++it) {
            mbar_wait(&full_bar[stage], phase);
            tcgen05_fence_after();
            const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
            const uint64_t a_hi = make_smem_desc(sbase), b_hi = make_smem_desc(sbase + OFF_B_HI);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
              umma_f16(tacc, a_hi + koff, b_hi + koff, idesc, acc);
              acc = 1;
              if (NPROD >= 2) {
                const uint64_t b_lo = make_smem_desc(sbase + OFF_B_LO);
                umma_f16(tacc, a_hi + koff, b_lo + koff, idesc, 1);
              }
              if (NPROD == 3) {
                const uint64_t a_lo = make_smem_desc(sbase + OFF_A_LO);
                umma_f16(tacc, a_lo + koff, b_hi + koff, idesc, 1);
              }
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull_bar[buf]);
        }
      }
    }
  } else {
    const int q = warp & 3;
    const bool vec_ok = ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
    uint32_t grp = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tm = item / tiles_n, tn = item - tm * tiles_n;
      float accv[128];
      for (int g0 = 0; g0 < total; g0 += glen, ++grp) {
        const uint32_t buf = grp & 1;
        mbar_wait(&tfull_bar[buf], (grp >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + buf * 128 + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int chunk = 0; chunk < 4; ++chunk) {
          if (chunk * 32 < bn) {   // uniform
            float v[32];
            tmem_ld32(taddr + (uint32_t)(chunk * 32), v);
            if (g0 == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) accv[chunk * 32 + j] = v[j];
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) accv[chunk * 32 + j] += v[j];
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty_bar[buf])) : "memory");
        }
      }
      const int64_t row = (int64_t)tm * BM + q * 32 + lane;
      if (row < Mrows) {
        int64_t orow = row;
        if (om.mode == 1) {
          const int64_t img = row / (om.gh * om.gw);
          const int rem = (int)(row - img * (om.gh * om.gw));
          const int i = rem / om.gw, j = rem - i * om.gw;
          orow = (img * om.H + (i * om.sh + om.ph)) * om.W + (j * om.sw + om.pw);
        }
#pragma unroll
        for (int chunk = 0; chunk < 4; ++chunk) {
          const int col0 = tn * bn + chunk * 32;
          if (chunk * 32 < bn && col0 < N) {
            float* drow = D + orow * ldd + col0;
            const float* v = accv + chunk * 32;
            if (vec_ok && col0 + 32 <= N) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(drow + j) = make_float4(alpha * v[j], alpha * v[j + 1], alpha * v[j + 2], alpha * v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < N) drow[j] = alpha * v[j];
            }
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

}  // namespace tc

static int make_tmap_nhwc(CUtensorMap* map, const void* ptr, int64_t Q, int H, int W, int64_t Kc, int64_t ld, int q_per_tile,
                          int box_h = 0) {
  PFN_encodeTiled enc = get_tensormap_encoder();
  LPB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)Kc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Q};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)tc::BK, (cuuint32_t)W, (cuuint32_t)(box_h > 0 ? box_h : H), (cuuint32_t)q_per_tile};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LPB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d) failed (%d) Q=%lld H=%d W=%d K=%lld ld=%lld", (int)r,
              (long long)Q, H, W, (long long)Kc, (long long)ld);
  return 0;
}

static int conv_nhwc_core(const void* X_hi, const void* X_lo, int64_t Q, int H, int W, int64_t Kc, int64_t ldx, const void* W_hi,
                          const void* W_lo, int64_t ldw, int64_t w_rows, int N, const tc::ConvTaps& taps, const tc::ConvOutMap& om,
                          float alpha, float* D, int64_t ldd, int fp16_operands, cudaStream_t st) {
  LPB_REQUIRE(Q > 0 && H > 0 && W > 0 && Kc > 0 && N > 0 && taps.n > 0, "conv_nhwc_bf16: bad extents");
  const bool big = H * W > 128;   // tiles of 128 / W image rows instead of whole images
  LPB_REQUIRE(big ? (128 % W == 0 && H % (128 / W) == 0) : (128 % (H * W) == 0),
              "conv_nhwc_bf16: %dx%d images do not tile 128-row blocks", H, W);
  LPB_REQUIRE((ldx % 8) == 0 && (ldw % 8) == 0 && ldx >= Kc && ldw >= Kc, "conv_nhwc_bf16: bad leading dimensions");
  LPB_REQUIRE(X_lo == nullptr || W_lo != nullptr, "conv_nhwc_bf16: an operand lo half needs the weight lo half");
  LPB_REQUIRE(ldd >= N, "conv_nhwc_bf16: ldd too small");
  const bool x3 = X_lo != nullptr;
  const bool x2 = !x3 && W_lo != nullptr;   // rounded operand rows (hi only) against exact (hi + lo) weights: two products
  const int q_per_tile = big ? 1 : 128 / (H * W);
  const int rows_per_tile = big ? 128 / W : H, tiles_per_img = big ? H / rows_per_tile : 0;
  CUtensorMap tX_hi, tX_lo, tW_hi, tW_lo;
  const int bn = N <= 64 ? 64 : 128;
  if (make_tmap_nhwc(&tX_hi, X_hi, Q, H, W, Kc, ldx, q_per_tile, rows_per_tile)) return 1;
  if (make_tmap_2d(&tW_hi, W_hi, w_rows, Kc, ldw, bn)) return 1;
  tX_lo = tX_hi; tW_lo = tW_hi;
  if (x3 && make_tmap_nhwc(&tX_lo, X_lo, Q, H, W, Kc, ldx, q_per_tile, rows_per_tile)) return 1;
  if ((x3 || x2) && make_tmap_2d(&tW_lo, W_lo, w_rows, Kc, ldw, bn)) return 1;
  const int64_t Mrows = Q * H * W;
  const int64_t tiles_m = big ? Q * tiles_per_img : ceil_div(Q, q_per_tile);
  const int tiles_n = (int)ceil_div(N, bn);
  const int kchunks = (int)ceil_div(Kc, tc::BK);
  const int b_bytes = bn * tc::BK * 2;
  const int stage_bytes = x3 ? 2 * (tc::TILE_BYTES + b_bytes) : (x2 ? tc::TILE_BYTES + 2 * b_bytes : tc::TILE_BYTES + b_bytes);
  const int64_t items = tiles_m * tiles_n;
  LPB_REQUIRE(items <= 2147483647LL, "conv_nhwc_bf16: too many tiles");
  const int pstages = (int)imin(8, (196 * 1024) / stage_bytes);
  const size_t psmem = (size_t)pstages * stage_bytes + (2 * pstages + 4) * sizeof(uint64_t) + 16 + 1024;
  static bool pattr1 = false, pattr2 = false, pattr3 = false;
  if (x2 && !pattr2) {
    if (check_cuda(cudaFuncSetAttribute(tc::conv_nhwc_tc_persistent_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        227 * 1024), "conv_nhwc_bf16 attr"))
      return 1;
    pattr2 = true;
  }
  if (x3 && !pattr3) {
    if (check_cuda(cudaFuncSetAttribute(tc::conv_nhwc_tc_persistent_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        227 * 1024), "conv_nhwc_bf16 attr"))
      return 1;
    pattr3 = true;
  }
  if (!x3 && !x2 && !pattr1) {
    if (check_cuda(cudaFuncSetAttribute(tc::conv_nhwc_tc_persistent_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        227 * 1024), "conv_nhwc_bf16 attr"))
      return 1;
    pattr1 = true;
  }
  const unsigned pgrid = (unsigned)imin(items, sm_count());
  if (x3)
    tc::conv_nhwc_tc_persistent_kernel<3><<<pgrid, tc::NUM_THREADS, psmem, st>>>(
        tX_hi, tX_lo, tW_hi, tW_lo, Mrows, N, alpha, D, ldd, q_per_tile, taps, om, kchunks, pstages, fp16_operands, bn, tiles_n,
        (int)items, tiles_per_img, rows_per_tile);
  else if (x2)
    tc::conv_nhwc_tc_persistent_kernel<2><<<pgrid, tc::NUM_THREADS, psmem, st>>>(
        tX_hi, tX_lo, tW_hi, tW_lo, Mrows, N, alpha, D, ldd, q_per_tile, taps, om, kchunks, pstages, fp16_operands, bn, tiles_n,
        (int)items, tiles_per_img, rows_per_tile);
  else
    tc::conv_nhwc_tc_persistent_kernel<1><<<pgrid, tc::NUM_THREADS, psmem, st>>>(
        tX_hi, tX_lo, tW_hi, tW_lo, Mrows, N, alpha, D, ldd, q_per_tile, taps, om, kchunks, pstages, fp16_operands, bn, tiles_n,
        (int)items, tiles_per_img, rows_per_tile);
  LPB_CHECK_LAUNCH("conv_nhwc_bf16 (persistent)");
  return 0;
}

// one-tile-per-CTA schedule (A/B timing only, LPB_CONV_MODE=0; whole-image tiles)
static int conv_nhwc_single(const void* X_hi, const void* X_lo, int64_t Q, int H, int W, int64_t Kc, int64_t ldx, const void* W_hi,
                            const void* W_lo, int64_t ldw, int N, int KH, int KW, int base_h, int base_w, int sgn, float alpha,
                            float* D, int64_t ldd, int fp16_operands, cudaStream_t st) {
  LPB_REQUIRE(H * W <= 128 && 128 % (H * W) == 0, "conv_nhwc_bf16: LPB_CONV_MODE=0 needs H*W dividing 128");
  LPB_REQUIRE((ldx % 8) == 0 && (ldw % 8) == 0 && ldx >= Kc && ldw >= Kc, "conv_nhwc_bf16: bad leading dimensions");
  const bool x3 = X_lo != nullptr;
  const int q_per_tile = 128 / (H * W);
  CUtensorMap tX_hi, tX_lo, tW_hi, tW_lo;
  const int bn = N <= 64 ? 64 : 128;
  if (make_tmap_nhwc(&tX_hi, X_hi, Q, H, W, Kc, ldx, q_per_tile)) return 1;
  if (make_tmap_2d(&tW_hi, W_hi, (int64_t)KH * KW * N, Kc, ldw, bn)) return 1;
  if (x3) {
    if (make_tmap_nhwc(&tX_lo, X_lo, Q, H, W, Kc, ldx, q_per_tile)) return 1;
    if (make_tmap_2d(&tW_lo, W_lo, (int64_t)KH * KW * N, Kc, ldw, bn)) return 1;
  } else {
    tX_lo = tX_hi; tW_lo = tW_hi;
  }
  const int64_t Mrows = Q * H * W;
  const int64_t tiles_m = ceil_div(Q, q_per_tile);
  const int tiles_n = (int)ceil_div(N, bn);
  LPB_REQUIRE(tiles_m <= 2147483647LL && tiles_n <= 65535, "conv_nhwc_bf16: too many tiles");
  const int kchunks = (int)ceil_div(Kc, tc::BK);
  const int stage_bytes = (x3 ? 2 : 1) * (tc::TILE_BYTES + bn * tc::BK * 2);
  const int num_stages = (int)imin(8, (200 * 1024) / stage_bytes);
  const size_t smem = (size_t)num_stages * stage_bytes + (2 * num_stages + 1) * sizeof(uint64_t) + 16 + 1024;
  static bool attr1 = false, attr3 = false;
  if (x3 && !attr3) {
    if (check_cuda(cudaFuncSetAttribute(tc::conv_nhwc_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
                   "conv_nhwc_bf16 attr"))
      return 1;
    attr3 = true;
  }
  if (!x3 && !attr1) {
    if (check_cuda(cudaFuncSetAttribute(tc::conv_nhwc_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
                   "conv_nhwc_bf16 attr"))
      return 1;
    attr1 = true;
  }
  dim3 grid((unsigned)tiles_m, (unsigned)tiles_n);
  if (x3)
    tc::conv_nhwc_tc_kernel<3><<<grid, tc::NUM_THREADS, smem, st>>>(tX_hi, tX_lo, tW_hi, tW_lo, Mrows, N, alpha, D, ldd,
                                                                    q_per_tile, KH, KW, base_h, base_w, sgn, kchunks,
                                                                    num_stages, fp16_operands, bn);
  else
    tc::conv_nhwc_tc_kernel<1><<<grid, tc::NUM_THREADS, smem, st>>>(tX_hi, tX_lo, tW_hi, tW_lo, Mrows, N, alpha, D, ldd,
                                                                    q_per_tile, KH, KW, base_h, base_w, sgn, kchunks,
                                                                    num_stages, fp16_operands, bn);
  LPB_CHECK_LAUNCH("conv_nhwc_bf16");
  return 0;
}

int conv_nhwc_bf16(const void* X_hi, const void* X_lo, int64_t Q, int H, int W, int64_t Kc, int64_t ldx, const void* W_hi,
                   const void* W_lo, int64_t ldw, int N, int KH, int KW, int base_h, int base_w, int sgn, float alpha,
                   float* D, int64_t ldd, int fp16_operands, cudaStream_t st) {
  LPB_REQUIRE(KH > 0 && KW > 0 && KH * KW <= tc::MAX_TAPS, "conv_nhwc_bf16: between 1 and %d taps", tc::MAX_TAPS);
  static int conv_mode = -1;
  if (conv_mode < 0) {
    const char* e = getenv("LPB_CONV_MODE");   // 0: one tile per CTA (A/B timing); default: persistent CTAs
    conv_mode = e ? (atoi(e) != 0 ? 1 : 0) : 1;
  }
  if (conv_mode == 0 && H * W <= 128)
    return conv_nhwc_single(X_hi, X_lo, Q, H, W, Kc, ldx, W_hi, W_lo, ldw, N, KH, KW, base_h, base_w, sgn, alpha, D, ldd,
                            fp16_operands, st);
  // taps whose shifted window lies entirely outside the image (|shift| >= extent: 3x3 kernels on 1x1 / 2x2 maps)
  // only ever read the zero padding -- drop them from the launch
  tc::ConvTaps taps = {};
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      const int dh = base_h + sgn * kh, dw = base_w + sgn * kw;
      if (dh >= H || dh <= -H || dw >= W || dw <= -W) continue;
      const int t = taps.n++;
      taps.w[t] = (short)(kh * KW + kw); taps.dh[t] = (short)dh; taps.dw[t] = (short)dw;
    }
  if (taps.n == 0) {   // every tap reads padding only: the output is zero
    return check_cuda(cudaMemset2DAsync(D, ldd * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)Q * H * W, st),
                      "conv_nhwc_bf16 memset");
  }
  tc::ConvOutMap om = {};
  return conv_nhwc_core(X_hi, X_lo, Q, H, W, Kc, ldx, W_hi, W_lo, ldw, (int64_t)KH * KW * N, N, taps, om, alpha, D, ldd,
                        fp16_operands, st);
}

// Backward-data of a STRIDED convolution as implicit GEMMs: input pixel (h, w) only receives taps with
// kh = h + PH (mod SH), kw = w + PW (mod SW), so the input gradient splits into SH*SW parity classes, each a stride-1
// convolution of the output-gradient grid with its own subset of taps, written to the class's pixels of the NHWC result
// (no [rows, KH*KW*C_in] intermediate, no col2im).  G: gradient rows [(q,oh,ow), Co]; Wt [(kh,kw,ci), Co] tap-major.
int conv_bwd_strided(const void* G_hi, const void* G_lo, int64_t Q, int OH, int OW, int64_t Co, int64_t ldg, const void* W_hi,
                     const void* W_lo, int64_t ldw, int Ci, int KH, int KW, int SH, int SW, int PH, int PW, int H, int W,
                     float* D, int64_t ldd, cudaStream_t st) {
  LPB_REQUIRE(SH >= 1 && SW >= 1 && KH >= 1 && KW >= 1 && KH * KW <= tc::MAX_TAPS, "conv_bwd_strided: bad geometry");
  LPB_REQUIRE(H == OH * SH && W == OW * SW, "conv_bwd_strided: input extent must be stride x output extent (got %dx%d vs %dx%d)",
              H, W, OH, OW);
  LPB_REQUIRE(ldd >= Ci, "conv_bwd_strided: ldd too small");
  // taps of parity class (p, q); a tap whose shifted window lies entirely outside the gradient grid reads padding only
  auto class_taps = [&](int p, int q) {
    tc::ConvTaps taps = {};
    for (int kh = 0; kh < KH; ++kh) {
      if ((p + PH - kh) % SH != 0) continue;              // h = oh*SH - PH + kh with h = i*SH + p
      for (int kw = 0; kw < KW; ++kw) {
        if ((q + PW - kw) % SW != 0) continue;
        const int dh = (p + PH - kh) / SH, dw = (q + PW - kw) / SW;   // oh = i + dh (exact division)
        if (dh >= OH || dh <= -OH || dw >= OW || dw <= -OW) continue;
        const int t = taps.n++;
        taps.w[t] = (short)(kh * KW + kw);
        taps.dh[t] = (short)dh;
        taps.dw[t] = (short)dw;
      }
    }
    return taps;
  };
  bool empty_class = false;
  for (int p = 0; p < SH; ++p)
    for (int q = 0; q < SW; ++q)
      if (class_taps(p, q).n == 0) empty_class = true;
  if (empty_class &&
      check_cuda(cudaMemset2DAsync(D, ldd * sizeof(float), 0, (size_t)Ci * sizeof(float), (size_t)Q * H * W, st), "conv_bwd_strided memset"))
    return 1;
  for (int p = 0; p < SH; ++p)
    for (int q = 0; q < SW; ++q) {
      const tc::ConvTaps taps = class_taps(p, q);
      if (taps.n == 0) continue;
      tc::ConvOutMap om = {1, OH, OW, SH, SW, p, q, H, W};
      if (conv_nhwc_core(G_hi, G_lo, Q, OH, OW, Co, ldg, W_hi, W_lo, ldw, (int64_t)KH * KW * Ci, Ci, taps, om, 1.0f, D, ldd, 0, st))
        return 1;
    }
  return 0;
}

}  // namespace lpb
