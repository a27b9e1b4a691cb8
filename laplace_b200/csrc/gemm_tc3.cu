// Persistent variant of the factor contraction (gemm_tc.cu):  D[M,N] (fp32) += alpha * A * B^T
//
// One CTA per SM walks a static list of work items (output tile x K-range).  Within an item the K-range is cut into
// groups of <= 16 k-chunks (K = 1024: the TMEM accumulator truncates, longer chains show a measurable bias); groups
// alternate between two TMEM accumulators, the epilogue warps drain one while the tensor pipe fills the other and sum
// the group results in registers (round-to-nearest fp32).  An item therefore ends with ONE set of reductions into D
// however long its K-range is.  Compared with the one-tile-per-CTA kernel this
//   * hides the epilogue (TMEM -> registers -> red.global) and the per-CTA prologue (barrier init, TMEM alloc,
//     descriptor fetch) behind the MMAs of the next group / item, and
//   * divides the split-K reduction traffic by the number of groups per item: a d = 64 factor over 2.6 M rows went
//     through 1280 CTAs x 64 KiB of atomics onto the same 64 KiB of D; here it is 148 x 64 KiB.
#include <stdlib.h>

#include "tc_common.cuh"

namespace lpb {

namespace tc {

constexpr int GROUP_CHUNKS = 16;      // k-chunks (K = 1024) chained in one TMEM tile: the accumulator truncates, the
                                      // bias grows ~3.5e-7 per chunk on same-sign sums; group sums are added in fp32 RN
constexpr int TMEM_COLS_P = 256;      // two 128-column accumulators

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct ItemGeom {
  int tm, tn, kc_begin, kc_end;
  bool diag;
};

__device__ __forceinline__ ItemGeom decode_item(int item, int num_tiles, int symmetric, int tiles_m, int tiles_n,
                                                int total_kchunks, int kchunks_per_split) {
  ItemGeom g;
  const int tile = item % num_tiles, split = item / num_tiles;   // neighbours share a K-range: operands meet in L2
  if (symmetric) {
    int t = tile, r = 0, cnt = tiles_m;
    while (t >= cnt) { t -= cnt; ++r; --cnt; }
    g.tm = r; g.tn = r + t;
  } else {
    g.tm = tile / tiles_n; g.tn = tile % tiles_n;
  }
  g.diag = symmetric && (g.tm == g.tn);
  g.kc_begin = split * kchunks_per_split;
  g.kc_end = min(total_kchunks, g.kc_begin + kchunks_per_split);
  return g;
}

// LOADER 0: K-major 2-D operands; 1: row (MN-major) 2-D operands; 2: implicit convolution patches (MN-major, 4-D);
// 3: A = row operand (output-gradient rows), B = implicit patches of the same samples, and every sample's product is
//    SQUARED before it is accumulated:  D[i,j] += alpha * sum_q ( sum_t A[(q,t), i] * patch[(q mod n_images, t), j] )^2
//    -- the diagonal GGN / EF of a convolution weight (per-sample weight gradients, never materialised).
template <int NPROD, int LOADER>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                          const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, int M,
                          int N, float alpha, float* __restrict__ D, int64_t ldd, int symmetric, int tiles_m, int tiles_n,
                          int num_tiles, int num_items, int total_kchunks, int kchunks_per_split, int num_stages,
                          int store_mode, int fp16_operands, PatchGeom pg) {
  constexpr bool MN = LOADER != 0;
  constexpr int TILES_PER_STAGE = NPROD == 3 ? 4 : 2;
  constexpr int STAGE_BYTES = TILES_PER_STAGE * TILE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)num_stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + num_stages;
  uint64_t* tfull_bar = empty_bar + num_stages;   // [2] accumulator ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;           // [2] accumulator drained (4 epilogue warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS_P)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const ItemGeom it = decode_item(item, num_tiles, symmetric, tiles_m, tiles_n, total_kchunks, kchunks_per_split);
        // implicit-patch operands: the (tap, channel block) of the tile's four 64-feature blocks is fixed over the K loop --
        // decode it ONCE per item, and step the sample coordinates (image, first image row) instead of dividing per chunk.
        // (r02: with the decode inside the loop -- ~6 integer divisions per chunk in the single producer thread -- the
        // loader, not the tensor pipe or the L2 fill, bound the kernel: 17 % tensor-pipe activity in the one-product mode.)
        int p_ci[4] = {0, 0, 0, 0}, p_dw[4] = {0, 0, 0, 0}, p_dh[4] = {0, 0, 0, 0};
        int n0 = 0, h0 = 0;
        if (LOADER == 2 || LOADER == 3) {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int fb = (b < 2 ? it.tm : it.tn) * 2 + (b & 1);
            const int live = fb / pg.blocks_per_tap;
            const int tap = pg.tap_of[live < pg.num_taps ? live : 0];
            const int kh = tap / pg.KW, kw = tap - kh * pg.KW;
            // feature blocks past the last tap read channel coordinate Ci: entirely out of bounds = zeros
            p_ci[b] = live < pg.num_taps ? (fb - live * pg.blocks_per_tap) * 64 : pg.Ci;
            p_dw[b] = kw - pg.PW;
            p_dh[b] = kh - pg.PH;
          }
          if (pg.chunks_per_img > 0) { n0 = it.kc_begin / pg.chunks_per_img; h0 = (it.kc_begin - n0 * pg.chunks_per_img) * pg.rows_per_chunk; }
          else { n0 = it.kc_begin * pg.imgs_per_chunk; h0 = 0; }
          if (LOADER == 3 && pg.n_images > 0) n0 %= pg.n_images;
        }
        const int img_rows = pg.chunks_per_img * pg.rows_per_chunk;   // image height when an image spans several chunks
        for (int kc = it.kc_begin; kc < it.kc_end; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], it.diag ? STAGE_BYTES / 2 : STAGE_BYTES);
          auto load_tile = [&](const CUtensorMap* map, uint8_t* dst, int tile, bool is_b) {
            if (LOADER == 2 || (LOADER == 3 && is_b)) {
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                const int q = (is_b ? 2 : 0) + b;
                tma_load_4d(map, &full_bar[stage], dst + b * (TILE_BYTES / 2), p_ci[q], p_dw[q], h0 + p_dh[q], n0);
              }
            } else if (MN) {
              tma_load_2d(map, &full_bar[stage], dst, tile * BM, kc * BK);
              tma_load_2d(map, &full_bar[stage], dst + TILE_BYTES / 2, tile * BM + 64, kc * BK);
            } else {
              tma_load_2d(map, &full_bar[stage], dst, kc * BK, tile * BM);
            }
          };
          load_tile(&tmA_hi, st, it.tm, false);
          if (!it.diag) load_tile(&tmB_hi, st + TILE_BYTES, it.tn, true);
          if (NPROD == 3) {
            load_tile(&tmA_lo, st + 2 * TILE_BYTES, it.tm, false);
            if (!it.diag) load_tile(&tmB_lo, st + 3 * TILE_BYTES, it.tn, true);
          }
          if (LOADER == 2 || LOADER == 3) {   // next chunk's sample coordinates
            if (pg.chunks_per_img > 0) {
              h0 += pg.rows_per_chunk;
              if (h0 >= img_rows) { h0 = 0; ++n0; }
            } else {
              n0 += pg.imgs_per_chunk;
            }
            if (LOADER == 3 && pg.n_images > 0 && n0 >= pg.n_images) n0 -= pg.n_images;
          }
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, BN, fp16_operands, MN ? 1 : 0);
      auto mk = [](uint32_t saddr) { return MN ? make_smem_desc_mn(saddr, TILE_BYTES / 2) : make_smem_desc(saddr); };
      int stage = 0; uint32_t phase = 0; uint32_t grp = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const ItemGeom it = decode_item(item, num_tiles, symmetric, tiles_m, tiles_n, total_kchunks, kchunks_per_split);
        const int n = it.kc_end - it.kc_begin;
        const int ngroups = (n + GROUP_CHUNKS - 1) / GROUP_CHUNKS;
        const int glen = (LOADER == 3) ? pg.group_chunks : (n + ngroups - 1) / ngroups;
        for (int g0 = 0; g0 < n; g0 += glen, ++grp) {
          const uint32_t buf = grp & 1;
          mbar_wait(&tempty_bar[buf], ((grp >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
          tcgen05_fence_after();
          const uint32_t tacc = tmem_base + buf * BN;
          uint32_t acc = 0;
          const int gend = min(n, g0 + glen);
          for (int c = g0; c < gend; ++c) {
            mbar_wait(&full_bar[stage], phase);
            tcgen05_fence_after();
            const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
            const uint64_t a_hi = mk(sbase), b_hi = it.diag ? a_hi : mk(sbase + TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
              umma_f16(tacc, a_hi + koff, b_hi + koff, idesc, acc);
              acc = 1;
              if (NPROD == 3) {
                const uint64_t a_lo = mk(sbase + 2 * TILE_BYTES);
                const uint64_t b_lo = it.diag ? a_lo : mk(sbase + 3 * TILE_BYTES);
                umma_f16(tacc, a_hi + koff, b_lo + koff, idesc, 1);
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
    // ================= epilogue: TMEM -> registers (sum over groups) -> D =================
    const int q = warp & 3;
    const bool vec_ok = ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
    uint32_t grp = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const ItemGeom it = decode_item(item, num_tiles, symmetric, tiles_m, tiles_n, total_kchunks, kchunks_per_split);
      const int n = it.kc_end - it.kc_begin;
      const int ngroups = (n + GROUP_CHUNKS - 1) / GROUP_CHUNKS;
      const int glen = (LOADER == 3) ? pg.group_chunks : (n + ngroups - 1) / ngroups;
      float accv[BN];
      for (int g0 = 0; g0 < n; g0 += glen, ++grp) {
        const uint32_t buf = grp & 1;
        mbar_wait(&tfull_bar[buf], (grp >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int chunk = 0; chunk < BN / 32; ++chunk) {
          float v[32];
          tmem_ld32(taddr + (uint32_t)(chunk * 32), v);
          if (LOADER == 3) {   // one group = one sample: square, then sum over samples
            if (g0 == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) accv[chunk * 32 + j] = v[j] * v[j];
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) accv[chunk * 32 + j] = fmaf(v[j], v[j], accv[chunk * 32 + j]);
            }
          } else if (g0 == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) accv[chunk * 32 + j] = v[j];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) accv[chunk * 32 + j] += v[j];
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      }
      // ---- one output pass per item ----
      const int row = it.tm * BM + q * 32 + lane;
      const bool mirror = symmetric && (it.tm != it.tn);
      if (row < M) {
#pragma unroll
        for (int chunk = 0; chunk < BN / 32; ++chunk) {
          const int col0 = it.tn * BN + chunk * 32;
          if (col0 >= N) continue;
          float* drow = D + (int64_t)row * ldd + col0;
          const float* v = accv + chunk * 32;
          if (store_mode) {
            if (vec_ok && col0 + 32 <= N) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(drow + j) = make_float4(alpha * v[j], alpha * v[j + 1], alpha * v[j + 2], alpha * v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < N) drow[j] = alpha * v[j];
            }
          } else if (vec_ok && col0 + 32 <= N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) red_add_v4(drow + j, alpha * v[j], alpha * v[j + 1], alpha * v[j + 2], alpha * v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) atomicAdd(drow + j, alpha * v[j]);
          }
          if (mirror) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) atomicAdd(D + (int64_t)(col0 + j) * ldd + row, alpha * v[j]);
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS_P) : "memory");
  }
}

}  // namespace tc

// Split-K choice for the persistent schedule: minimise  waves x (chunks per item + epilogue)  over the number of splits.
int launch_gemm_tc_persistent(bool mn, bool x3, const CUtensorMap& tA_hi, const CUtensorMap& tA_lo, const CUtensorMap& tB_hi,
                              const CUtensorMap& tB_lo, int64_t M, int64_t N, float alpha, float* D, int64_t ldd,
                              int symmetric, int tiles_m, int tiles_n, int64_t num_tiles, int total_kchunks,
                              int kchunks_per_split, int splits, int store_mode, int fp16_operands, int ctas,
                              cudaStream_t st, const tc::PatchGeom* patches);

void persistent_schedule(int64_t tiles, int total_kchunks, int ctas, bool allow_single_store, int* kchunks_per_split,
                         int* splits, bool* single) {
  const double EPI = 3.0;   // one item's output pass, in k-chunk equivalents (mostly hidden behind the next item)
  double best = 1e300, cost1 = 1e300;
  int best_s = 1;
  auto consider = [&](int64_t s64) {
    if (s64 < 1 || s64 > total_kchunks) return;
    const int kps = (int)ceil_div(total_kchunks, s64);
    const int s = (int)ceil_div(total_kchunks, kps);   // the split count this chunk size really produces
    const double waves = (double)ceil_div(tiles * s, (int64_t)ctas);
    const double cost = waves * (kps + EPI);
    if (s == 1) cost1 = cost;
    if (cost < best * 0.999 || (cost < best * 1.001 && s < best_s)) { best = cost; best_s = s; }
  };
  // only split counts that fill w waves exactly (or nearly) can be optimal: s = floor(w * ctas / tiles), w = 1, 2, ...
  consider(1);
  for (int w = 1; w <= 64; ++w) {
    const int64_t s = (int64_t)w * ctas / tiles;
    consider(s);
    consider(s + 1);
    if (s >= total_kchunks) break;
  }
  // overwrite semantics: plain stores (no memset, no atomics) if a single split is nearly as good
  if (allow_single_store && cost1 <= 1.15 * best) best_s = 1;
  *splits = best_s;
  *kchunks_per_split = (int)ceil_div(total_kchunks, best_s);
  *single = best_s == 1;
}

int launch_gemm_tc_persistent(bool mn, bool x3, const CUtensorMap& tA_hi, const CUtensorMap& tA_lo, const CUtensorMap& tB_hi,
                              const CUtensorMap& tB_lo, int64_t M, int64_t N, float alpha, float* D, int64_t ldd,
                              int symmetric, int tiles_m, int tiles_n, int64_t num_tiles, int total_kchunks,
                              int kchunks_per_split, int splits, int store_mode, int fp16_operands, int ctas,
                              cudaStream_t st, const tc::PatchGeom* patches) {
  const int64_t items = num_tiles * splits;
  LPB_REQUIRE(items <= 2147483647LL, "gemm_tc_persistent: too many work items");
  const int stage_bytes = (x3 ? 4 : 2) * tc::TILE_BYTES;
  const int num_stages = x3 ? 3 : 6;
  const size_t smem = (size_t)num_stages * stage_bytes + (2 * num_stages + 4) * sizeof(uint64_t) + 16 + 1024;
  const unsigned grid = (unsigned)imin(items, ctas);
  tc::PatchGeom pg = {};
  if (patches) pg = *patches;
#define LPB_LAUNCH_P(NP, MNV)                                                                                            \
  do {                                                                                                                   \
    static bool attr_done = false;                                                                                       \
    if (!attr_done) {                                                                                                    \
      if (check_cuda(cudaFuncSetAttribute(tc::gemm_tc_persistent_kernel<NP, MNV>,                                        \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),                      \
                     "gemm_tc_persistent attr"))                                                                         \
        return 1;                                                                                                        \
      attr_done = true;                                                                                                  \
    }                                                                                                                    \
    tc::gemm_tc_persistent_kernel<NP, MNV><<<grid, tc::NUM_THREADS, smem, st>>>(                                         \
        tA_hi, tA_lo, tB_hi, tB_lo, (int)M, (int)N, alpha, D, ldd, symmetric, tiles_m, tiles_n, (int)num_tiles,          \
        (int)items, total_kchunks, kchunks_per_split, num_stages, store_mode, fp16_operands, pg);                        \
  } while (0)
  if (patches && patches->group_chunks > 0) {
    if (x3) LPB_LAUNCH_P(3, 3);
    else LPB_LAUNCH_P(1, 3);
  } else if (patches) {
    if (x3) LPB_LAUNCH_P(3, 2);
    else LPB_LAUNCH_P(1, 2);
  } else if (x3 && mn) LPB_LAUNCH_P(3, 1);
  else if (x3) LPB_LAUNCH_P(3, 0);
  else if (mn) LPB_LAUNCH_P(1, 1);
  else LPB_LAUNCH_P(1, 0);
#undef LPB_LAUNCH_P
  LPB_CHECK_LAUNCH("gemm_tc_persistent");
  return 0;
}

// ------------------------------------------------------------------------------------------
// KFAC input factor of a stride-1 "same" convolution straight from the NHWC activation rows:
//     D[(t,ci),(t',cj)] (+)= alpha * sum_{n,h,w} x[n, h+kh-PH, w+kw-PW, ci] * x[n, h+kh'-PH, w+kw'-PW, cj]
// (tap-major feature order; taps_to_param_accumulate() below folds it into the parameter order (ci,kh,kw)).
// The patch matrix [(n,h,w), KH*KW*Ci] -- 9x the activation for 3x3 kernels -- is never written or read.
// row-major [rows, cols] 16-bit matrix, boxes of 64 columns x 64 rows (the MN-major operand tiles of gemm_tc.cu)
static int make_tmap_rows_ext(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld) {
  PFN_encodeTiled enc = get_tensormap_encoder();
  LPB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LPB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rows) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows,
              (long long)cols, (long long)ld);
  return 0;
}

static int make_tmap_patches(CUtensorMap* map, const void* ptr, int64_t Q, int H, int W, int64_t Ci, int64_t ld, int box_h,
                             int box_n) {
  PFN_encodeTiled enc = get_tensormap_encoder();
  LPB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)Ci, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Q};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)box_h, (cuuint32_t)box_n};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LPB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(patches) failed (%d) Q=%lld H=%d W=%d Ci=%lld ld=%lld", (int)r,
              (long long)Q, H, W, (long long)Ci, (long long)ld);
  return 0;
}

int syrk_conv_patches(const void* X_hi, const void* X_lo, int64_t ldx, int64_t Q, int H, int W, int Ci, int KH, int KW, int PH,
                      int PW, float alpha, int accumulate, float* D, int64_t ldd, int fp16_operands, cudaStream_t st) {
  LPB_REQUIRE(Q > 0 && H > 0 && W > 0 && Ci > 0 && KH > 0 && KW > 0, "syrk_conv_patches: bad extents");
  LPB_REQUIRE(2 * PH == KH - 1 && 2 * PW == KW - 1, "syrk_conv_patches: stride-1 'same' convolutions only");
  LPB_REQUIRE((ldx % 8) == 0 && ldx >= Ci && ((uintptr_t)X_hi % 16) == 0 && ((uintptr_t)X_lo % 16) == 0,
              "syrk_conv_patches: operand rows must be 16-byte aligned");
  const int HW = H * W;
  tc::PatchGeom pg = {};
  // channel counts that are not a multiple of 64 are padded per tap: feature (t, ci) sits at t * Ci_pad + ci, the
  // TMA unit zero-fills channel coordinates >= Ci, so the padded rows / columns of D come out exactly zero
  const int Ci_pad = (int)ceil_div(Ci, 64) * 64;
  pg.KW = KW; pg.PH = PH; pg.PW = PW; pg.Ci = Ci; pg.blocks_per_tap = Ci_pad / 64;
  LPB_REQUIRE(KH * KW <= 16, "syrk_conv_patches: at most 16 taps");
  // taps whose window lies entirely in the padding (|shift| >= image extent: 3x3 kernels on 1x1 maps) contribute zero
  // rows / columns: only the `num_taps` live ones are computed, D is [num_taps * Ci_pad]^2 over the live taps in
  // ascending kernel position (live_taps() tells the caller which)
  pg.num_taps = 0;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw)
      if (kh - PH < H && PH - kh < H && kw - PW < W && PW - kw < W) pg.tap_of[pg.num_taps++] = (unsigned char)(kh * KW + kw);
  int box_h, box_n;
  if (HW >= 64) {
    LPB_REQUIRE(64 % W == 0 && H % (64 / W) == 0, "syrk_conv_patches: %dx%d images do not tile 64-row chunks", H, W);
    pg.rows_per_chunk = 64 / W; pg.chunks_per_img = HW / 64; pg.imgs_per_chunk = 0;
    box_h = pg.rows_per_chunk; box_n = 1;
  } else {
    LPB_REQUIRE(64 % HW == 0, "syrk_conv_patches: %dx%d images do not tile 64-row chunks", H, W);
    pg.rows_per_chunk = H; pg.chunks_per_img = 0; pg.imgs_per_chunk = 64 / HW;
    box_h = H; box_n = pg.imgs_per_chunk;
  }
  const int64_t d = (int64_t)pg.num_taps * Ci_pad;
  LPB_REQUIRE(ldd >= d, "syrk_conv_patches: ldd too small (D is [live_taps*Ci_pad]^2, Ci_pad = C_in rounded up to 64)");
  const int64_t kchunks64 = HW >= 64 ? Q * pg.chunks_per_img : ceil_div(Q, (int64_t)pg.imgs_per_chunk);
  LPB_REQUIRE(kchunks64 < (1LL << 31), "syrk_conv_patches: too many sample rows");
  const int total_kchunks = (int)kchunks64;
  const bool x3 = X_lo != nullptr;
  CUtensorMap tX_hi, tX_lo;
  if (make_tmap_patches(&tX_hi, X_hi, Q, H, W, Ci, ldx, box_h, box_n)) return 1;
  if (x3) {
    if (make_tmap_patches(&tX_lo, X_lo, Q, H, W, Ci, ldx, box_h, box_n)) return 1;
  } else {
    tX_lo = tX_hi;
  }
  const int tiles_m = (int)ceil_div(d, tc::BM);
  const int64_t tiles = (int64_t)tiles_m * (tiles_m + 1) / 2;
  const int sms = sm_count();
  int kps = 0, nsplit = 1;
  bool single = false;
  persistent_schedule(tiles, total_kchunks, sms, false, &kps, &nsplit, &single);
  if (!accumulate && check_cuda(cudaMemset2DAsync(D, ldd * sizeof(float), 0, d * sizeof(float), d, st), "syrk_conv_patches memset"))
    return 1;
  return launch_gemm_tc_persistent(true, x3, tX_hi, tX_lo, tX_hi, tX_lo, d, d, alpha, D, ldd, 1, tiles_m, tiles_m, tiles,
                                   total_kchunks, kps, nsplit, 0, fp16_operands, sms, st, &pg);
}

// Diagonal GGN / EF of a stride-1 'same' convolution weight on the tensor cores:
//     D[co, (t,ci)] (+)= alpha * sum_{q=(col,n)} ( sum_{h,w} G[(q,h,w), co] * x[n, h+kh-PH, w+kw-PW, ci] )^2
// G: output-gradient rows of all folded columns (bf16 hi/lo), x: NHWC activation rows of the Nimg images (same
// format).  One TMEM accumulation per sample (H*W/64 k-chunks), squared by the epilogue and summed over the CTA's
// samples in registers; tap-major columns with Ci padded to 64 (taps_to_param_rect() folds them into (ci,kh,kw)).
int diag_conv_sq(const void* G_hi, const void* G_lo, int64_t ldg, const void* X_hi, const void* X_lo, int64_t ldx, int64_t Qtot,
                 int64_t Nimg, int H, int W, int Ci, int Co, int KH, int KW, int PH, int PW, float alpha, int accumulate,
                 float* D, int64_t ldd, cudaStream_t st) {
  LPB_REQUIRE(Qtot > 0 && Nimg > 0 && Qtot % Nimg == 0, "diag_conv_sq: sample rows must be a multiple of the image count");
  LPB_REQUIRE(2 * PH == KH - 1 && 2 * PW == KW - 1, "diag_conv_sq: stride-1 'same' convolutions only");
  const int HW = H * W;
  LPB_REQUIRE(HW >= 64 && 64 % W == 0 && H % (64 / W) == 0, "diag_conv_sq: %dx%d images do not tile 64-row chunks", H, W);
  LPB_REQUIRE((ldx % 8) == 0 && ldx >= Ci && (ldg % 8) == 0 && ldg >= Co, "diag_conv_sq: bad leading dimensions");
  LPB_REQUIRE((G_lo == nullptr) == (X_lo == nullptr), "diag_conv_sq: lo operands must both be given or both NULL");
  const int Ci_pad = (int)ceil_div(Ci, 64) * 64;
  tc::PatchGeom pg = {};
  pg.KW = KW; pg.PH = PH; pg.PW = PW; pg.Ci = Ci; pg.num_taps = KH * KW; pg.blocks_per_tap = Ci_pad / 64;
  pg.rows_per_chunk = 64 / W; pg.chunks_per_img = HW / 64; pg.imgs_per_chunk = 0;
  pg.n_images = (int)Nimg; pg.group_chunks = pg.chunks_per_img;
  LPB_REQUIRE(KH * KW <= 16, "diag_conv_sq: at most 16 taps");
  for (int t = 0; t < KH * KW; ++t) pg.tap_of[t] = (unsigned char)t;
  LPB_REQUIRE(pg.group_chunks <= 64, "diag_conv_sq: images larger than 4096 pixels are not supported");
  const int64_t Ncols = (int64_t)KH * KW * Ci_pad;
  LPB_REQUIRE(ldd >= Ncols, "diag_conv_sq: ldd too small");
  const int64_t kchunks64 = Qtot * pg.chunks_per_img;
  LPB_REQUIRE(kchunks64 < (1LL << 31) && Nimg < (1LL << 31), "diag_conv_sq: too many sample rows");
  const bool x3 = G_lo != nullptr;
  CUtensorMap tG_hi, tG_lo, tX_hi, tX_lo;
  if (make_tmap_rows_ext(&tG_hi, G_hi, Qtot * HW, Co, ldg)) return 1;
  if (make_tmap_patches(&tX_hi, X_hi, Nimg, H, W, Ci, ldx, pg.rows_per_chunk, 1)) return 1;
  if (x3) {
    if (make_tmap_rows_ext(&tG_lo, G_lo, Qtot * HW, Co, ldg)) return 1;
    if (make_tmap_patches(&tX_lo, X_lo, Nimg, H, W, Ci, ldx, pg.rows_per_chunk, 1)) return 1;
  } else {
    tG_lo = tG_hi; tX_lo = tX_hi;
  }
  const int tiles_m = (int)ceil_div(Co, tc::BM), tiles_n = (int)ceil_div(Ncols, tc::BN);
  const int64_t tiles = (int64_t)tiles_m * tiles_n;
  const int sms = sm_count();
  // split over samples (a sample's chunks stay together): schedule in units of one sample
  int qps = 0, nsplit = 1;
  bool single = false;
  persistent_schedule(tiles, (int)imin(Qtot, 2147483647LL), sms, false, &qps, &nsplit, &single);
  if (!accumulate && check_cuda(cudaMemset2DAsync(D, ldd * sizeof(float), 0, Ncols * sizeof(float), Co, st), "diag_conv_sq memset"))
    return 1;
  return launch_gemm_tc_persistent(true, x3, tG_hi, tG_lo, tX_hi, tX_lo, Co, Ncols, alpha, D, ldd, 0, tiles_m, tiles_n, tiles,
                                   (int)kchunks64, qps * pg.chunks_per_img, nsplit, 0, 0, sms, st, &pg);
}

// out[co, ci*KK + t] += Dt[co, t*Cp + ci]
__global__ void __launch_bounds__(256) taps_to_param_rect_kernel(const float* __restrict__ Dt, int64_t ldt, int Co, int Ci, int Cp,
                                                                 int KK, float* __restrict__ out, int64_t ldo) {
  const int64_t total = (int64_t)Co * Ci * KK;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t co = e / (Ci * KK);
    const int r = (int)(e - co * Ci * KK);
    const int ci = r / KK, t = r - ci * KK;
    out[co * ldo + r] += Dt[co * ldt + (int64_t)t * Cp + ci];
  }
}

int taps_to_param_rect(const float* Dt, int64_t ldt, int Co, int Ci, int Ci_pad, int KK, float* out, int64_t ldo,
                       cudaStream_t st) {
  const int64_t total = (int64_t)Co * Ci * KK;
  if (total == 0) return 0;
  const int blocks = (int)imin(ceil_div(total, 256), (int64_t)sm_count() * 16);
  taps_to_param_rect_kernel<<<blocks, 256, 0, st>>>(Dt, ldt, Co, Ci, Ci_pad, KK, out, ldo);
  LPB_CHECK_LAUNCH("taps_to_param_rect");
  return 0;
}

// out[(ci*KK + t), (cj*KK + t')] += T[(t*Cp + ci), (t'*Cp + cj)]   (KK = KH*KW <= 9, Cp = padded channel stride of T).
// One CTA per (ci, 32 cj's): the KK x KK x 32 block is read in 128-byte runs, staged in shared memory and written as
// KK runs of 32*KK floats.
// number of kernel positions whose window overlaps a H x W image at all (the others only ever read zero padding)
int syrk_conv_live_taps(int KH, int KW, int PH, int PW, int H, int W) {
  int n = 0;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw)
      if (kh - PH < H && PH - kh < H && kw - PW < W && PW - kw < W) ++n;
  return n;
}

struct TapMap {
  int n;                 // live taps (rows / columns of T per channel block)
  unsigned char of[16];  // live index -> kernel position
};

__global__ void __launch_bounds__(256) taps_to_param_kernel(const float* __restrict__ T, int64_t ldt, int Ci, int Cp, int KK,
                                                            TapMap tm, float* __restrict__ out, int64_t ldo) {
  __shared__ float s[9 * 9 * 32];
  const int ci = blockIdx.y, cj0 = blockIdx.x * 32;
  const int ncj = min(32, Ci - cj0);
  const int L = tm.n;
  for (int e = threadIdx.x; e < L * L * 32; e += blockDim.x) {
    const int cjl = e & 31, tt = e >> 5;          // tt = i * L + i'  (live indices)
    const int t = tt / L, t2 = tt - t * L;
    if (cjl < ncj) s[e] = T[(int64_t)(t * Cp + ci) * ldt + t2 * Cp + cj0 + cjl];
  }
  __syncthreads();
  const int run = ncj * L;
  for (int e = threadIdx.x; e < L * run; e += blockDim.x) {
    const int t = e / run, r = e - t * run;
    const int cjl = r / L, t2 = r - cjl * L;
    float* dst = out + (int64_t)(ci * KK + tm.of[t]) * ldo + (int64_t)(cj0 + cjl) * KK + tm.of[t2];
    *dst += s[(t * L + t2) * 32 + cjl];
  }
}

int taps_to_param_accumulate(const float* T, int64_t ldt, int Ci, int Ci_pad, int KH, int KW, int PH, int PW, int H, int W,
                             float* out, int64_t ldo, cudaStream_t st) {
  const int KK = KH * KW;
  LPB_REQUIRE(Ci_pad >= Ci, "taps_to_param_accumulate: padded channel stride smaller than the channel count");
  TapMap tm = {};
  for (int kh = 0; kh < KH; ++kh)     // same live-tap rule as syrk_conv_patches
    for (int kw = 0; kw < KW; ++kw)
      if (kh - PH < H && PH - kh < H && kw - PW < W && PW - kw < W) tm.of[tm.n++] = (unsigned char)(kh * KW + kw);
  LPB_REQUIRE(ldt >= (int64_t)tm.n * Ci_pad, "taps_to_param_accumulate: ldt too small");
  LPB_REQUIRE(KK >= 1 && KK <= 9, "taps_to_param_accumulate: kernel window larger than 9 taps");
  LPB_REQUIRE(Ci > 0 && Ci <= 65535, "taps_to_param_accumulate: bad channel count");
  dim3 grid((unsigned)ceil_div(Ci, 32), (unsigned)Ci);
  taps_to_param_kernel<<<grid, 256, 0, st>>>(T, ldt, Ci, Ci_pad, KK, tm, out, ldo);
  LPB_CHECK_LAUNCH("taps_to_param_accumulate");
  return 0;
}

}  // namespace lpb
