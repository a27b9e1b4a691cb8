// tcgen05 tensor-core contraction for the factor GEMMs (K1/K2/K3 of SURVEY 2.3):
//
//     D[M,N] (fp32) += alpha * A[M,K] * B[N,K]^T        A, B: bf16, K-major
//
// B200-native structure (no library GEMM):
//   * operands staged by TMA (cp.async.bulk.tensor.2d, 128B swizzle) into a multi-stage shared
//     memory ring guarded by full/empty mbarriers;
//   * one elected thread issues tcgen05.mma (cta_group::1, kind::f16, M=128 x N=128 x K=16) with the
//     fp32 accumulator resident in TMEM; tcgen05.commit releases ring slots / signals the epilogue;
//   * 4 epilogue warps read TMEM with tcgen05.ld and add the tile into the accumulated factor
//     buffer with vector fp32 reductions (split-K CTAs all reduce into the same buffer);
//   * SYRK mode visits only tiles on/above the diagonal and mirrors them;
//   * optional error-compensated mode (NPROD = 3): hi*hi + hi*lo + lo*hi with bf16 hi/lo splits,
//     all three products accumulated in the same TMEM tile (relative product error ~2^-16).
#include <stdlib.h>

#include "tc_common.cuh"

namespace lpb {

namespace tc {

// MN = false: operands K-major  (A[M, K], B[N, K], contraction index contiguous)          D = A B^T
// MN = true : operands MN-major (A[K, M], B[K, N], row-major "sample rows x features")   D = A^T B
//             -- the layout activations / gradients already have, so XᵀX needs no transposing pack.
template <int NPROD, bool MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_nt_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, int M, int N,
                  float alpha, float* __restrict__ D, int64_t ldd, int symmetric, int tiles_m, int tiles_n,
                  int total_kchunks, int kchunks_per_split, int num_stages, int store_mode, int fp16_operands) {
  constexpr int TILES_PER_STAGE = NPROD == 3 ? 4 : 2;
  constexpr int STAGE_BYTES = TILES_PER_STAGE * TILE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)num_stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + num_stages;
  uint64_t* tmem_full_bar = empty_bar + num_stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  // ---- tile / split decode (uniform per CTA) ----
  int tm, tn;
  if (symmetric) {
    int t = blockIdx.x, r = 0, cnt = tiles_m;
    while (t >= cnt) { t -= cnt; ++r; --cnt; }
    tm = r; tn = r + t;
  } else {
    tm = blockIdx.x / tiles_n; tn = blockIdx.x % tiles_n;
  }
  const bool diag = symmetric && (tm == tn);
  const int kc_begin = blockIdx.y * kchunks_per_split;
  const int kc_end = min(total_kchunks, kc_begin + kchunks_per_split);
  if (kc_begin >= kc_end) return;

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
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
        // diagonal SYRK tiles use the A tile for both operands: half the TMA traffic
        mbar_expect_tx(&full_bar[stage], diag ? STAGE_BYTES / 2 : STAGE_BYTES);
        auto load_tile = [&](const CUtensorMap* map, uint8_t* dst, int tile) {
          if (MN) {  // two boxes of 64 features x 64 rows (128 B per row, 8 KiB each)
            tma_load_2d(map, &full_bar[stage], dst, tile * BM, kc * BK);
            tma_load_2d(map, &full_bar[stage], dst + TILE_BYTES / 2, tile * BM + 64, kc * BK);
          } else {
            tma_load_2d(map, &full_bar[stage], dst, kc * BK, tile * BM);
          }
        };
        load_tile(&tmA_hi, st, tm);
        if (!diag) load_tile(&tmB_hi, st + TILE_BYTES, tn);
        if (NPROD == 3) {
          load_tile(&tmA_lo, st + 2 * TILE_BYTES, tm);
          if (!diag) load_tile(&tmB_lo, st + 3 * TILE_BYTES, tn);
        }
        if (++stage == num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, BN, fp16_operands, MN ? 1 : 0);
      auto mk = [](uint32_t saddr) { return MN ? make_smem_desc_mn(saddr, TILE_BYTES / 2) : make_smem_desc(saddr); };
      int stage = 0; uint32_t phase = 0; uint32_t acc = 0;
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint64_t a_hi = mk(sbase), b_hi = diag ? a_hi : mk(sbase + TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: 16 elements = 32 B along the swizzled row; MN-major: 16 rows = two 8-row groups = 2048 B
          const uint64_t koff = (uint64_t)((MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
          umma_f16(tmem_base, a_hi + koff, b_hi + koff, idesc, acc);
          acc = 1;
          if (NPROD == 3) {
            const uint64_t a_lo = mk(sbase + 2 * TILE_BYTES);
            const uint64_t b_lo = diag ? a_lo : mk(sbase + 3 * TILE_BYTES);
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
    // ================= epilogue: TMEM -> registers -> fp32 reductions into D =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int row = tm * BM + q * 32 + lane;
    const bool mirror = symmetric && (tm != tn);
    const bool vec_ok = ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
#pragma unroll 1
    for (int chunk = 0; chunk < BN / 32; ++chunk) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(chunk * 32), v);
      const int col0 = tn * BN + chunk * 32;
      if (row < M) {
        float* drow = D + (int64_t)row * ldd + col0;
        if (store_mode) {
          // single split, overwrite semantics: plain (vector) stores, no atomics, no pre-zeroing
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
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
PFN_encodeTiled get_tensormap_encoder() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// row-major [rows, cols] matrix, boxes of 64 columns (128 B) x 64 rows: the MN-major operand tiles
static int make_tmap_rows(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld) {
  PFN_encodeTiled enc = get_tensormap_encoder();
  LPB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)tc::BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LPB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rows) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
              (long long)rows, (long long)cols, (long long)ld);
  return 0;
}

int make_tmap_2d(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  PFN_encodeTiled enc = get_tensormap_encoder();
  LPB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LPB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld ld=%lld", (int)r, (long long)rows,
              (long long)K, (long long)ld);
  return 0;
}

static int gemm_tc(bool mn, const void* A_hi, const void* A_lo, int64_t lda, const void* B_hi, const void* B_lo, int64_t ldb,
                   int64_t M, int64_t N, int64_t K, float alpha, int accumulate, float* D, int64_t ldd, int symmetric,
                   int fp16_operands, cudaStream_t st);
int launch_gemm_tc_pair(bool mn, bool x3, const CUtensorMap& tA_hi, const CUtensorMap& tA_lo, const CUtensorMap& tB_hi,
                        const CUtensorMap& tB_lo, int64_t M, int64_t N, float alpha, float* D, int64_t ldd, int symmetric,
                        int total_kchunks, int kchunks_per_split, int splits, int store_mode, int fp16_operands,
                        cudaStream_t st);

// Schedules: persistent CTAs with double-buffered TMEM (gemm_tc3.cu) by default -- measured 10-40 % faster than one
// tile per CTA on every factor shape of the benchmark; 128 x 128 one-tile CTAs (this file) and 256 x 256 CTA-pair
// tiles (gemm_tc2.cu) stay selectable through LPB_GEMM_MODE / lpb_set_gemm_tile_mode() for tests and A/B timing.
void persistent_schedule(int64_t tiles, int total_kchunks, int ctas, bool allow_single_store, int* kchunks_per_split,
                         int* splits, bool* single);
int launch_gemm_tc_persistent(bool mn, bool x3, const CUtensorMap& tA_hi, const CUtensorMap& tA_lo, const CUtensorMap& tB_hi,
                              const CUtensorMap& tB_lo, int64_t M, int64_t N, float alpha, float* D, int64_t ldd,
                              int symmetric, int tiles_m, int tiles_n, int64_t num_tiles, int total_kchunks,
                              int kchunks_per_split, int splits, int store_mode, int fp16_operands, int ctas,
                              cudaStream_t st, const tc::PatchGeom* patches);

// -2: read LPB_GEMM_MODE on first use; -1 auto; 0 one 128x128 tile per CTA; 1 CTA-pair tiles whenever M, N >= 256;
// 2 persistent CTAs with double-buffered TMEM (gemm_tc3.cu)
static int g_pair_mode = -2;
void set_gemm_pair_mode(int mode) { g_pair_mode = mode < 0 ? -1 : imin(mode, 2); }
static int gemm_mode() {
  if (g_pair_mode == -2) {
    const char* e = getenv("LPB_GEMM_MODE");
    g_pair_mode = e ? (int)imin(imax(atoi(e), -1), 2) : -1;
  }
  return g_pair_mode;
}
static bool use_pair_tiles(int64_t M, int64_t N) { return gemm_mode() == 1 && M >= 256 && N >= 256; }

int gemm_nt_bf16(const void* A_hi, const void* A_lo, int64_t lda, const void* B_hi, const void* B_lo, int64_t ldb,
                 int64_t M, int64_t N, int64_t K, float alpha, int accumulate, float* D, int64_t ldd, int symmetric,
                 int fp16_operands, cudaStream_t st) {
  return gemm_tc(false, A_hi, A_lo, lda, B_hi, B_lo, ldb, M, N, K, alpha, accumulate, D, ldd, symmetric, fp16_operands, st);
}

int gemm_tn_rows(const void* A_hi, const void* A_lo, int64_t lda, const void* B_hi, const void* B_lo, int64_t ldb,
                 int64_t M, int64_t N, int64_t K, float alpha, int accumulate, float* D, int64_t ldd, int symmetric,
                 int fp16_operands, cudaStream_t st) {
  return gemm_tc(true, A_hi, A_lo, lda, B_hi, B_lo, ldb, M, N, K, alpha, accumulate, D, ldd, symmetric, fp16_operands, st);
}

static int gemm_tc(bool mn, const void* A_hi, const void* A_lo, int64_t lda, const void* B_hi, const void* B_lo, int64_t ldb,
                   int64_t M, int64_t N, int64_t K, float alpha, int accumulate, float* D, int64_t ldd, int symmetric,
                   int fp16_operands, cudaStream_t st) {
  LPB_REQUIRE(!symmetric || M == N, "gemm_nt_bf16: symmetric needs M == N");
  // TMA coordinates are 32-bit: the contraction index (sample rows) and the feature extents must fit
  LPB_REQUIRE(K < (1LL << 31) - 64 && M < (1LL << 31) - 256 && N < (1LL << 31) - 256,
              "gemm_nt_bf16: extent beyond 32-bit tensor-map coordinates (M=%lld N=%lld K=%lld)", (long long)M, (long long)N,
              (long long)K);
  LPB_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0, "gemm_nt_bf16: leading dimensions must be multiples of 8 elements");
  LPB_REQUIRE(((uintptr_t)A_hi % 16) == 0 && ((uintptr_t)B_hi % 16) == 0 && ((uintptr_t)A_lo % 16) == 0 &&
                  ((uintptr_t)B_lo % 16) == 0,
              "gemm_nt_bf16: operands must be 16-byte aligned");
  if (M == 0 || N == 0) return 0;
  const bool pair = use_pair_tiles(M, N);
  const int tile_edge = pair ? 256 : tc::BM;
  const int tiles_m = (int)ceil_div(M, tile_edge), tiles_n = (int)ceil_div(N, tile_edge);
  // CTAs per split: one per 128 x 128 tile, or two per 256 x 256 pair tile
  const int64_t tiles = (symmetric ? (int64_t)tiles_m * (tiles_m + 1) / 2 : (int64_t)tiles_m * tiles_n) * (pair ? 2 : 1);
  const int total_kchunks = (int)ceil_div(K, tc::BK);
  const int sms = sm_count();
  if ((gemm_mode() == 2 || gemm_mode() == -1) && K > 0) {
    int kps = 0, nsplit = 1;
    bool single = false;
    persistent_schedule(tiles, total_kchunks, sms, !accumulate && !symmetric, &kps, &nsplit, &single);
    const bool store = !accumulate && !symmetric && single;
    if (!accumulate && !store &&
        check_cuda(cudaMemset2DAsync(D, ldd * sizeof(float), 0, N * sizeof(float), M, st), "gemm_tc memset"))
      return 1;
    const bool x3p = A_lo != nullptr;
    CUtensorMap pA_hi, pA_lo, pB_hi, pB_lo;
    auto mk = [&](CUtensorMap* m, const void* p, int64_t feat, int64_t ld) {
      return mn ? make_tmap_rows(m, p, K, feat, ld) : make_tmap_2d(m, p, feat, K, ld);
    };
    if (mk(&pA_hi, A_hi, M, lda) || mk(&pB_hi, B_hi, N, ldb)) return 1;
    if (x3p) {
      if (mk(&pA_lo, A_lo, M, lda) || mk(&pB_lo, B_lo, N, ldb)) return 1;
    } else {
      pA_lo = pA_hi; pB_lo = pB_hi;
    }
    return launch_gemm_tc_persistent(mn, x3p, pA_hi, pA_lo, pB_hi, pB_lo, M, N, alpha, D, ldd, symmetric, tiles_m, tiles_n,
                                     tiles, total_kchunks, kps, nsplit, store ? 1 : 0, fp16_operands, sms, st, nullptr);
  }
  // overwrite + enough tiles (or a short K): one split per tile with plain stores
  const bool store_mode = !accumulate && !symmetric && K > 0 && (tiles >= sms / 2 || total_kchunks <= 16) && total_kchunks <= 128;
  if (!accumulate && !store_mode) {
    if (check_cuda(cudaMemset2DAsync(D, ldd * sizeof(float), 0, N * sizeof(float), M, st), "gemm_nt_bf16 memset"))
      return 1;
  }
  if (K == 0) return 0;
  const bool x3 = A_lo != nullptr;
  CUtensorMap tA_hi, tA_lo, tB_hi, tB_lo;
  auto mkmap = [&](CUtensorMap* m, const void* p, int64_t feat, int64_t ld) {
    return mn ? make_tmap_rows(m, p, K, feat, ld) : make_tmap_2d(m, p, feat, K, ld);
  };
  if (mkmap(&tA_hi, A_hi, M, lda) || mkmap(&tB_hi, B_hi, N, ldb)) return 1;
  if (x3) {
    if (mkmap(&tA_lo, A_lo, M, lda) || mkmap(&tB_lo, B_lo, N, ldb)) return 1;
  } else {
    tA_lo = tA_hi; tB_lo = tB_hi;
  }
  // split K so that about two CTAs per SM exist while each CTA keeps >= 4 k-chunks; never accumulate more
  // than 32 chunks (K = 2048) in one TMEM tile: the tensor-core accumulator truncates, and the bias grows
  // linearly with the number of chained MMAs (measured 4.5e-5 relative at K = 8192) -- split-K partial sums are
  // combined with round-to-nearest fp32 reductions instead.
  int64_t splits = imax(1, (2 * (int64_t)sms) / tiles);
  splits = imin(splits, imax(1, total_kchunks / 4));
  splits = imax(splits, ceil_div(total_kchunks, 32));
  splits = imin(splits, 65535);
  if (store_mode) splits = 1;
  const int kchunks_per_split = (int)ceil_div(total_kchunks, splits);
  splits = ceil_div(total_kchunks, kchunks_per_split);
  const int stage_bytes = (x3 ? 4 : 2) * tc::TILE_BYTES;
  const int num_stages = x3 ? 3 : 6;
  const size_t smem = (size_t)num_stages * stage_bytes + (2 * num_stages + 1) * sizeof(uint64_t) + 16 + 1024;
  LPB_REQUIRE(tiles <= 2147483647LL, "gemm_tc: too many tiles");
  if (pair)
    return launch_gemm_tc_pair(mn, x3, tA_hi, tA_lo, tB_hi, tB_lo, M, N, alpha, D, ldd, symmetric, total_kchunks,
                               kchunks_per_split, (int)splits, store_mode ? 1 : 0, fp16_operands, st);
  dim3 grid((unsigned)tiles, (unsigned)splits);
#define LPB_LAUNCH_TC(NP, MNV)                                                                                          \
  do {                                                                                                                  \
    static bool attr_done = false;                                                                                      \
    if (!attr_done) {                                                                                                   \
      if (check_cuda(cudaFuncSetAttribute(tc::gemm_nt_tc_kernel<NP, MNV>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          227 * 1024),                                                                  \
                     "gemm_tc attr"))                                                                                   \
        return 1;                                                                                                       \
      attr_done = true;                                                                                                 \
    }                                                                                                                   \
    tc::gemm_nt_tc_kernel<NP, MNV><<<grid, tc::NUM_THREADS, smem, st>>>(                                                \
        tA_hi, tA_lo, tB_hi, tB_lo, (int)M, (int)N, alpha, D, ldd, symmetric, tiles_m, tiles_n, total_kchunks,          \
        kchunks_per_split, num_stages, store_mode ? 1 : 0, fp16_operands);                                              \
  } while (0)
  if (x3 && mn) LPB_LAUNCH_TC(3, true);
  else if (x3) LPB_LAUNCH_TC(3, false);
  else if (mn) LPB_LAUNCH_TC(1, true);
  else LPB_LAUNCH_TC(1, false);
#undef LPB_LAUNCH_TC
  LPB_CHECK_LAUNCH("gemm_nt_bf16");
  return 0;
}

}  // namespace lpb
