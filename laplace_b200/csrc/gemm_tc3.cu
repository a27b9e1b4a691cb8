// Persistent variant of the factor contraction (gemm_tc.cu):  D[M,N] (fp32) += alpha * A * B^T
//
// One CTA per SM walks a static list of work items (output tile x K-range).  Within an item the K-range is cut into
// groups of <= 32 k-chunks (K = 2048: the TMEM accumulator truncates, longer chains show a measurable bias); groups
// alternate between two TMEM accumulators, the epilogue warps drain one while the tensor pipe fills the other and sum
// the group results in registers (round-to-nearest fp32).  An item therefore ends with ONE set of reductions into D
// however long its K-range is.  Compared with the one-tile-per-CTA kernel this
//   * hides the epilogue (TMEM -> registers -> red.global) and the per-CTA prologue (barrier init, TMEM alloc,
//     descriptor fetch) behind the MMAs of the next group / item, and
//   * divides the split-K reduction traffic by the number of groups per item: a d = 64 factor over 2.6 M rows went
//     through 1280 CTAs x 64 KiB of atomics onto the same 64 KiB of D; here it is 148 x 64 KiB.
#include "tc_common.cuh"

namespace lpb {

namespace tc {

constexpr int GROUP_CHUNKS = 32;      // k-chunks accumulated in one TMEM tile
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

template <int NPROD, bool MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                          const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, int M,
                          int N, float alpha, float* __restrict__ D, int64_t ldd, int symmetric, int tiles_m, int tiles_n,
                          int num_tiles, int num_items, int total_kchunks, int kchunks_per_split, int num_stages,
                          int store_mode, int fp16_operands) {
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
        for (int kc = it.kc_begin; kc < it.kc_end; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], it.diag ? STAGE_BYTES / 2 : STAGE_BYTES);
          auto load_tile = [&](const CUtensorMap* map, uint8_t* dst, int tile) {
            if (MN) {
              tma_load_2d(map, &full_bar[stage], dst, tile * BM, kc * BK);
              tma_load_2d(map, &full_bar[stage], dst + TILE_BYTES / 2, tile * BM + 64, kc * BK);
            } else {
              tma_load_2d(map, &full_bar[stage], dst, kc * BK, tile * BM);
            }
          };
          load_tile(&tmA_hi, st, it.tm);
          if (!it.diag) load_tile(&tmB_hi, st + TILE_BYTES, it.tn);
          if (NPROD == 3) {
            load_tile(&tmA_lo, st + 2 * TILE_BYTES, it.tm);
            if (!it.diag) load_tile(&tmB_lo, st + 3 * TILE_BYTES, it.tn);
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
        const int glen = (n + ngroups - 1) / ngroups;
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
      const int glen = (n + ngroups - 1) / ngroups;
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
          if (g0 == 0) {
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
void persistent_schedule(int64_t tiles, int total_kchunks, int ctas, bool allow_single_store, int* kchunks_per_split,
                         int* splits, bool* single) {
  const double EPI = 3.0;   // one item's output pass, in k-chunk equivalents (mostly hidden behind the next item)
  double best = 1e300, cost1 = 1e300;
  int best_s = 1;
  const int max_s = (int)imin(total_kchunks, 16384);
  for (int s = 1; s <= max_s; ++s) {
    const int kps = (int)ceil_div(total_kchunks, s);
    if ((int)ceil_div(total_kchunks, kps) != s) continue;
    const double waves = (double)ceil_div(tiles * s, (int64_t)ctas);
    const double cost = waves * (kps + EPI);
    if (s == 1) cost1 = cost;
    if (cost < best * 0.999) { best = cost; best_s = s; }
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
                              cudaStream_t st) {
  const int64_t items = num_tiles * splits;
  LPB_REQUIRE(items <= 2147483647LL, "gemm_tc_persistent: too many work items");
  const int stage_bytes = (x3 ? 4 : 2) * tc::TILE_BYTES;
  const int num_stages = x3 ? 3 : 6;
  const size_t smem = (size_t)num_stages * stage_bytes + (2 * num_stages + 4) * sizeof(uint64_t) + 16 + 1024;
  const unsigned grid = (unsigned)imin(items, ctas);
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
        (int)items, total_kchunks, kchunks_per_split, num_stages, store_mode, fp16_operands);                            \
  } while (0)
  if (x3 && mn) LPB_LAUNCH_P(3, true);
  else if (x3) LPB_LAUNCH_P(3, false);
  else if (mn) LPB_LAUNCH_P(1, true);
  else LPB_LAUNCH_P(1, false);
#undef LPB_LAUNCH_P
  LPB_CHECK_LAUNCH("gemm_tc_persistent");
  return 0;
}

}  // namespace lpb
