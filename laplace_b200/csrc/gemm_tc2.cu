// CTA-pair (cta_group::2) variant of the factor contraction in gemm_tc.cu:
//
//     D[M,N] (fp32) += alpha * A * B^T     one 256 x 256 output tile per pair of CTAs (a 2-CTA cluster on one TPC)
//
// Why: with 128 x 128 tiles and the error-compensated three-product mode every SM has to pull 64 KiB of operands from
// L2 per 64-deep k-chunk while its tensor pipe needs only ~770 clocks for the chunk's MMAs -- 85 B/clk against the
// ~42 B/clk one SM sustains from L2, so the pipe sits at ~50 % (profiles/r01_syrk_mn_major_ncu_run15.md).  A CTA pair
// computes a 256 x 256 tile from the same 64 KiB per SM: each CTA stages 128 rows of A and 128 rows of B, the
// leader issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) which reads A from the CTA's own shared memory and
// the two B halves from both, and each CTA's TMEM holds its 128 accumulator rows.  Bytes per flop halve.
//
// Protocol (per ring stage):
//   both producers : wait local empty[s]  ->  TMA (cta_group::2) into local smem, completing on the LEADER's full[s]
//   leader         : arrive.expect_tx(full[s], bytes of both CTAs); MMA thread waits full[s], issues the MMAs, then
//                    tcgen05.commit.cta_group::2 ... multicast -> empty[s] of both CTAs
//   after the last chunk the commit multicasts to tmem_full of both CTAs; each CTA's epilogue drains its own TMEM.
#include "tc_common.cuh"

namespace lpb {

namespace tc {

constexpr int TMEM_COLS2 = 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(rank));
  return out;
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t leader_bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

template <int NPROD, bool MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, int M, int N,
                    float alpha, float* __restrict__ D, int64_t ldd, int symmetric, int tiles_m, int tiles_n,
                    int total_kchunks, int kchunks_per_split, int num_stages, int store_mode, int fp16_operands) {
  constexpr int TILES_PER_STAGE = NPROD == 3 ? 4 : 2;
  constexpr int STAGE_BYTES = TILES_PER_STAGE * TILE_BYTES;   // per CTA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)num_stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + num_stages;
  uint64_t* tmem_full_bar = empty_bar + num_stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  // ---- pair tile / split decode (uniform over the pair) ----
  const int pair = blockIdx.x >> 1;
  int tm, tn;
  if (symmetric) {
    int t = pair, r = 0, cnt = tiles_m;
    while (t >= cnt) { t -= cnt; ++r; --cnt; }
    tm = r; tn = r + t;
  } else {
    tm = pair / tiles_n; tn = pair % tiles_n;
  }
  const bool diag = symmetric && (tm == tn);
  const int kc_begin = blockIdx.y * kchunks_per_split;
  const int kc_end = min(total_kchunks, kc_begin + kchunks_per_split);
  if (kc_begin >= kc_end) return;   // same decision in both CTAs of the pair

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS2)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised, TMEM address published
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // feature rows staged by this CTA: its half of the A tile and its half of the B tile
  const int a_row0 = tm * 256 + (int)rank * 128;
  const int b_row0 = tn * 256 + (int)rank * 128;

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t pair_bytes = 2u * (diag ? STAGE_BYTES / 2 : STAGE_BYTES);
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
        const uint32_t lbar = mapa_rank(smem_u32(&full_bar[stage]), 0);
        if (leader) mbar_expect_tx(&full_bar[stage], pair_bytes);
        auto load_tile = [&](const CUtensorMap* map, uint8_t* dst, int row0) {
          if (MN) {
            tma_load_2d_pair(map, lbar, dst, row0, kc * BK);
            tma_load_2d_pair(map, lbar, dst + TILE_BYTES / 2, row0 + 64, kc * BK);
          } else {
            tma_load_2d_pair(map, lbar, dst, kc * BK, row0);
          }
        };
        load_tile(&tmA_hi, st, a_row0);
        if (!diag) load_tile(&tmB_hi, st + TILE_BYTES, b_row0);
        if (NPROD == 3) {
          load_tile(&tmA_lo, st + 2 * TILE_BYTES, a_row0);
          if (!diag) load_tile(&tmB_lo, st + 3 * TILE_BYTES, b_row0);
        }
        if (++stage == num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc(256, 256, fp16_operands, MN ? 1 : 0);
      auto mk = [](uint32_t saddr) { return MN ? make_smem_desc_mn(saddr, TILE_BYTES / 2) : make_smem_desc(saddr); };
      int stage = 0; uint32_t phase = 0; uint32_t acc = 0;
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint64_t a_hi = mk(sbase), b_hi = diag ? a_hi : mk(sbase + TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t koff = (uint64_t)((MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
          umma_f16_pair(tmem_base, a_hi + koff, b_hi + koff, idesc, acc);
          acc = 1;
          if (NPROD == 3) {
            const uint64_t a_lo = mk(sbase + 2 * TILE_BYTES);
            const uint64_t b_lo = diag ? a_lo : mk(sbase + 3 * TILE_BYTES);
            umma_f16_pair(tmem_base, a_hi + koff, b_lo + koff, idesc, 1);
            umma_f16_pair(tmem_base, a_lo + koff, b_hi + koff, idesc, 1);
          }
        }
        umma_commit_pair(&empty_bar[stage]);
        if (++stage == num_stages) { stage = 0; phase ^= 1; }
      }
      umma_commit_pair(tmem_full_bar);
    }
  } else {
    // ================= epilogue (both CTAs: 128 accumulator rows each, 256 columns) =================
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int row = a_row0 + q * 32 + lane;
    const bool mirror = symmetric && (tm != tn);
    const bool vec_ok = ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
#pragma unroll 1
    for (int chunk = 0; chunk < 256 / 32; ++chunk) {
      const int col0 = tn * 256 + chunk * 32;
      if (col0 >= N) break;   // warp-uniform
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(chunk * 32), v);
      if (row < M) {
        float* drow = D + (int64_t)row * ldd + col0;
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
    tcgen05_fence_before();
  }
  // neither CTA may exit (or free TMEM) while its peer can still read its shared memory / signal its barriers
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS2) : "memory");
  }
}

}  // namespace tc

// Launch the pair kernel.  Tensor maps are built by the caller (gemm_tc.cu) with 128-row (K-major) or 64 x 64
// (MN-major) boxes -- the same maps the single-CTA kernel uses.
int launch_gemm_tc_pair(bool mn, bool x3, const CUtensorMap& tA_hi, const CUtensorMap& tA_lo, const CUtensorMap& tB_hi,
                        const CUtensorMap& tB_lo, int64_t M, int64_t N, float alpha, float* D, int64_t ldd, int symmetric,
                        int total_kchunks, int kchunks_per_split, int splits, int store_mode, int fp16_operands,
                        cudaStream_t st) {
  const int tiles_m = (int)ceil_div(M, 256), tiles_n = (int)ceil_div(N, 256);
  const int64_t pairs = symmetric ? (int64_t)tiles_m * (tiles_m + 1) / 2 : (int64_t)tiles_m * tiles_n;
  LPB_REQUIRE(2 * pairs <= 2147483647LL, "gemm_tc_pair: too many tiles");
  const int stage_bytes = (x3 ? 4 : 2) * tc::TILE_BYTES;
  const int num_stages = x3 ? 3 : 6;
  const size_t smem = (size_t)num_stages * stage_bytes + (2 * num_stages + 1) * sizeof(uint64_t) + 16 + 1024;
  dim3 grid((unsigned)(2 * pairs), (unsigned)splits);
#define LPB_LAUNCH_PAIR(NP, MNV)                                                                                         \
  do {                                                                                                                   \
    static bool attr_done = false;                                                                                       \
    if (!attr_done) {                                                                                                    \
      if (check_cuda(cudaFuncSetAttribute(tc::gemm_tc_pair_kernel<NP, MNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          227 * 1024),                                                                   \
                     "gemm_tc_pair attr"))                                                                               \
        return 1;                                                                                                        \
      attr_done = true;                                                                                                  \
    }                                                                                                                    \
    tc::gemm_tc_pair_kernel<NP, MNV><<<grid, tc::NUM_THREADS, smem, st>>>(                                               \
        tA_hi, tA_lo, tB_hi, tB_lo, (int)M, (int)N, alpha, D, ldd, symmetric, tiles_m, tiles_n, total_kchunks,           \
        kchunks_per_split, num_stages, store_mode, fp16_operands);                                                       \
  } while (0)
  if (x3 && mn) LPB_LAUNCH_PAIR(3, true);
  else if (x3) LPB_LAUNCH_PAIR(3, false);
  else if (mn) LPB_LAUNCH_PAIR(1, true);
  else LPB_LAUNCH_PAIR(1, false);
#undef LPB_LAUNCH_PAIR
  LPB_CHECK_LAUNCH("gemm_tc_pair");
  return 0;
}

}  // namespace lpb
