// Device-side building blocks shared by the tcgen05 kernels (gemm_tc.cu, conv_tc.cu): mbarrier / TMA /
// tcgen05 wrappers (inline PTX), shared-memory and instruction descriptors, TMEM loads, vector reductions.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace lpb {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 64;        // BK * 2 B = 128 B = one swizzle row
constexpr int TILE_BYTES = BM * BK * 2;           // 16 KiB per operand tile
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;                  // warp0 TMA, warp1 MMA/TMEM, warps2-5 epilogue
constexpr int TMEM_COLS = 128;

// Implicit convolution-patch operand of the persistent contraction (gemm_tc3.cu): feature f = tap * Ci + ci of sample
// row (n,h,w) is x[n, h + kh - PH, w + kw - PW, ci] (zero outside the image), fetched as a shifted 4-D TMA box.
struct PatchGeom {
  int KW, PH, PW, Ci, num_taps, blocks_per_tap, chunks_per_img, rows_per_chunk, imgs_per_chunk;
  unsigned char tap_of[16];   // live tap i -> kernel position kh * KW + kw (taps that only read padding are dropped)
  int n_images;     // loader 3: sample rows (col, n) wrap around the images every n_images samples (0: no wrap)
  int group_chunks; // > 0: TMEM groups of exactly this many k-chunks (= one sample), SQUARED before they are summed
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// shared-memory matrix descriptor: K-major tile, 128B swizzle, rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address        bits [0,14)
  d |= (uint64_t)1 << 16;                          // leading byte offset  (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset   bits [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // layout: SWIZZLE_128B
  return d;
}

// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=128
// fp16_operands != 0 selects F16 (format 0) instead of BF16 (format 1) for A and B; mn_major != 0 marks both
// operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int fp16_operands = 0, int mn_major = 0) {
  return (1u << 4) | ((fp16_operands ? 0u : 1u) << 7) | ((fp16_operands ? 0u : 1u) << 10) | ((mn_major ? 1u : 0u) << 15) |
         ((mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// MN-major tile, 128B swizzle: canonical layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in 16-bit elements --
// 64 features contiguous (128 B), successive sample rows 128 B apart, 8-row groups SBO = 1024 B apart,
// 64-feature blocks LBO = `mn_block_bytes` apart (one TMA box).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t mn_block_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((mn_block_bytes >> 4) & 0x3FFF) << 16;   // leading byte offset
  d |= (uint64_t)(1024 >> 4) << 32;                        // stride byte offset
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}


__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

}  // namespace tc

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_tensormap_encoder();
int make_tmap_2d(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows = 128);

}  // namespace lpb
