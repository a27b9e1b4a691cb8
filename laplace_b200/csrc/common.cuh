// Shared helpers for the laplace_b200 native library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace lpb {

// thread-local last error string surfaced through lpb_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

inline int check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

#define LPB_CHECK_LAUNCH(what)                                 \
  do {                                                         \
    if (lpb::check_cuda(cudaGetLastError(), what)) return 1;   \
  } while (0)

#define LPB_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      lpb::set_error(__VA_ARGS__);    \
      return 1;                       \
    }                                 \
  } while (0)

__host__ __device__ inline int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ inline int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }
__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// output element kinds of the pack kernels
enum OutKind : int { OUT_F32 = 0, OUT_BF16 = 1, OUT_BF16_HILO = 2, OUT_F16_HILO = 3 };
enum PackFlags : int { PACK_SQUARE = 1 };

// store one packed value `v` at element index `idx` of the K-major staging buffers
template <int KIND>
__device__ __forceinline__ void store_packed(void* hi, void* lo, int64_t idx, float v) {
  if constexpr (KIND == OUT_F32) {
    reinterpret_cast<float*>(hi)[idx] = v;
  } else if constexpr (KIND == OUT_F16_HILO) {
    // fp16 hi + fp16 lo: 22 significant bits (|x - hi - lo| <= 2^-23 |x| while lo stays normal); callers keep
    // operands inside fp16 range (activations / weights); gradients use the bf16 split (8-bit exponent)
    const __half h = __float2half_rn(v);
    reinterpret_cast<__half*>(hi)[idx] = h;
    reinterpret_cast<__half*>(lo)[idx] = __float2half_rn(v - __half2float(h));
  } else {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    reinterpret_cast<__nv_bfloat16*>(hi)[idx] = h;
    if constexpr (KIND == OUT_BF16_HILO) {
      reinterpret_cast<__nv_bfloat16*>(lo)[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

int sm_count();

struct ConvGeom {
  int N, C, H, W, KH, KW, SH, SW, PH, PW, DH, DW, OH, OW;
};

}  // namespace lpb
