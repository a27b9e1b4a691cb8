// C ABI of laplace_b200 (see include/laplace_b200.h for the contract of every entry point).
#include <stdarg.h>

#include "../../include/laplace_b200.h"
#include "common.cuh"

namespace lpb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}


int pack_rows_t(const float*, int64_t, int64_t, int64_t, const float*, int, float, int, void*, void*, int, int64_t,
                int64_t, cudaStream_t);
int pack_conv2d_t(const float*, const ConvGeom&, float, int, int, void*, void*, int, int64_t, int64_t, cudaStream_t);
int pack_nchw_t(const float*, int64_t, int, int, float, int, int, void*, void*, int, int64_t, int64_t, cudaStream_t);
int pack_conv2d_rows(const float*, const ConvGeom&, void*, void*, int, int64_t, cudaStream_t);
int pack_nchw_rows(const float*, int64_t, int, int, void*, void*, int, int64_t, cudaStream_t);
int pack_cast(const float*, int64_t, int64_t, int64_t, void*, void*, int, int64_t, cudaStream_t);
int col2im(const float*, int64_t, const ConvGeom&, float*, cudaStream_t);
int gemm_nt_f32(const float*, int64_t, const float*, int64_t, int64_t, int64_t, int64_t, float, int, float*, int64_t, int,
                cudaStream_t);
int gemm_nt_bf16(const void*, const void*, int64_t, const void*, const void*, int64_t, int64_t, int64_t, int64_t, float,
                 int, float*, int64_t, int, int, cudaStream_t);
int scale_channels(const float*, const float*, float*, int64_t, int, int64_t, cudaStream_t);
int relu_bwd(const float*, const float*, float*, int64_t, int, cudaStream_t);
void set_mask_major_min(int64_t);
int maxpool2d_bwd(const float*, const int64_t*, float*, int64_t, int, int, int, int, int, int, int, int, int, cudaStream_t);
int maxpool2d_bwd_pack_nhwc(const float*, const int64_t*, const float*, const float*, void*, void*, int64_t, int64_t, int, int, int,
                            int, int, int, int, int, int, cudaStream_t);
int gemm_tn_rows(const void*, const void*, int64_t, const void*, const void*, int64_t, int64_t, int64_t, int64_t, float,
                 int, float*, int64_t, int, int, cudaStream_t);
int conv_nhwc_bf16(const void*, const void*, int64_t, int, int, int64_t, int64_t, const void*, const void*, int64_t, int, int,
                   int, int, int, int, float, float*, int64_t, int, cudaStream_t);
int shared_weight_contract(int, const float*, int64_t, const float*, int64_t, int, int, int, int, int, float, float*,
                           int64_t, int64_t, int64_t, cudaStream_t);
int jac_linear_write(const float*, const float*, int, int, int, int, float*, int64_t, int64_t, int64_t, int64_t,
                     cudaStream_t);
int ll_jacobian_write(const float*, int, int, int, int, float*, cudaStream_t);
int batched_pair_dot(const float*, const float*, const float*, int64_t, int, int, int, int, int64_t, int64_t, int64_t,
                     int64_t, int, float*, cudaStream_t);
int ll_ggn_expand(const float*, int, int, int, int, float*, cudaStream_t);
int ll_sigma_gather(const float*, int, int, int, float*, cudaStream_t);
int eigh_jacobi(const float*, int, int, float*, float*, int, cudaStream_t);
int kron_conv_quadform(const float*, int64_t, int64_t, const float*, int64_t, int, int, int, int, int, const float*, const float*,
                       float, int, float*, cudaStream_t);

void set_gemm_pair_mode(int mode);
int syrk_conv_patches(const void*, const void*, int64_t, int64_t, int, int, int, int, int, int, int, float, int, float*, int64_t,
                      int, cudaStream_t);
int taps_to_param_accumulate(const float*, int64_t, int, int, int, int, int, int, int, int, float*, int64_t, cudaStream_t);
int syrk_conv_live_taps(int KH, int KW, int PH, int PW, int H, int W);
int diag_conv_sq(const void*, const void*, int64_t, const void*, const void*, int64_t, int64_t, int64_t, int, int, int, int, int,
                 int, int, int, float, int, float*, int64_t, cudaStream_t);
int taps_to_param_rect(const float*, int64_t, int, int, int, int, float*, int64_t, cudaStream_t);
int conv_bwd_strided(const void*, const void*, int64_t, int, int, int64_t, int64_t, const void*, const void*, int64_t, int, int, int,
                     int, int, int, int, int, int, float*, int64_t, cudaStream_t);
int pack_cast_fused(const float*, int64_t, int64_t, int64_t, const float*, const float*, int64_t, int64_t, void*, void*, int,
                    int64_t, cudaStream_t);
int col2im_nhwc(const float*, int64_t, const ConvGeom&, float*, cudaStream_t);
int maxpool2d_bwd_nhwc(const float*, const int64_t*, float*, int64_t, int, int, int, int, int, int, int, int, int, cudaStream_t);
}  // namespace lpb

using lpb::set_error;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int lpb_version(void) { return 100; }

const char* lpb_last_error(void) { return lpb::get_error(); }

int lpb_set_gemm_tile_mode(int mode) {
  lpb::set_gemm_pair_mode(mode);
  return 0;
}

int lpb_set_mask_major_min(int64_t min_mask_elems) {
  lpb::set_mask_major_min(min_mask_elems);
  return 0;
}

int lpb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  if (lpb::check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
  int sms = 0, maj = 0, min = 0;
  if (lpb::check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "attr sm count")) return 1;
  if (lpb::check_cuda(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev), "attr cc major")) return 1;
  if (lpb::check_cuda(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev), "attr cc minor")) return 1;
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return 0;
}

int lpb_pack_rows_t(const float* src, int64_t rows, int64_t cols, int64_t ld_src, const float* row_scale, int nrep,
                    float scale, int flags, void* dst_hi, void* dst_lo, int out_kind, int64_t ldk, int64_t k0,
                    void* stream) {
  LPB_REQUIRE(rows >= 0 && cols >= 0 && ld_src >= cols, "lpb_pack_rows_t: bad extents");
  LPB_REQUIRE(k0 >= 0 && k0 + rows <= ldk, "lpb_pack_rows_t: rows [%lld, %lld) exceed ldk=%lld", (long long)k0,
              (long long)(k0 + rows), (long long)ldk);
  return lpb::pack_rows_t(src, rows, cols, ld_src, row_scale, nrep, scale, flags, dst_hi, dst_lo, out_kind, ldk, k0,
                          ST(stream));
}

int lpb_pack_conv2d_t(const float* x, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW, int DH,
                      int DW, float scale, int flags, int reduce_mean, void* dst_hi, void* dst_lo, int out_kind,
                      int64_t ldk, int64_t k0, void* stream) {
  LPB_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && KH > 0 && KW > 0 && SH > 0 && SW > 0 && DH > 0 && DW > 0,
              "lpb_pack_conv2d_t: bad geometry");
  lpb::ConvGeom g{N, C, H, W, KH, KW, SH, SW, PH, PW, DH, DW, 0, 0};
  g.OH = (H + 2 * PH - DH * (KH - 1) - 1) / SH + 1;
  g.OW = (W + 2 * PW - DW * (KW - 1) - 1) / SW + 1;
  LPB_REQUIRE(g.OH > 0 && g.OW > 0, "lpb_pack_conv2d_t: empty output");
  const int64_t K = reduce_mean ? N : (int64_t)N * g.OH * g.OW;
  LPB_REQUIRE(k0 >= 0 && k0 + K <= ldk, "lpb_pack_conv2d_t: rows exceed ldk");
  return lpb::pack_conv2d_t(x, g, scale, flags, reduce_mean, dst_hi, dst_lo, out_kind, ldk, k0, ST(stream));
}

int lpb_pack_nchw_t(const float* g, int64_t Nn, int Cc, int HW, float scale, int flags, int reduce_sum, void* dst_hi,
                    void* dst_lo, int out_kind, int64_t ldk, int64_t k0, void* stream) {
  LPB_REQUIRE(Nn >= 0 && Cc > 0 && HW > 0, "lpb_pack_nchw_t: bad extents");
  const int64_t K = reduce_sum ? Nn : Nn * HW;
  LPB_REQUIRE(k0 >= 0 && k0 + K <= ldk, "lpb_pack_nchw_t: rows exceed ldk");
  return lpb::pack_nchw_t(g, Nn, Cc, HW, scale, flags, reduce_sum, dst_hi, dst_lo, out_kind, ldk, k0, ST(stream));
}

static int make_geom(lpb::ConvGeom& g, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW, int DH,
                     int DW) {
  LPB_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && KH > 0 && KW > 0 && SH > 0 && SW > 0 && DH > 0 && DW > 0,
              "bad convolution geometry");
  g = lpb::ConvGeom{N, C, H, W, KH, KW, SH, SW, PH, PW, DH, DW, 0, 0};
  g.OH = (H + 2 * PH - DH * (KH - 1) - 1) / SH + 1;
  g.OW = (W + 2 * PW - DW * (KW - 1) - 1) / SW + 1;
  LPB_REQUIRE(g.OH > 0 && g.OW > 0, "empty convolution output");
  return 0;
}

int lpb_pack_conv2d_rows(const float* x, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW, int DH,
                         int DW, void* dst_hi, void* dst_lo, int out_kind, int64_t ld, void* stream) {
  lpb::ConvGeom g;
  if (make_geom(g, N, C, H, W, KH, KW, SH, SW, PH, PW, DH, DW)) return 1;
  LPB_REQUIRE(ld >= (int64_t)C * KH * KW, "lpb_pack_conv2d_rows: ld too small");
  return lpb::pack_conv2d_rows(x, g, dst_hi, dst_lo, out_kind, ld, ST(stream));
}

int lpb_pack_nchw_rows(const float* g, int64_t Q, int Cc, int HW, void* dst_hi, void* dst_lo, int out_kind, int64_t ld,
                       void* stream) {
  LPB_REQUIRE(Q >= 0 && Cc > 0 && HW > 0 && ld >= Cc, "lpb_pack_nchw_rows: bad extents");
  return lpb::pack_nchw_rows(g, Q, Cc, HW, dst_hi, dst_lo, out_kind, ld, ST(stream));
}

int lpb_pack_cast(const float* src, int64_t rows, int64_t cols, int64_t ld_src, void* dst_hi, void* dst_lo, int out_kind,
                  int64_t ld, void* stream) {
  LPB_REQUIRE(rows >= 0 && cols >= 0 && ld_src >= cols && ld >= cols, "lpb_pack_cast: bad extents");
  return lpb::pack_cast(src, rows, cols, ld_src, dst_hi, dst_lo, out_kind, ld, ST(stream));
}

int lpb_col2im(const float* Dc, int64_t ldd, int Q, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW,
               int DH, int DW, float* grad_in, void* stream) {
  lpb::ConvGeom g;
  if (make_geom(g, Q, C, H, W, KH, KW, SH, SW, PH, PW, DH, DW)) return 1;
  LPB_REQUIRE(ldd >= (int64_t)Q * g.OH * g.OW, "lpb_col2im: ldd too small");
  return lpb::col2im(Dc, ldd, g, grad_in, ST(stream));
}

int lpb_syrk_conv_patches_tc(const void* X_hi, const void* X_lo, int64_t ldx, int64_t Q, int H, int W, int Ci, int KH, int KW,
                             int PH, int PW, float alpha, int accumulate, float* D, int64_t ldd, int fp16_operands,
                             void* stream) {
  return lpb::syrk_conv_patches(X_hi, X_lo, ldx, Q, H, W, Ci, KH, KW, PH, PW, alpha, accumulate, D, ldd, fp16_operands,
                                ST(stream));
}

int lpb_conv_live_taps(int KH, int KW, int PH, int PW, int H, int W) { return lpb::syrk_conv_live_taps(KH, KW, PH, PW, H, W); }

int lpb_taps_to_param_accumulate(const float* T, int64_t ldt, int Ci, int Ci_pad, int KH, int KW, int PH, int PW, int H, int W,
                                 float* out, int64_t ldo, void* stream) {
  LPB_REQUIRE(KH > 0 && KW > 0 && KH * KW <= 9, "lpb_taps_to_param_accumulate: kernel window larger than 9 taps");
  LPB_REQUIRE(ldo >= (int64_t)Ci * KH * KW, "lpb_taps_to_param_accumulate: leading dimension too small");
  return lpb::taps_to_param_accumulate(T, ldt, Ci, Ci_pad, KH, KW, PH, PW, H, W, out, ldo, ST(stream));
}

int lpb_diag_conv_sq_tc(const void* G_hi, const void* G_lo, int64_t ldg, const void* X_hi, const void* X_lo, int64_t ldx,
                        int64_t Qtot, int64_t Nimg, int H, int W, int Ci, int Co, int KH, int KW, int PH, int PW, float alpha,
                        int accumulate, float* D, int64_t ldd, void* stream) {
  return lpb::diag_conv_sq(G_hi, G_lo, ldg, X_hi, X_lo, ldx, Qtot, Nimg, H, W, Ci, Co, KH, KW, PH, PW, alpha, accumulate, D, ldd,
                           ST(stream));
}

int lpb_taps_to_param_rect(const float* Dt, int64_t ldt, int Co, int Ci, int Ci_pad, int KK, float* out, int64_t ldo,
                           void* stream) {
  LPB_REQUIRE(Ci_pad >= Ci && ldt >= (int64_t)Ci_pad * KK && ldo >= (int64_t)Ci * KK, "lpb_taps_to_param_rect: bad extents");
  return lpb::taps_to_param_rect(Dt, ldt, Co, Ci, Ci_pad, KK, out, ldo, ST(stream));
}

int lpb_pack_cast_fused(const float* src, int64_t rows, int64_t cols, int64_t ld_src, const float* scale, const float* y,
                        int64_t rows_y, int64_t ld_y, void* dst_hi, void* dst_lo, int out_kind, int64_t ld, void* stream) {
  LPB_REQUIRE(ld_src >= cols && ld >= cols, "lpb_pack_cast_fused: leading dimension too small");
  return lpb::pack_cast_fused(src, rows, cols, ld_src, scale, y, rows_y, ld_y, dst_hi, dst_lo, out_kind, ld, ST(stream));
}

int lpb_col2im_nhwc(const float* Dc, int64_t ldd, int Q, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW,
                    int DH, int DW, float* grad_in, void* stream) {
  lpb::ConvGeom g;
  if (make_geom(g, Q, C, H, W, KH, KW, SH, SW, PH, PW, DH, DW)) return 1;
  LPB_REQUIRE(ldd >= (int64_t)KH * KW * C, "lpb_col2im_nhwc: ldd too small");
  return lpb::col2im_nhwc(Dc, ldd, g, grad_in, ST(stream));
}

int lpb_maxpool2d_bwd_nhwc(const float* g, const int64_t* idx, float* out, int64_t Q, int Nb, int C, int H, int W, int OH,
                           int OW, int k, int s, int p, void* stream) {
  return lpb::maxpool2d_bwd_nhwc(g, idx, out, Q, Nb, C, H, W, OH, OW, k, s, p, ST(stream));
}

int lpb_gemm_nt_f32(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                    float alpha, int accumulate, float* D, int64_t ldd, int symmetric, void* stream) {
  LPB_REQUIRE(lda >= K && ldb >= K && ldd >= N, "lpb_gemm_nt_f32: leading dimension too small");
  return lpb::gemm_nt_f32(A, lda, B, ldb, M, N, K, alpha, accumulate, D, ldd, symmetric, ST(stream));
}

int lpb_gemm_nt_tc(const void* A_hi, const void* A_lo, int64_t lda, const void* B_hi, const void* B_lo, int64_t ldb,
                   int64_t M, int64_t N, int64_t K, float alpha, int accumulate, float* D, int64_t ldd, int symmetric,
                   int fp16_operands, void* stream) {
  LPB_REQUIRE(lda >= K && ldb >= K && ldd >= N, "lpb_gemm_nt_tc: leading dimension too small");
  LPB_REQUIRE((A_lo == nullptr) == (B_lo == nullptr), "lpb_gemm_nt_tc: A_lo and B_lo must both be given or both NULL");
  return lpb::gemm_nt_bf16(A_hi, A_lo, lda, B_hi, B_lo, ldb, M, N, K, alpha, accumulate, D, ldd, symmetric, fp16_operands,
                           ST(stream));
}

int lpb_scale_channels(const float* g, const float* scale, float* out, int64_t n, int C, int64_t inner, void* stream) {
  return lpb::scale_channels(g, scale, out, n, C, inner, ST(stream));
}

int lpb_relu_bwd(const float* g, const float* y, float* out, int64_t n, int reps, void* stream) {
  return lpb::relu_bwd(g, y, out, n, reps, ST(stream));
}

int lpb_maxpool2d_bwd(const float* g, const int64_t* idx, float* out, int64_t Q, int Nb, int C, int H, int W, int OH, int OW,
                      int k, int s, int p, void* stream) {
  return lpb::maxpool2d_bwd(g, idx, out, Q, Nb, C, H, W, OH, OW, k, s, p, ST(stream));
}

int lpb_gemm_tn_tc(const void* A_hi, const void* A_lo, int64_t lda, const void* B_hi, const void* B_lo, int64_t ldb,
                   int64_t M, int64_t N, int64_t K, float alpha, int accumulate, float* D, int64_t ldd, int symmetric,
                   int fp16_operands, void* stream) {
  LPB_REQUIRE(lda >= M && ldb >= N && ldd >= N, "lpb_gemm_tn_tc: leading dimension too small");
  LPB_REQUIRE((A_lo == nullptr) == (B_lo == nullptr), "lpb_gemm_tn_tc: A_lo and B_lo must both be given or both NULL");
  return lpb::gemm_tn_rows(A_hi, A_lo, lda, B_hi, B_lo, ldb, M, N, K, alpha, accumulate, D, ldd, symmetric, fp16_operands,
                           ST(stream));
}

int lpb_conv_nhwc_tc(const void* X_hi, const void* X_lo, int64_t Q, int H, int W, int64_t Kc, int64_t ldx, const void* W_hi,
                     const void* W_lo, int64_t ldw, int N, int KH, int KW, int base_h, int base_w, int sgn, float alpha,
                     float* D, int64_t ldd, int fp16_operands, void* stream) {
  LPB_REQUIRE(sgn == 1 || sgn == -1, "lpb_conv_nhwc_tc: sgn must be +1 or -1");
  return lpb::conv_nhwc_bf16(X_hi, X_lo, Q, H, W, Kc, ldx, W_hi, W_lo, ldw, N, KH, KW, base_h, base_w, sgn, alpha, D, ldd,
                             fp16_operands, ST(stream));
}

int lpb_conv_bwd_strided_tc(const void* G_hi, const void* G_lo, int64_t Q, int OH, int OW, int64_t Co, int64_t ldg,
                            const void* W_hi, const void* W_lo, int64_t ldw, int Ci, int KH, int KW, int SH, int SW, int PH,
                            int PW, int H, int W, float* D, int64_t ldd, void* stream) {
  return lpb::conv_bwd_strided(G_hi, G_lo, Q, OH, OW, Co, ldg, W_hi, W_lo, ldw, Ci, KH, KW, SH, SW, PH, PW, H, W, D, ldd,
                               ST(stream));
}

int lpb_shared_weight_contract(int mode, const float* G, int64_t ldg, const float* A, int64_t lda, int d_out, int d_in,
                               int T, int Nn, int ncols, float scale, float* out, int64_t out_ld, int64_t js_stride_n,
                               int64_t js_stride_c, void* stream) {
  LPB_REQUIRE(mode == 0 || mode == 1, "lpb_shared_weight_contract: mode must be 0 or 1");
  LPB_REQUIRE(T > 0 && ldg >= (int64_t)Nn * ncols * T && lda >= (int64_t)Nn * T, "lpb_shared_weight_contract: bad extents");
  return lpb::shared_weight_contract(mode, G, ldg, A, lda, d_out, d_in, T, Nn, ncols, scale, out, out_ld, js_stride_n,
                                     js_stride_c, ST(stream));
}

int lpb_maxpool2d_bwd_pack_nhwc(const float* g, const int64_t* idx, const float* scale, const float* y, void* dst_hi, void* dst_lo,
                                int64_t ld, int64_t Q, int Nb, int C, int H, int W, int OH, int OW, int k, int s, int p, void* stream) {
  LPB_REQUIRE(g != nullptr && idx != nullptr && dst_hi != nullptr && dst_lo != nullptr, "lpb_maxpool2d_bwd_pack_nhwc: null operand");
  return lpb::maxpool2d_bwd_pack_nhwc(g, idx, scale, y, dst_hi, dst_lo, ld, Q, Nb, C, H, W, OH, OW, k, s, p, ST(stream));
}

int lpb_kron_conv_quadform(const float* Gt, int64_t ldg, int64_t g_stride_c, const float* At, int64_t lda, int d_out, int d_in,
                           int T, int Nn, int C, const float* l1, const float* l2, float delta, int damping, float* out,
                           void* stream) {
  LPB_REQUIRE(T > 0 && lda >= (int64_t)Nn * T && g_stride_c >= (int64_t)Nn * T && ldg >= (int64_t)(C - 1) * g_stride_c + (int64_t)Nn * T,
              "lpb_kron_conv_quadform: bad extents");
  return lpb::kron_conv_quadform(Gt, ldg, g_stride_c, At, lda, d_out, d_in, T, Nn, C, l1, l2, delta, damping, out, ST(stream));
}

// ---- layer-level KFAC entry points (lpb_kfac_accum_A / _B, lpb_workspace_bytes) -------------------------
// Compositions of the kernels above for a caller that holds plain fp32 activations / gradients and wants one call per
// factor: operand split into 16-bit hi/lo rows in the caller's workspace, then the MN-major tcgen05 SYRK.
static int64_t ld_round8(int64_t k) { return (k + 7) / 8 * 8 < 8 ? 8 : (k + 7) / 8 * 8; }

int64_t lpb_workspace_bytes(int64_t rows, int64_t d) {
  if (rows < 0 || d <= 0) return -1;
  return 2 * rows * ld_round8(d) * 2 + 512;   // hi + lo 16-bit rows, 256-byte aligned halves
}

static int split_workspace(void* ws, int64_t ws_bytes, int64_t rows, int64_t d, void** hi, void** lo, int64_t* ld) {
  LPB_REQUIRE(ws != nullptr && ws_bytes >= lpb_workspace_bytes(rows, d), "workspace too small: need %lld bytes",
              (long long)lpb_workspace_bytes(rows, d));
  *ld = ld_round8(d);
  uintptr_t base = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
  *hi = (void*)base;
  *lo = (void*)((base + (uintptr_t)(rows * *ld * 2) + 255) & ~(uintptr_t)255);
  return 0;
}

int lpb_kfac_accum_rows(const float* X, int64_t rows, int64_t d, int64_t ldx, float alpha, int fp16_operands, void* workspace,
                        int64_t workspace_bytes, float* out, int64_t ldo, void* stream) {
  LPB_REQUIRE(X != nullptr && out != nullptr && rows > 0 && d > 0 && ldx >= d && ldo >= d, "lpb_kfac_accum_rows: bad extents");
  void *hi, *lo;
  int64_t ld;
  if (split_workspace(workspace, workspace_bytes, rows, d, &hi, &lo, &ld)) return 1;
  if (lpb::pack_cast(X, rows, d, ldx, hi, lo, fp16_operands ? lpb::OUT_F16_HILO : lpb::OUT_BF16_HILO, ld, ST(stream))) return 1;
  return lpb::gemm_tn_rows(hi, lo, ld, hi, lo, ld, d, d, rows, alpha, 1, out, ldo, 1, fp16_operands ? 1 : 0, ST(stream));
}

int lpb_kfac_accum_conv_input(const float* x, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW, int DH,
                          int DW, float alpha, void* workspace, int64_t workspace_bytes, float* out, int64_t ldo, void* stream) {
  lpb::ConvGeom g;
  if (make_geom(g, N, C, H, W, KH, KW, SH, SW, PH, PW, DH, DW)) return 1;
  const int64_t rows = (int64_t)N * g.OH * g.OW, d = (int64_t)C * KH * KW;
  LPB_REQUIRE(out != nullptr && ldo >= d, "lpb_kfac_accum_conv_input: bad output extents");
  void *hi, *lo;
  int64_t ld;
  if (split_workspace(workspace, workspace_bytes, rows, d, &hi, &lo, &ld)) return 1;
  if (lpb::pack_conv2d_rows(x, g, hi, lo, lpb::OUT_F16_HILO, ld, ST(stream))) return 1;
  return lpb::gemm_tn_rows(hi, lo, ld, hi, lo, ld, d, d, rows, alpha, 1, out, ldo, 1, 1, ST(stream));
}

int lpb_jac_linear_write(const float* g, const float* a, int Nn, int C, int d_out, int d_in, float* Js,
                         int64_t js_stride_n, int64_t js_stride_c, int64_t off_w, int64_t off_b, void* stream) {
  return lpb::jac_linear_write(g, a, Nn, C, d_out, d_in, Js, js_stride_n, js_stride_c, off_w, off_b, ST(stream));
}

int lpb_ll_jacobian_write(const float* phi, int Nn, int C, int D, int has_bias, float* Js, void* stream) {
  return lpb::ll_jacobian_write(phi, Nn, C, D, has_bias, Js, ST(stream));
}

int lpb_batched_pair_dot(const float* X, const float* Z, const float* m, int64_t m_stride, int Nn, int CX, int CZ, int d,
                         int64_t x_stride_n, int64_t x_stride_c, int64_t z_stride_n, int64_t z_stride_c, int accumulate,
                         float* out, void* stream) {
  return lpb::batched_pair_dot(X, Z, m, m_stride, Nn, CX, CZ, d, x_stride_n, x_stride_c, z_stride_n, z_stride_c,
                               accumulate, out, ST(stream));
}

int lpb_ll_ggn_expand(const float* G, int C, int D, int has_bias, int accumulate, float* H, void* stream) {
  return lpb::ll_ggn_expand(G, C, D, has_bias, accumulate, H, ST(stream));
}

int lpb_ll_sigma_gather(const float* Sigma, int C, int D, int has_bias, float* Sg, void* stream) {
  return lpb::ll_sigma_gather(Sigma, C, D, has_bias, Sg, ST(stream));
}

int lpb_eigh_jacobi(const float* A, int batch, int n, float* evals, float* Q, int max_sweeps, void* stream) {
  return lpb::eigh_jacobi(A, batch, n, evals, Q, max_sweeps, ST(stream));
}

}  // extern "C"
