/* laplace_b200 -- C ABI of the B200-native curvature hot path for the Laplace library.
 *
 * Drop-in boundary: the reference's curvature backends are Python classes deriving from
 * laplace.curvature.CurvatureInterface (reference laplace/curvature/curvature.py:12-291).  The
 * per-batch arithmetic those classes perform is what this library replaces; the Python subclass
 * `laplace_b200.B200GGN / B200EF` binds these entry points through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current device unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - all calls are asynchronous on `stream`; nothing is allocated or freed by the library
 *     except where a `workspace` argument is documented; buffers are owned by the caller;
 *   - return value 0 = success, non-zero = failure with a message in lpb_last_error()
 *     (thread-local).  The Python wrapper raises RuntimeError (reference error convention:
 *     Python exceptions only);
 *   - "K-major operand": matrix X[rows, ld] with the contraction index k contiguous
 *     (element (j, k) at X[j*ld + k]).  The pack entry points produce this layout.
 *   - bf16 buffers are passed as void* (uint16 storage).  out_kind: 0 = fp32, 1 = bf16,
 *     2 = bf16 hi + bf16 lo (error-compensated split: x ~= hi + lo, |x - hi - lo| <= 2^-17 |x|).
 */
#ifndef LAPLACE_B200_H_
#define LAPLACE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPB_OUT_F32 0
#define LPB_OUT_BF16 1
#define LPB_OUT_BF16_HILO 2
#define LPB_OUT_F16_HILO 3 /* fp16 hi + fp16 lo: 22 significant bits, operands must stay inside fp16 range */
#define LPB_PACK_SQUARE 1

/* ---- library ---------------------------------------------------------------------------- */
int lpb_version(void);
const char* lpb_last_error(void);
/* sm count / compute capability of the current device */
int lpb_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Schedule of lpb_gemm_nt_tc / lpb_gemm_tn_tc: -1 automatic (default), 0 one 128 x 128 tile per CTA,
 * 1 256 x 256 CTA-pair tiles (tcgen05 cta_group::2) whenever M, N >= 256, 2 persistent CTAs with double-buffered
 * TMEM accumulators.  Results agree to fp32 rounding. */
int lpb_set_gemm_tile_mode(int mode);
/* lpb_pack_cast_fused / lpb_relu_bwd read their ReLU mask y once per mask element and walk the gradient rows that share
 * it ("mask-major") when the mask has at least this many elements (default 4 Mi: larger masks do not survive in L2 between
 * the folded curvature columns); smaller masks keep the row-major kernels.  < 0: always row-major.  Bit-identical results. */
int lpb_set_mask_major_min(int64_t min_mask_elems);

/* ---- pack: layer inputs / output gradients -> K-major staging ----------------------------
 * Front end of the KFAC factor contractions that curvlinops performs as einsum("b i,b j->i j")
 * behind reference laplace/curvature/curvlinops.py:100 (linop._compute_kfac()).              */

/* src [rows, cols] fp32 row-major (ld_src)  ->  dst[(z*cols + j), k0 + k] = f(scale*src[k, j]) * row_scale[z*rows + k]
 * (f = square if flags & LPB_PACK_SQUARE; row_scale may be NULL when nrep == 1).            */
int lpb_pack_rows_t(const float* src, int64_t rows, int64_t cols, int64_t ld_src, const float* row_scale, int nrep,
                    float scale, int flags, void* dst_hi, void* dst_lo, int out_kind, int64_t ldk, int64_t k0,
                    void* stream);

/* x [N, C, H, W] fp32 -> unfolded patches dst[(ci,kh,kw), k0 + (n,oh,ow)] (the rows KFAC-expand
 * uses for nn.Conv2d); reduce_mean != 0: KFAC-reduce rows dst[(ci,kh,kw), k0 + n] = mean over (oh,ow). */
int lpb_pack_conv2d_t(const float* x, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW, int DH,
                      int DW, float scale, int flags, int reduce_mean, void* dst_hi, void* dst_lo, int out_kind,
                      int64_t ldk, int64_t k0, void* stream);

/* g [Nn, Cc, HW] fp32 (NCHW gradient of a conv output) -> dst[ch, k0 + n*HW + hw];
 * reduce_sum != 0: dst[ch, k0 + n] = sum_hw (KFAC-reduce).                                    */
int lpb_pack_nchw_t(const float* g, int64_t Nn, int Cc, int HW, float scale, int flags, int reduce_sum, void* dst_hi,
                    void* dst_lo, int out_kind, int64_t ldk, int64_t k0, void* stream);

/* same contraction on ROW-major operands (sample rows x features, the layout activations and gradients already
 * have): D[M, N] (+)= alpha * A[K, M]^T * B[K, N]; tcgen05 with MN-major shared-memory descriptors, so X^T X
 * needs no transposing pack.  lda >= M, ldb >= N (multiples of 8 elements).                                    */
int lpb_gemm_tn_tc(const void* A_hi, const void* A_lo, int64_t lda, const void* B_hi, const void* B_lo, int64_t ldb,
                   int64_t M, int64_t N, int64_t K, float alpha, int accumulate, float* D, int64_t ldd, int symmetric,
                   int fp16_operands, void* stream);

/* ---- convolution engine operands (forward / backward-data of nn.Conv2d as GEMMs) ----------
 * Replaces the model-side torch.func / autograd convolution passes the reference runs below
 * CurvatureInterface.jacobians (curvature/curvature.py:111-117) when fp32-accurate Jacobians are required. */
/* patch-major im2col: x [N,C,H,W] -> dst[(n,oh,ow), (ci,kh,kw)], row stride ld                      */
int lpb_pack_conv2d_rows(const float* x, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW, int DH,
                         int DW, void* dst_hi, void* dst_lo, int out_kind, int64_t ld, void* stream);
/* g [Q, Cc, HW] -> dst[(q,hw), ch], row stride ld (channel index contiguous)                            */
int lpb_pack_nchw_rows(const float* g, int64_t Q, int Cc, int HW, void* dst_hi, void* dst_lo, int out_kind, int64_t ld,
                       void* stream);
/* src [rows, cols] fp32 (ld_src) -> same layout in out_kind (row stride ld)                               */
int lpb_pack_cast(const float* src, int64_t rows, int64_t cols, int64_t ld_src, void* dst_hi, void* dst_lo, int out_kind,
                  int64_t ld, void* stream);
/* pack_cast with the element-wise maps that precede a convolution's reverse pass fused in:
 * dst[r, c] = split(src[r, c] * scale[c] * (y[r % rows_y, c] > 0)); scale / y may be NULL.  Replaces the vmapped
 * threshold_backward + native_batch_norm_backward (eval) of the reference's reverse pass plus the operand cast.   */
int lpb_pack_cast_fused(const float* src, int64_t rows, int64_t cols, int64_t ld_src, const float* scale, const float* y,
                        int64_t rows_y, int64_t ld_y, void* dst_hi, void* dst_lo, int out_kind, int64_t ld, void* stream);
/* col2im gather: Dc [(ci,kh,kw), ldd] with columns (q,oh,ow) -> grad_in [Q, C, H, W] (overwrites)        */
int lpb_col2im(const float* Dc, int64_t ldd, int Q, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW,
               int DH, int DW, float* grad_in, void* stream);
/* channels-last form: Dc [(q,oh,ow), ldd] with columns (kh,kw,ci) -> grad_in [Q, H, W, C] (overwrites)     */
int lpb_col2im_nhwc(const float* Dc, int64_t ldd, int Q, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW,
                    int DH, int DW, float* grad_in, void* stream);

/* ---- contractions ------------------------------------------------------------------------
 * D[M, N] (fp32, ldd)  (+)=  alpha * A[M, K] * B[N, K]^T   on K-major operands.
 * symmetric != 0 (requires A == B, M == N): SYRK -- only tiles on/above the diagonal are
 * computed and mirrored.  accumulate == 0 overwrites D.
 * Replaces: KFAC factor einsums (curvlinops.py:100), full GGN einsum "bcp,bck,bkq->pq"
 * (curvature.py:406/408), EF einsum "bp,bq->pq" (curvature.py:492), the eigenbasis rotations of
 * KronDecomposed._bmm (utils/matrix.py:447-450) and J Sigma J^T (baselaplace.py:1683-1684).  */

/* exact fp32 SIMT path */
int lpb_gemm_nt_f32(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                    float alpha, int accumulate, float* D, int64_t ldd, int symmetric, void* stream);

/* tcgen05 tensor-core path: 16-bit operands (TMA -> smem -> tcgen05.mma, fp32 accumulation in TMEM);
 * fp16_operands = 0: bf16, 1: fp16 (both operands).  A_lo/B_lo non-NULL: error-compensated 3-product mode
 * hi*hi + hi*lo + lo*hi.  Requirements: lda, ldb multiples of 8 elements, operand pointers 16-byte aligned. */
int lpb_gemm_nt_tc(const void* A_hi, const void* A_lo, int64_t lda, const void* B_hi, const void* B_lo, int64_t ldb,
                   int64_t M, int64_t N, int64_t K, float alpha, int accumulate, float* D, int64_t ldd, int symmetric,
                   int fp16_operands, void* stream);

/* implicit-GEMM convolution, stride 1, same padding, NHWC bf16 / fp16 (hi/lo) operands, fp32 NHWC output:
 *   D[(q,h,w), n] = alpha * sum_{kh,kw,k} X[q, h + base_h + sgn*kh, w + base_w + sgn*kw, k] * Wt[(kh*KW+kw)*N + n, k]
 * X [Q,H,W,ldx] (ldx >= Kc), Wt [KH*KW*N, ldw]; H*W must divide 128; out-of-range taps read zeros (TMA fill).
 * forward of nn.Conv2d: (base, sgn) = (-pad, +1); backward-data: (+pad, -1) with Wt[(tap), ci, co].          */
int lpb_conv_nhwc_tc(const void* X_hi, const void* X_lo, int64_t Q, int H, int W, int64_t Kc, int64_t ldx, const void* W_hi,
                     const void* W_lo, int64_t ldw, int N, int KH, int KW, int base_h, int base_w, int sgn, float alpha,
                     float* D, int64_t ldd, int fp16_operands, void* stream);

/* backward-data of a STRIDED convolution as implicit GEMMs (no [rows, KH*KW*Ci] intermediate, no col2im): the input
 * gradient splits into SH*SW parity classes, each a stride-1 convolution of the output-gradient grid with the taps
 * kh = h + PH (mod SH), kw = w + PW (mod SW), stored to the class's pixels of D [Q, H, W, ldd] (all of D is written).
 * G [(q,oh,ow), ldg] 16-bit hi(/lo) rows, Wt [(kh,kw,ci), ldw] tap-major; needs H == OH*SH, W == OW*SW and an
 * OH x OW grid that tiles 128-row blocks.                                                                           */
int lpb_conv_bwd_strided_tc(const void* G_hi, const void* G_lo, int64_t Q, int OH, int OW, int64_t Co, int64_t ldg,
                            const void* W_hi, const void* W_lo, int64_t ldw, int Ci, int KH, int KW, int SH, int SW, int PH,
                            int PW, int H, int W, float* D, int64_t ldd, void* stream);

/* ---- KFAC input factor of a stride-1 'same' convolution without im2col -----------------------------------------
 * D[(t,ci),(t',cj)] (+)= alpha * sum_{n,h,w} x[n,h+kh-PH,w+kw-PW,ci] * x[n,h+kh'-PH,w+kw'-PW,cj]   (zero padding),
 * t = kh*KW + kw; the i-th LIVE tap's feature (t,ci) sits at index i*Ci_pad + ci with Ci_pad = Ci rounded up to 64
 * (D is [live*Ci_pad]^2, live = lpb_conv_live_taps(...), padded rows / columns exactly zero); x given as 16-bit hi(/lo) NHWC rows [(n,h,w), Ci] (ldx).
 * Replaces unfold + einsum("b t i, b t j -> i j") behind reference laplace/curvature/curvlinops.py:100 for those
 * layers; the 9x larger patch matrix is never formed (shifted 4-D TMA boxes feed the tensor cores directly).
 * Needs H*W dividing 64 or (W | 64 and 64/W | H).  D is symmetric (both triangles written).                      */
int lpb_syrk_conv_patches_tc(const void* X_hi, const void* X_lo, int64_t ldx, int64_t Q, int H, int W, int Ci, int KH, int KW,
                             int PH, int PW, float alpha, int accumulate, float* D, int64_t ldd, int fp16_operands,
                             void* stream);
/* Diagonal GGN / EF of a stride-1 'same' convolution weight, per-sample weight gradients squared and summed on the
 * tensor cores (replaces einsum("bcp,bcp->p") / ("bp,bp->p") on materialised per-sample gradients,
 * reference laplace/curvature/curvature.py:429-431, 504):
 *   D[co, t*Ci_pad + ci] (+)= alpha * sum_{q < Qtot} ( sum_{h,w} G[(q,h,w), co] * x[q % Nimg, h+kh-PH, w+kw-PW, ci] )^2
 * G: 16-bit hi(/lo) output-gradient rows [(q,h,w), Co] of all folded curvature columns, x: NHWC rows [(n,h,w), Ci] in
 * the same format (bf16).  Needs H*W >= 64 with W | 64 and 64/W | H.  lpb_taps_to_param_rect folds the tap-major,
 * channel-padded columns into the parameter order (ci,kh,kw).                                                       */
int lpb_diag_conv_sq_tc(const void* G_hi, const void* G_lo, int64_t ldg, const void* X_hi, const void* X_lo, int64_t ldx,
                        int64_t Qtot, int64_t Nimg, int H, int W, int Ci, int Co, int KH, int KW, int PH, int PW, float alpha,
                        int accumulate, float* D, int64_t ldd, void* stream);
int lpb_taps_to_param_rect(const float* Dt, int64_t ldt, int Co, int Ci, int Ci_pad, int KK, float* out, int64_t ldo,
                           void* stream);
/* Kernel positions whose window overlaps an H x W image at all (|kh-PH| < H and |kw-PW| < W); the others only read
 * zero padding.  lpb_syrk_conv_patches_tc computes the LIVE taps only: D is [live*Ci_pad]^2, live taps in ascending
 * kernel position.                                                                                               */
int lpb_conv_live_taps(int KH, int KW, int PH, int PW, int H, int W);
/* out[(ci*KK+t), (cj*KK+t')] += T[(i*Ci_pad+ci), (i'*Ci_pad+cj)] for live taps t = tap(i), t' = tap(i'): tap-major
 * factor of lpb_syrk_conv_patches_tc -> parameter order (ci,kh,kw); KK = KH*KW <= 9                              */
int lpb_taps_to_param_accumulate(const float* T, int64_t ldt, int Ci, int Ci_pad, int KH, int KW, int PH, int PW, int H, int W,
                                 float* out, int64_t ldo, void* stream);

/* ---- reverse-pass element-wise maps of the convolution engine (columns folded into the batch) -----------------
 * out[i] = g[i] * scale[(i / inner) % C]                 frozen BatchNorm as per-channel affine map (backward)   */
int lpb_scale_channels(const float* g, const float* scale, float* out, int64_t n, int C, int64_t inner, void* stream);
/* out[r*n + i] = y[i] > 0 ? g[r*n + i] : 0               ReLU backward, forward output y shared by `reps` columns */
int lpb_relu_bwd(const float* g, const float* y, float* out, int64_t n, int reps, void* stream);
/* max-pool backward (gather form): g [Q,C,OH,OW], idx [Nb,C,OH,OW] argmax (h*W+w) of image q % Nb -> out [Q,C,H,W] */
int lpb_maxpool2d_bwd(const float* g, const int64_t* idx, float* out, int64_t Q, int Nb, int C, int H, int W, int OH, int OW,
                      int k, int s, int p, void* stream);
/* the same on channels-last memory: g [Q,OH,OW,C], idx [Nb,OH,OW,C] -> out [Q,H,W,C]                                  */
int lpb_maxpool2d_bwd_nhwc(const float* g, const int64_t* idx, float* out, int64_t Q, int Nb, int C, int H, int W, int OH,
                           int OW, int k, int s, int p, void* stream);

/* ---- weight-sharing layers: per-sample layer Jacobians ------------------------------------
 * P_q[i,j] = sum_t G[i, q*T+t] * A[j, (q % Nn)*T + t], q = c*Nn + n over ncols back-propagated columns.
 * mode 0: out[i*out_ld + j] += scale * sum_q P_q[i,j]^2      (diag GGN / EF, curvature.py:429-431, :504)
 * mode 1: Js[n*js_stride_n + c*js_stride_c + i*d_in + j] = P_q[i,j]   (Jacobian rows, curvature.py:115-124) */
int lpb_shared_weight_contract(int mode, const float* G, int64_t ldg, const float* A, int64_t lda, int d_out, int d_in,
                               int T, int Nn, int ncols, float scale, float* out, int64_t out_ld, int64_t js_stride_n,
                               int64_t js_stride_c, void* stream);

/* max-pool reverse map FUSED with the operand split of the convolution chain in front of the pool (conv -> frozen BN -> ReLU
 * -> max-pool, the ResNet stem): rows [(col, n, h, w), C] bf16 hi/lo = split( unpool(g) * scale[c] * (y > 0) ); g, idx, y
 * channels-last (y = the pool's input, NULL: no mask; scale NULL: 1).  Replaces lpb_maxpool2d_bwd_nhwc + lpb_pack_cast_fused
 * for that chain: the fp32 un-pooled gradient is never written.                                                      */
int lpb_maxpool2d_bwd_pack_nhwc(const float* g, const int64_t* idx, const float* scale, const float* y, void* dst_hi, void* dst_lo,
                                int64_t ld, int64_t Q, int Nb, int C, int H, int W, int OH, int OW, int k, int s, int p, void* stream);

/* ---- layer-level KFAC entry points (compositions of the kernels above) ------------------------------------
 * For a caller that holds plain fp32 tensors and wants ONE call per Kronecker factor of CurvlinopsInterface.kron
 * (curvature/curvlinops.py:77-108).  `workspace`: device scratch of at least lpb_workspace_bytes(rows, d) bytes, owned by
 * the caller.  `out` [d, d] fp32 is ACCUMULATED into (the caller zeroes it once per fit).
 *   lpb_kfac_accum_rows   out += alpha * X^T X for fp32 rows X [rows, d]: the input factor A of an nn.Linear (rows = the
 *                         (n, t) layer inputs, alpha = sqrt(factor)/(N*T), fp16_operands = 1) or the output-gradient factor
 *                         B of any layer (rows = the (col, n, t) gradient rows, alpha = sqrt(factor), fp16_operands = 0: bf16
 *                         hi/lo keeps the gradients' dynamic range).  Three tensor-core products, fp32 accumulation.
 *   lpb_kfac_accum_conv_input out += alpha * P^T P for the unfolded patches P [(n,oh,ow), C*KH*KW] of an NCHW fp32 input
 *                         (curvlinops' unfold + einsum): workspace rows = N*OH*OW, d = C*KH*KW.                        */
int64_t lpb_workspace_bytes(int64_t rows, int64_t d);
int lpb_kfac_accum_rows(const float* X, int64_t rows, int64_t d, int64_t ldx, float alpha, int fp16_operands, void* workspace,
                        int64_t workspace_bytes, float* out, int64_t ldo, void* stream);
int lpb_kfac_accum_conv_input(const float* x, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW, int DH,
                          int DW, float alpha, void* workspace, int64_t workspace_bytes, float* out, int64_t ldo, void* stream);

/* Kron GLM-predictive quadratic form of a weight-sharing layer without the dense Jacobian (replaces the per-(n,c) dense
 * rotations of KronDecomposed._bmm / inv_square_form, utils/matrix.py:406-461, for convolution / token-shared layers):
 *   out[n,c,k] += sum_ij w(i,j) Z_c[i,j] Z_k[i,j],  Z_c[i,j] = sum_t Gt[i, c*g_stride_c + n*T + t] * At[j, n*T + t],
 *   w = 1/(l1[i]*l2[j] + delta)  (damping != 0: 1/((l1[i]+sqrt(delta))*(l2[j]+sqrt(delta)))).
 * Gt [d_out, ldg], At [d_in, lda]: eigenbasis-rotated output-gradient / unfolded-input rows, K-major fp32; out [Nn, C, C]
 * (accumulated into; C <= 12).                                                                                      */
int lpb_kron_conv_quadform(const float* Gt, int64_t ldg, int64_t g_stride_c, const float* At, int64_t lda, int d_out, int d_in,
                           int T, int Nn, int C, const float* l1, const float* l2, float delta, int damping, float* out,
                           void* stream);

/* ---- Jacobian writers (reference CurvatureInterface.jacobians / last_layer_jacobians) ----- */
/* no weight sharing: Js[n,c, off_w + i*d_in + j] = g[c,n,i]*a[n,j]; Js[n,c, off_b + i] = g[c,n,i]
 * (off_w / off_b < 0: skip that block).  g [C, Nn, d_out], a [Nn, d_in].                      */
int lpb_jac_linear_write(const float* g, const float* a, int Nn, int C, int d_out, int d_in, float* Js,
                         int64_t js_stride_n, int64_t js_stride_c, int64_t off_w, int64_t off_b, void* stream);
/* J_n = [I_C (x) phi_n^T, I_C] (curvature.py:157-165); Js [Nn, C, C*D (+C)] contiguous.        */
int lpb_ll_jacobian_write(const float* phi, int Nn, int C, int D, int has_bias, float* Js, void* stream);

/* ---- batched pair reductions (epilogues of the predictive quadratic forms) -----------------
 * out[n, c, k] (+)= sum_i X[n*x_stride_n + c*x_stride_c + i] * Z[n*z_stride_n + k*z_stride_c + i] * (m ? m[n*m_stride + i] : 1)
 * Replaces torch.bmm(W, SW^T) in KronDecomposed.inv_square_form (utils/matrix.py:458-461).     */
int lpb_batched_pair_dot(const float* X, const float* Z, const float* m, int64_t m_stride, int Nn, int CX, int CZ, int d,
                         int64_t x_stride_n, int64_t x_stride_c, int64_t z_stride_n, int64_t z_stride_c, int accumulate,
                         float* out, void* stream);

/* ---- last-layer full GGN (structured form of curvature.py:398-408 with last_layer=True) ---- */
/* G [Dt, C(C+1)/2, Dt] (Dt = D + has_bias) -> H [P, P], P = C*D (+C), ordering [vec(W) row-major; b] */
int lpb_ll_ggn_expand(const float* G, int C, int D, int has_bias, int accumulate, float* H, void* stream);
/* Sigma [P, P] -> Sg [C*C, Dt, Dt] with Sg[(c,k), et, dt] = Sigma[idx(c,dt), idx(k,et)]         */
int lpb_ll_sigma_gather(const float* Sigma, int C, int D, int has_bias, float* Sg, void* stream);

/* ---- symmetric eigendecomposition (reference Kron.decompose -> symeig, utils/utils.py:193-228)
 * Batched cyclic one-sided Jacobi for n <= LPB_EIGH_MAX_N; A [batch, n, n] fp32 symmetric (upper
 * triangle read) -> eigenvalues ascending, clamped at 0, NaN -> 0; eigenvectors in columns of Q.
 * Returns non-zero (with message) when n exceeds the limit.                                     */
#define LPB_EIGH_MAX_N 128
int lpb_eigh_jacobi(const float* A, int batch, int n, float* evals, float* Q, int max_sweeps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAPLACE_B200_H_ */
