"""Tensor-level wrappers over the C ABI (``include/laplace_b200.h``).

PyTorch is used here only for device memory and streams: every function below turns
``torch.Tensor`` arguments into raw device pointers and launches one native kernel on the
current CUDA stream.  There is no torch/CPU fallback; a missing library or a non-CUDA tensor
raises.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib

F32, BF16, BF16X3, F16X3 = 0, 1, 2, 3  # out_kind of the pack kernels (LPB_OUT_*)
F16 = 4                                # host-side only: the hi half of an F16X3 operand used alone (one fp16 product)
FP16_KINDS = (F16X3, F16)
KIND_OF = {"fp32": F32, "bf16": BF16, "bf16x3": BF16X3, "fp16x3": F16X3}

LAUNCHES = 0  # number of native kernel launches issued through this module (bench.py reads it)


def _bump(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _ptr(t: torch.Tensor | None):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(t: torch.Tensor, dtype=torch.float32, name="tensor") -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"laplace_b200 kernels need CUDA tensors ({name} is on {t.device}); there is no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t


@dataclass
class Packed:
    """K-major operand ``[rows, ldk]`` (contraction index contiguous)."""

    hi: torch.Tensor
    lo: torch.Tensor | None
    kind: int
    rows: int
    K: int

    @property
    def ldk(self) -> int:
        return self.hi.shape[1]


def hi_only(P: Packed) -> Packed:
    """The leading 16-bit half of a hi/lo operand as a single-product operand (no copy)."""
    if P.lo is None:
        return P
    return Packed(P.hi, None, F16 if P.kind == F16X3 else BF16, P.rows, P.K)


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def alloc_packed(rows: int, K: int, kind: int, device) -> Packed:
    ldk = max(round_up(K, 8), 8)
    if kind == F32:
        return Packed(torch.empty(rows, ldk, device=device, dtype=torch.float32), None, kind, rows, K)
    dt = torch.float16 if kind == F16X3 else torch.bfloat16
    hi = torch.empty(rows, ldk, device=device, dtype=dt)
    lo = torch.empty(rows, ldk, device=device, dtype=dt) if kind in (BF16X3, F16X3) else None
    return Packed(hi, lo, kind, rows, K)


def device_info():
    import ctypes as C

    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    _lib.call("lpb_device_info", C.byref(sm), C.byref(maj), C.byref(mnr))
    return sm.value, maj.value, mnr.value


def set_gemm_tile_mode(mode: int) -> None:
    """-1 automatic, 0 one 128x128 tile per CTA, 1 256x256 CTA-pair tiles whenever M, N >= 256, 2 persistent CTAs."""
    _lib.call("lpb_set_gemm_tile_mode", int(mode))


def set_mask_major_min(min_mask_elems: int) -> None:
    """Mask size from which ``pack_cast_fused`` / ``relu_bwd`` switch to their mask-major kernels (``lpb_set_mask_major_min``;
    default 4 Mi elements, negative: never).  Same results bit for bit; for tests and A/B timing."""
    _lib.call("lpb_set_mask_major_min", int(min_mask_elems))


# ------------------------------------------------------------------------------ pack
def pack_rows(src: torch.Tensor, kind: int, out: Packed | None = None, k0: int = 0, scale: float = 1.0,
              square: bool = False, row_scale: torch.Tensor | None = None, nrep: int = 1, total_K: int | None = None) -> Packed:
    """``src [K, d]`` fp32 (last dim contiguous) -> K-major ``[nrep*d, ldk]`` at columns ``k0..k0+K``."""
    _check(src, name="src")
    assert src.dim() == 2 and src.stride(1) == 1
    K, d = src.shape
    if out is None:
        out = alloc_packed(nrep * d, total_K if total_K is not None else K, kind, src.device)
    if row_scale is not None:
        _check(row_scale, name="row_scale")
        assert row_scale.is_contiguous() and row_scale.numel() == nrep * K
    _lib.call("lpb_pack_rows_t", _ptr(src), K, d, src.stride(0), _ptr(row_scale), nrep, scale, 1 if square else 0,
              _ptr(out.hi), _ptr(out.lo), out.kind, out.ldk, k0, _stream())
    _bump()
    return out


def conv_out_hw(x_shape, mod) -> tuple[int, int]:
    H, W = x_shape[-2:]
    kh, kw = mod.kernel_size
    sh, sw = mod.stride
    ph, pw = mod.padding
    dh, dw = mod.dilation
    return (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1, (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1


def pack_conv(x: torch.Tensor, mod, kind: int, reduce_mean: bool = False, square: bool = False) -> tuple[Packed, int]:
    """Unfolded conv patches, K-major ``[C_in*kh*kw, N*OH*OW]`` (or ``[.., N]`` means for KFAC-reduce).
    Returns the packed operand and ``T = OH*OW``."""
    _check(x, name="x")
    x = x.contiguous()
    if isinstance(mod.padding, str):
        raise ValueError("string padding modes are not supported")
    N, Cin, H, W = x.shape
    OH, OW = conv_out_hw(x.shape, mod)
    kh, kw = mod.kernel_size
    K = N if reduce_mean else N * OH * OW
    out = alloc_packed(Cin * kh * kw, K, kind, x.device)
    _lib.call("lpb_pack_conv2d_t", _ptr(x), N, Cin, H, W, kh, kw, mod.stride[0], mod.stride[1], mod.padding[0],
              mod.padding[1], mod.dilation[0], mod.dilation[1], 1.0, 1 if square else 0, 1 if reduce_mean else 0,
              _ptr(out.hi), _ptr(out.lo), out.kind, out.ldk, 0, _stream())
    _bump()
    return out, OH * OW


def pack_nchw(g: torch.Tensor, kind: int, reduce_sum: bool = False, square: bool = False) -> Packed:
    """``g [Nn, Cc, HW]`` fp32 contiguous -> K-major ``[Cc, Nn*HW]`` (or ``[Cc, Nn]`` sums)."""
    _check(g, name="g")
    assert g.dim() == 3 and g.is_contiguous()
    Nn, Cc, HW = g.shape
    K = Nn if reduce_sum else Nn * HW
    out = alloc_packed(Cc, K, kind, g.device)
    _lib.call("lpb_pack_nchw_t", _ptr(g), Nn, Cc, HW, 1.0, 1 if square else 0, 1 if reduce_sum else 0, _ptr(out.hi),
              _ptr(out.lo), out.kind, out.ldk, 0, _stream())
    _bump()
    return out


# ------------------------------------------------------------------------------ contractions
def gemm_nt(A: Packed, B: Packed, out: torch.Tensor, alpha: float = 1.0, accumulate: bool = True,
            symmetric: bool = False) -> torch.Tensor:
    """``out[M,N] (+)= alpha * A[M,K] @ B[N,K]^T`` ; fp32 SIMT or tcgen05 bf16 / bf16x3 by operand kind."""
    _check(out, name="out")
    assert out.dim() == 2 and out.stride(1) == 1
    M, N = out.shape
    assert A.rows == M and B.rows == N and A.K == B.K and A.kind == B.kind, (A.rows, M, B.rows, N, A.K, B.K)
    if symmetric:
        assert A.hi.data_ptr() == B.hi.data_ptr() and M == N
    if A.kind == F32:
        _lib.call("lpb_gemm_nt_f32", _ptr(A.hi), A.ldk, _ptr(B.hi), B.ldk, M, N, A.K, alpha, 1 if accumulate else 0,
                  _ptr(out), out.stride(0), 1 if symmetric else 0, _stream())
    else:
        _lib.call("lpb_gemm_nt_tc", _ptr(A.hi), _ptr(A.lo), A.ldk, _ptr(B.hi), _ptr(B.lo), B.ldk, M, N, A.K, alpha,
                  1 if accumulate else 0, _ptr(out), out.stride(0), 1 if symmetric else 0, 1 if A.kind == F16X3 else 0,
                  _stream())
    _bump()
    return out


def shared_weight_contract(mode: int, G: Packed, A: Packed, d_out: int, d_in: int, T: int, Nn: int, ncols: int,
                           out: torch.Tensor, scale: float = 1.0, out_ld: int = 0, js_stride_n: int = 0,
                           js_stride_c: int = 0) -> None:
    assert G.kind == F32 and A.kind == F32
    _check(out, name="out")
    _lib.call("lpb_shared_weight_contract", mode, _ptr(G.hi), G.ldk, _ptr(A.hi), A.ldk, d_out, d_in, T, Nn, ncols,
              scale, _ptr(out), out_ld, js_stride_n, js_stride_c, _stream())
    _bump()


def kron_conv_quadform(Gt: torch.Tensor, At: torch.Tensor, T: int, Nn: int, C: int, l1: torch.Tensor, l2: torch.Tensor,
                       delta: float, damping: bool, out: torch.Tensor) -> torch.Tensor:
    """``out[n,c,k] += sum_ij w(i,j) Z_c[i,j] Z_k[i,j]`` with ``Z_c = sum_t Gt[:, (c,n,t)] At[:, (n,t)]^T`` formed tile by
    tile in shared memory (``lpb_kron_conv_quadform``).  ``Gt [d_out, C*Nn*T]``, ``At [d_in, Nn*T]`` K-major fp32
    (eigenbasis-rotated rows), ``l1 [d_out]``, ``l2 [d_in]`` eigenvalues, ``out [Nn, C, C]``."""
    _check(Gt, name="Gt"), _check(At, name="At"), _check(out, name="out"), _check(l1, name="l1"), _check(l2, name="l2")
    assert Gt.dim() == 2 and At.dim() == 2 and Gt.stride(1) == 1 and At.stride(1) == 1 and out.is_contiguous()
    d_out, d_in = Gt.shape[0], At.shape[0]
    assert Gt.shape[1] >= C * Nn * T and At.shape[1] >= Nn * T and out.shape == (Nn, C, C)
    assert l1.numel() == d_out and l2.numel() == d_in and l1.is_contiguous() and l2.is_contiguous()
    for n0 in range(0, Nn, 65535):      # gridDim.z
        nn = min(65535, Nn - n0)
        _lib.call("lpb_kron_conv_quadform", Gt.data_ptr() + 4 * n0 * T, Gt.stride(0), Nn * T, At.data_ptr() + 4 * n0 * T,
                  At.stride(0), d_out, d_in, T, nn, C, _ptr(l1), _ptr(l2), float(delta), 1 if damping else 0,
                  out.data_ptr() + 4 * n0 * C * C, _stream())
        _bump()
    return out


def jac_linear_write(g: torch.Tensor, a: torch.Tensor, Js_view: torch.Tensor, stride_n: int, stride_c: int, off_w: int,
                     off_b: int) -> None:
    """``g [C, Nn, d_out]``, ``a [Nn, d_in]`` contiguous; ``Js_view`` = base pointer tensor of the Jacobian rows."""
    _check(g, name="g"), _check(a, name="a"), _check(Js_view, name="Js")
    assert g.is_contiguous() and a.is_contiguous()
    Cc, Nn, d_out = g.shape
    d_in = a.shape[1]
    _lib.call("lpb_jac_linear_write", _ptr(g), _ptr(a), Nn, Cc, d_out, d_in, _ptr(Js_view), stride_n, stride_c, off_w,
              off_b, _stream())
    _bump()


def ll_jacobian_write(phi: torch.Tensor, C_out: int, has_bias: bool) -> torch.Tensor:
    _check(phi, name="phi")
    phi = phi.contiguous()
    Nn, D = phi.shape
    P = C_out * D + (C_out if has_bias else 0)
    Js = torch.empty(Nn, C_out, P, device=phi.device, dtype=torch.float32)
    _lib.call("lpb_ll_jacobian_write", _ptr(phi), Nn, C_out, D, 1 if has_bias else 0, _ptr(Js), _stream())
    _bump()
    return Js


def batched_pair_dot(X: torch.Tensor, Z: torch.Tensor, m: torch.Tensor | None, out: torch.Tensor,
                     accumulate: bool = False) -> torch.Tensor:
    """``out[n,c,k] (+)= sum_i X[n,c,i] Z[n,k,i] m[n,i]`` for 3-D (possibly permuted) views ``X [Nn,CX,d]``,
    ``Z [Nn,CZ,d]`` whose last dim is contiguous; ``m [Nn,d]`` or ``[d]`` contiguous."""
    _check(X, name="X"), _check(Z, name="Z"), _check(out, name="out")
    assert X.stride(2) == 1 and Z.stride(2) == 1 and out.is_contiguous()
    Nn, CX, d = X.shape
    CZ = Z.shape[1]
    assert Z.shape[0] == Nn and Z.shape[2] == d and out.shape == (Nn, CX, CZ)
    m_stride = 0
    if m is not None:
        _check(m, name="m")
        assert m.is_contiguous() and m.shape[-1] == d
        m_stride = d if m.dim() == 2 else 0
    _lib.call("lpb_batched_pair_dot", _ptr(X), _ptr(Z), _ptr(m), m_stride, Nn, CX, CZ, d, X.stride(0), X.stride(1),
              Z.stride(0), Z.stride(1), 1 if accumulate else 0, _ptr(out), _stream())
    _bump()
    return out


def ll_ggn_expand(G: torch.Tensor, C_out: int, D: int, has_bias: bool, H: torch.Tensor, accumulate: bool) -> None:
    _check(G, name="G"), _check(H, name="H")
    assert G.is_contiguous() and H.is_contiguous()
    _lib.call("lpb_ll_ggn_expand", _ptr(G), C_out, D, 1 if has_bias else 0, 1 if accumulate else 0, _ptr(H), _stream())
    _bump()


def ll_sigma_gather(Sigma: torch.Tensor, C_out: int, D: int, has_bias: bool) -> torch.Tensor:
    _check(Sigma, name="Sigma")
    Sigma = Sigma.contiguous()
    Dt = D + (1 if has_bias else 0)
    Sg = torch.empty(C_out * C_out * Dt, Dt, device=Sigma.device, dtype=torch.float32)
    _lib.call("lpb_ll_sigma_gather", _ptr(Sigma), C_out, D, 1 if has_bias else 0, _ptr(Sg), _stream())
    _bump()
    return Sg


EIGH_MAX_N = 128


def eigh_jacobi(A: torch.Tensor, max_sweeps: int = 30):
    """Batched symmetric eigendecomposition ``A [batch, n, n]`` (n <= 128): ascending clamped eigenvalues, Q columns."""
    _check(A, name="A")
    A = A.contiguous()
    batch, n, _ = A.shape
    ev = torch.empty(batch, n, device=A.device, dtype=torch.float32)
    Q = torch.empty(batch, n, n, device=A.device, dtype=torch.float32)
    _lib.call("lpb_eigh_jacobi", _ptr(A), batch, n, _ptr(ev), _ptr(Q), max_sweeps, _stream())
    _bump()
    return ev, Q


# ------------------------------------------------------------------------------ convolution engine operands
def alloc_rows(rows: int, cols: int, kind: int, device) -> Packed:
    """Row-major operand ``[rows, ld]`` whose *columns* are the contraction index (ld multiple of 8)."""
    return alloc_packed(rows, cols, kind, device)


def _conv_args(shape, mod):
    N, C, H, W = shape
    return (N, C, H, W, mod.kernel_size[0], mod.kernel_size[1], mod.stride[0], mod.stride[1], mod.padding[0],
            mod.padding[1], mod.dilation[0], mod.dilation[1])


def pack_conv_rows(x: torch.Tensor, mod, kind: int) -> Packed:
    """Patch-major im2col ``[(n,oh,ow), C_in*kh*kw]`` (contraction over the patch index)."""
    _check(x, name="x")
    x = x.contiguous()
    N, Cin, H, W = x.shape
    OH, OW = conv_out_hw(x.shape, mod)
    out = alloc_rows(N * OH * OW, Cin * mod.kernel_size[0] * mod.kernel_size[1], kind, x.device)
    _lib.call("lpb_pack_conv2d_rows", _ptr(x), *_conv_args(x.shape, mod), _ptr(out.hi), _ptr(out.lo), out.kind, out.ldk,
              _stream())
    _bump()
    return out


def pack_nchw_rows(g: torch.Tensor, kind: int) -> Packed:
    """``g [Q, Cc, HW]`` -> ``[(q,hw), Cc]`` (contraction over channels)."""
    _check(g, name="g")
    assert g.dim() == 3 and g.is_contiguous()
    Q, Cc, HW = g.shape
    out = alloc_rows(Q * HW, Cc, kind, g.device)
    _lib.call("lpb_pack_nchw_rows", _ptr(g), Q, Cc, HW, _ptr(out.hi), _ptr(out.lo), out.kind, out.ldk, _stream())
    _bump()
    return out


def pack_cast(src: torch.Tensor, kind: int) -> Packed:
    """``src [rows, cols]`` fp32 -> same layout as a K-major operand (contraction over ``cols``)."""
    _check(src, name="src")
    assert src.dim() == 2 and src.stride(1) == 1
    rows, cols = src.shape
    out = alloc_rows(rows, cols, kind, src.device)
    _lib.call("lpb_pack_cast", _ptr(src), rows, cols, src.stride(0), _ptr(out.hi), _ptr(out.lo), out.kind, out.ldk,
              _stream())
    _bump()
    return out


def pack_cast_fused(src: torch.Tensor, kind: int, scale: torch.Tensor | None = None, y: torch.Tensor | None = None) -> Packed:
    """``pack_cast(src * scale[None, :] * (y > 0))`` in one pass: ``src [rows, cols]`` fp32, ``scale [cols]`` (frozen
    BatchNorm as an affine map), ``y [rows_y, cols]`` (ReLU output, ``rows % rows_y == 0``: shared by the folded
    curvature columns).  Either may be ``None``."""
    _check(src, name="src")
    assert src.dim() == 2 and src.stride(1) == 1
    rows, cols = src.shape
    rows_y = ld_y = 0
    if y is not None:
        _check(y, name="y")
        assert y.dim() == 2 and y.stride(1) == 1 and y.shape[1] == cols and rows % y.shape[0] == 0
        rows_y, ld_y = y.shape[0], y.stride(0)
    if scale is not None:
        _check(scale, name="scale")
        scale = scale.contiguous()
        assert scale.numel() == cols
    out = alloc_rows(rows, cols, kind, src.device)
    _lib.call("lpb_pack_cast_fused", _ptr(src), rows, cols, src.stride(0), _ptr(scale), _ptr(y), rows_y, ld_y, _ptr(out.hi),
              _ptr(out.lo), out.kind, out.ldk, _stream())
    _bump()
    return out


def col2im(Dc: torch.Tensor, in_shape, mod) -> torch.Tensor:
    """``Dc [C_in*kh*kw, Q*OH*OW]`` -> ``grad_in [Q, C_in, H, W]``."""
    _check(Dc, name="Dc")
    assert Dc.dim() == 2 and Dc.stride(1) == 1
    out = torch.empty(tuple(in_shape), device=Dc.device, dtype=torch.float32)
    _lib.call("lpb_col2im", _ptr(Dc), Dc.stride(0), *_conv_args(in_shape, mod), _ptr(out), _stream())
    _bump()
    return out


def col2im_nhwc(Dc: torch.Tensor, in_shape, mod) -> torch.Tensor:
    """``Dc [Q*OH*OW, kh*kw*C_in]`` (tap-major columns) -> ``grad_in [Q, C_in, H, W]`` as a channels-last view."""
    _check(Dc, name="Dc")
    assert Dc.dim() == 2 and Dc.stride(1) == 1
    Q, Cin, H, W = in_shape
    out = torch.empty(Q, H, W, Cin, device=Dc.device, dtype=torch.float32)
    _lib.call("lpb_col2im_nhwc", _ptr(Dc), Dc.stride(0), *_conv_args(in_shape, mod), _ptr(out), _stream())
    _bump()
    return out.permute(0, 3, 1, 2)


def conv_nhwc(X: Packed, Q: int, H: int, W: int, Wt: Packed, N: int, KH: int, KW: int, base_h: int, base_w: int, sgn: int,
              out: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """Implicit-GEMM stride-1 convolution on NHWC bf16(hi/lo) rows ``X [(q,h,w), Kc]`` with tap-major weights
    ``Wt [(tap, n), Kc]``; ``out [(q,h,w), N]`` fp32 is overwritten."""
    _check(out, name="out")
    # operand rows may be given as their hi half alone against hi/lo weights: two products instead of three
    lean = X.lo is None and Wt.lo is not None and (X.kind, Wt.kind) in ((BF16, BF16X3), (F16, F16X3))
    assert X.kind in (BF16, BF16X3, F16X3, F16) and (X.kind == Wt.kind or lean)
    assert X.rows == Q * H * W and Wt.rows == KH * KW * N and X.K == Wt.K
    assert out.shape == (Q * H * W, N) and out.stride(1) == 1
    _lib.call("lpb_conv_nhwc_tc", _ptr(X.hi), _ptr(X.lo), Q, H, W, X.K, X.ldk, _ptr(Wt.hi), _ptr(Wt.lo), Wt.ldk, N, KH, KW,
              base_h, base_w, sgn, alpha, _ptr(out), out.stride(0), 1 if X.kind in FP16_KINDS else 0, _stream())
    _bump()
    return out


def conv_bwd_strided(G: Packed, Q: int, OH: int, OW: int, Wt: Packed, mod, in_shape) -> torch.Tensor:
    """Input gradient of a strided convolution from the output-gradient rows ``G [(q,oh,ow), C_out]`` and the tap-major
    weights ``Wt [(kh,kw,ci), C_out]``: one implicit GEMM per stride parity class, written in place into the NHWC result
    (returned as a channels-last view ``[Q, C_in, H, W]``)."""
    _, Ci, H, W = in_shape
    kh, kw = mod.kernel_size
    lean = G.lo is None and Wt.lo is not None and (G.kind, Wt.kind) == (BF16, BF16X3)
    assert G.kind in (BF16, BF16X3) and (G.kind == Wt.kind or lean) and G.rows == Q * OH * OW and Wt.rows == kh * kw * Ci and G.K == Wt.K
    out = torch.empty(Q, H, W, Ci, device=G.hi.device, dtype=torch.float32)
    _lib.call("lpb_conv_bwd_strided_tc", _ptr(G.hi), _ptr(G.lo), Q, OH, OW, G.K, G.ldk, _ptr(Wt.hi), _ptr(Wt.lo), Wt.ldk, Ci,
              kh, kw, mod.stride[0], mod.stride[1], mod.padding[0], mod.padding[1], H, W, _ptr(out), out.stride(2), _stream())
    _bump(mod.stride[0] * mod.stride[1])
    return out.permute(0, 3, 1, 2)


def gemm_tn(A: Packed, B: Packed, out: torch.Tensor, alpha: float = 1.0, accumulate: bool = True,
            symmetric: bool = False) -> torch.Tensor:
    """``out[M,N] (+)= alpha * A^T B`` for ROW-major 16-bit operands ``A [K_rows, M]``, ``B [K_rows, N]``
    (``Packed.rows`` = sample rows, ``Packed.K`` = features): the tcgen05 kernel with MN-major descriptors."""
    _check(out, name="out")
    M, N = out.shape
    assert A.kind in (BF16, BF16X3, F16X3, F16) and A.kind == B.kind and A.K == M and B.K == N and A.rows == B.rows
    if symmetric:
        assert A.hi.data_ptr() == B.hi.data_ptr() and M == N
    _lib.call("lpb_gemm_tn_tc", _ptr(A.hi), _ptr(A.lo), A.ldk, _ptr(B.hi), _ptr(B.lo), B.ldk, M, N, A.rows, alpha,
              1 if accumulate else 0, _ptr(out), out.stride(0), 1 if symmetric else 0, 1 if A.kind in FP16_KINDS else 0, _stream())
    _bump()
    return out


def conv_patches_ok(Ci: int, H: int, W: int, kh: int, kw: int) -> bool:
    """Shapes ``syrk_conv_patches`` accepts (images tiling 64-row chunks, <= 9 taps) and pays off on (channel padding
    to 64 per tap wastes at most ~1/3 of the tensor work)."""
    if kh * kw > 9 or Ci < 48 or (-(-Ci // 64) * 64) > 1.34 * Ci:
        return False
    hw = H * W
    return (64 % hw == 0) if hw < 64 else (64 % W == 0 and H % (64 // W) == 0)


def syrk_conv_patches(X: Packed, Q: int, H: int, W: int, mod, out: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """``out[d_in, d_in] += alpha * P^T P`` for the im2col patch matrix ``P [(n,h,w), C_in*kh*kw]`` of a stride-1 'same'
    convolution, computed from the NHWC rows ``X [(n,h,w), C_in]`` without forming ``P``: implicit (shifted 4-D TMA)
    SYRK into a tap-major scratch factor, then a permuting accumulate into the parameter order ``(ci,kh,kw)``."""
    _check(out, name="out")
    kh, kw = mod.kernel_size
    Ci = X.K
    d = Ci * kh * kw
    Ci_pad = -(-Ci // 64) * 64
    ph, pw = mod.padding
    live = sum(1 for a in range(kh) for b in range(kw) if abs(a - ph) < H and abs(b - pw) < W)   # == lpb_conv_live_taps
    dp = Ci_pad * live
    assert out.shape == (d, d) and X.rows == Q * H * W and X.kind in (BF16, BF16X3, F16X3, F16)
    T = torch.empty(dp, dp, device=out.device, dtype=torch.float32)
    _lib.call("lpb_syrk_conv_patches_tc", _ptr(X.hi), _ptr(X.lo), X.ldk, Q, H, W, Ci, kh, kw, mod.padding[0], mod.padding[1],
              alpha, 0, _ptr(T), T.stride(0), 1 if X.kind in FP16_KINDS else 0, _stream())
    _lib.call("lpb_taps_to_param_accumulate", _ptr(T), T.stride(0), Ci, Ci_pad, kh, kw, ph, pw, H, W, _ptr(out), out.stride(0),
              _stream())
    _bump(2)
    return out


def diag_conv_ok(Ci: int, H: int, W: int, kh: int, kw: int) -> bool:
    """Shapes ``diag_conv_sq`` accepts: one TMEM accumulation per sample needs whole 64-pixel chunks per image."""
    hw = H * W
    return hw >= 64 and 64 % W == 0 and H % (64 // W) == 0 and hw <= 4096 and Ci >= 16


def diag_conv_sq(G: Packed, X: Packed, Nimg: int, H: int, W: int, mod, out: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """``out[co, (ci,kh,kw)] += alpha * sum_q (per-sample weight gradient)^2`` of a stride-1 'same' convolution on the
    tensor cores: ``G [(q,h,w), C_out]`` gradient rows of all folded columns (``q = col * Nimg + n``), ``X [(n,h,w), C_in]``
    NHWC input rows, both bf16 hi/lo.  Per-sample gradients are formed in TMEM, squared by the epilogue, never stored."""
    _check(out, name="out")
    kh, kw = mod.kernel_size
    Ci, Co = X.K, G.K
    Ci_pad = -(-Ci // 64) * 64
    assert out.shape == (Co, Ci * kh * kw) and X.rows == Nimg * H * W and G.rows % X.rows == 0 and G.kind == X.kind
    assert G.kind in (BF16, BF16X3)
    Dt = torch.empty(Co, kh * kw * Ci_pad, device=out.device, dtype=torch.float32)
    _lib.call("lpb_diag_conv_sq_tc", _ptr(G.hi), _ptr(G.lo), G.ldk, _ptr(X.hi), _ptr(X.lo), X.ldk, G.rows // (H * W), Nimg, H, W,
              Ci, Co, kh, kw, mod.padding[0], mod.padding[1], alpha, 0, _ptr(Dt), Dt.stride(0), _stream())
    _lib.call("lpb_taps_to_param_rect", _ptr(Dt), Dt.stride(0), Co, Ci, Ci_pad, kh * kw, _ptr(out), out.stride(0), _stream())
    _bump(2)
    return out


# ------------------------------------------------------------------------------ reverse-pass element-wise maps
def scale_channels(g: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """``g [Q, C, H, W]`` (NCHW- or channels_last-dense) times a per-channel ``scale [C]``; same layout out."""
    _check(g, name="g"), _check(scale, name="scale")
    Q, C, H, W = g.shape
    if g.is_contiguous():
        inner = H * W
    elif g.is_contiguous(memory_format=torch.channels_last):
        inner = 1
    else:
        g, inner = g.contiguous(), H * W
    out = torch.empty_like(g)
    _lib.call("lpb_scale_channels", _ptr(g), _ptr(scale.contiguous()), _ptr(out), g.numel(), C, inner, _stream())
    _bump()
    return out


def relu_bwd(g: torch.Tensor, y: torch.Tensor, reps: int) -> torch.Tensor:
    """``g [reps*B, ...]`` masked by the forward output ``y [B, ...]`` (same dense layout per image block)."""
    _check(g, name="g"), _check(y, name="y")
    out = torch.empty_like(g)
    _lib.call("lpb_relu_bwd", _ptr(g), _ptr(y), _ptr(out), y.numel(), reps, _stream())
    _bump()
    return out


def maxpool2d_bwd(g: torch.Tensor, idx: torch.Tensor, in_shape, k: int, s: int, p: int) -> torch.Tensor:
    """``g [Q, C, OH, OW]``, ``idx [Nb, C, OH, OW]`` int64 argmax -> ``[Q, C, H, W]``.  Channels-last ``g`` and ``idx``
    stay channels-last (no layout copies); anything else goes through the NCHW kernel."""
    _check(g, name="g")
    Q, C, OH, OW = g.shape
    H, W = in_shape[-2:]
    cl = torch.channels_last
    if C > 1 and g.is_contiguous(memory_format=cl) and idx.is_contiguous(memory_format=cl) and not g.is_contiguous():
        out = torch.empty(Q, H, W, C, device=g.device, dtype=torch.float32)
        _lib.call("lpb_maxpool2d_bwd_nhwc", _ptr(g), _ptr(idx), _ptr(out), Q, idx.shape[0], C, H, W, OH, OW, k, s, p,
                  _stream())
        _bump()
        return out.permute(0, 3, 1, 2)
    g = g.contiguous()
    idx = idx.contiguous()
    Q, C, OH, OW = g.shape
    H, W = in_shape[-2:]
    out = torch.empty(Q, C, H, W, device=g.device, dtype=torch.float32)
    _lib.call("lpb_maxpool2d_bwd", _ptr(g), _ptr(idx), _ptr(out), Q, idx.shape[0], C, H, W, OH, OW, k, s, p, _stream())
    _bump()
    return out


def maxpool2d_bwd_pack(g: torch.Tensor, idx: torch.Tensor, in_shape, k: int, s: int, p: int, scale: torch.Tensor | None,
                       y: torch.Tensor | None) -> Packed:
    """Un-pooled gradient of a max-pool times the ReLU mask ``y > 0`` and the channel ``scale``, emitted directly as bf16
    hi/lo operand rows ``[(q, h, w), C]`` (``lpb_maxpool2d_bwd_pack_nhwc``).  ``g [Q, C, OH, OW]``, ``idx [Nb, C, OH, OW]`` and
    ``y [Nb, C, H, W]`` channels-last."""
    _check(g, name="g")
    Q, C, OH, OW = g.shape
    H, W = in_shape[-2:]
    cl = torch.channels_last
    assert g.is_contiguous(memory_format=cl) and idx.is_contiguous(memory_format=cl) and C % 8 == 0
    if y is not None:
        _check(y, name="y")
        assert y.is_contiguous(memory_format=cl) and tuple(y.shape) == (idx.shape[0], C, H, W)
    if scale is not None:
        _check(scale, name="scale")
        scale = scale.contiguous()
    out = alloc_rows(Q * H * W, C, BF16X3, g.device)
    _lib.call("lpb_maxpool2d_bwd_pack_nhwc", _ptr(g), _ptr(idx), _ptr(scale), _ptr(y), _ptr(out.hi), _ptr(out.lo), out.ldk, Q,
              idx.shape[0], C, H, W, OH, OW, k, s, p, _stream())
    _bump()
    return out
