"""Convolution engine: forward and backward-data of ``nn.Conv2d`` as GEMMs on the tcgen05 kernel.

Why it exists.  The curvature columns have to be propagated through the network in *true* fp32: PyTorch's
default lets cuDNN use TF32 for convolutions, which perturbs the per-layer gradients by ~1e-3 and breaks the
1e-4 parity gate on the KFAC/GGN factors; with TF32 disabled cuDNN falls back to fp32 SIMT / FFT kernels that
run the C-times-batched reverse pass of a ResNet-18 at ~3 TFLOP/s (78 % of a KFAC step, profiles/).  The
engine keeps autograd for the graph and every element-wise / pooling / normalisation op, but computes the two
convolution products itself:

* forward        stride-1 "same" convolutions: implicit GEMM on NHWC rows (shifted 4-D TMA boxes, no im2col);
                 anything else: patch-major im2col + GEMM-NT ``out[(n,t), co] = patches[(n,t), :] . W[co, :]``
* backward-data  stride 1: the same implicit GEMM with flipped taps; stride > 1: one implicit GEMM per stride parity
                 class, written in place into the NHWC input gradient; anything else:
                 ``Dc[(q,t), (kh,kw,ci)] = g[(q,t), :] . W^T[(kh,kw,ci), :]`` + ``col2im_nhwc`` gather

both with 16-bit hi/lo operands and three tensor-core products per tile (relative error ~2^-16, fp32
accumulation), i.e. fp32-accurate at tensor-core speed.  Everything the engine emits or consumes is
channels-last.  The reverse pass for all curvature columns is ONE call per layer: ``torch.func.vmap`` over
``autograd.grad`` reaches ``_ConvBwdData.vmap`` (or the fused ``_ConvFusedBwd.vmap``), which folds the column
dimension into the batch.
"""
from __future__ import annotations

import torch
from torch import nn

from . import kernels as K

import math
import os
import weakref

KIND = K.BF16X3       # reverse pass: gradients span many orders of magnitude -> bf16 hi/lo (8-bit exponent)
KIND_FWD = K.F16X3    # forward pass: activations / weights sit inside fp16 range -> fp16 hi/lo (22 bits); the
                      # ~2e-7 forward error keeps ReLU masks identical to an fp32 forward (a 1e-5 error flips a
                      # few masks per 10^5 units, each flip is a 100 % error on that unit's gradient)
# > 0: backward-data convolutions over at least this many gradient rows use the bf16 hi half of the gradient alone
# against the exact (hi + lo) weights -- two tensor-core products instead of three.  OFF by default: each gradient element
# then carries ~2e-3 of relative noise (the 2^-9 rounding of the rows does not average out inside one dot product), which
# costs 1.1e-4 .. 1.8e-4 rel-fro on the KFAC B factors at batch 64 .. 256 (tests/diagnostics/gpu_lean_diag.py).  Kept as a measured
# option (backend.LEAN_BACKWARD_MIN_ROWS) for batches where rows / d_out >= ~2e5.
LEAN_BWD_MIN_ROWS = 0
FUSE_POOL = os.environ.get("LPB_NO_POOL_FUSION") != "1"   # max-pool reverse map fused with the stem chain's operand split
USE_IMPLICIT = os.environ.get("LPB_NO_IMPLICIT") != "1"
USE_STRIDED = os.environ.get("LPB_NO_STRIDED") != "1"     # strided reverse passes as per-parity implicit GEMMs


class _WeightCache:
    """Packed weights are reused across batches (weights are constant during ``fit``).  Entries live in a
    ``WeakKeyDictionary`` keyed by the module (they die with it: no leak across models, no stale hit when a new model
    re-uses a freed module's ``id``) and are valid only for the very tensor object, storage and version they were
    packed from."""

    def __init__(self):
        self.store = weakref.WeakKeyDictionary()

    def get(self, mod: nn.Module, which: str):
        w = mod.weight
        per_mod = self.store.setdefault(mod, {})
        tag = (w.data_ptr(), w._version, tuple(w.shape), w.device)
        hit = per_mod.get(which)
        if hit is not None and hit[0] == tag and hit[2]() is w:
            return hit[1]
        w2 = w.detach().reshape(w.shape[0], -1)
        w2 = w2 if w2.dtype == torch.float32 else w2.float()
        scale = 1.0
        if which in ("fwd", "fwd_taps"):
            # fp16 hi/lo operands: weights are tiny (|w| <= 1/sqrt(fan_in)), their lo halves would fall into the fp16
            # subnormals -- pre-scale by a power of two so that max|w| sits near 2^10 (one host sync per weight version)
            amax = float(w2.abs().max())
            scale = 2.0 ** math.floor(math.log2(1024.0 / amax)) if amax > 0 else 1.0
        if which == "fwd":
            packed = K.pack_cast((w2 * scale).contiguous(), KIND_FWD)
        elif which == "bwd":
            packed = K.pack_rows(w2.contiguous(), KIND)
        else:  # tap-major weights of the implicit-GEMM kernel
            w4 = w.detach().float() * scale
            perm = (2, 3, 0, 1) if which == "fwd_taps" else (2, 3, 1, 0)
            w4 = w4.permute(*perm).reshape(-1, w4.shape[1] if which == "fwd_taps" else w4.shape[0]).contiguous()
            packed = K.pack_cast(w4, KIND_FWD if which == "fwd_taps" else KIND)
        packed.inv_scale = 1.0 / scale
        per_mod[which] = (tag, packed, weakref.ref(w))
        return packed


_CACHE = _WeightCache()

# Row-major 16-bit operands produced while running the passes of the current batch, keyed by id(module):
#   "P": patch rows [(n,t), d_in] of explicit-path forwards, "G": output-gradient rows [(col,n,t), C_out].
# The KFAC factor SYRKs consume them directly through the MN-major GEMM (no second, transposing pack).
STASH: dict = {}
# Reverse-pass counter (bumped by the backend before every autograd.grad call): a layer whose gradient rows are packed
# twice within one pass has two consumers in the graph, which the fused chains cannot represent.
PASS_ID = [0]


class FusionConflict(RuntimeError):
    """A fused conv -> BatchNorm -> ReLU chain whose intermediate is also consumed elsewhere (e.g. a pre-activation
    residual that adds the raw convolution output).  The backend catches it and repeats the batch unfused."""


def implicit_ok(mod: nn.Conv2d, H: int, W: int) -> bool:
    """Stride-1 'same' convolutions whose images tile 128-row MMA blocks -- whole images (``H*W`` divides 128) or
    ``128 / W`` image rows of a larger image: no im2col / col2im at all."""
    kh, kw = mod.kernel_size
    if not USE_IMPLICIT:
        return False
    if not (tuple(mod.stride) == (1, 1) and tuple(mod.dilation) == (1, 1) and 2 * mod.padding[0] == kh - 1
            and 2 * mod.padding[1] == kw - 1):
        return False
    if H * W <= 128:
        return 128 % (H * W) == 0
    return W <= 128 and 128 % W == 0 and H % (128 // W) == 0


def strided_ok(mod: nn.Conv2d, H: int, W: int) -> bool:
    """Strided convolutions whose reverse pass runs as implicit GEMMs per stride parity class: the input extent is an
    exact multiple of the stride and the output grid tiles 128-row MMA blocks."""
    if not USE_IMPLICIT or not USE_STRIDED or tuple(mod.dilation) != (1, 1) or tuple(mod.stride) == (1, 1):
        return False
    sh, sw = mod.stride
    kh, kw = mod.kernel_size
    if H % sh or W % sw or kh * kw > 64:
        return False
    OH, OW = (H + 2 * mod.padding[0] - kh) // sh + 1, (W + 2 * mod.padding[1] - kw) // sw + 1
    if OH * sh != H or OW * sw != W:
        return False
    if OH * OW <= 128:
        return 128 % (OH * OW) == 0
    return OW <= 128 and 128 % OW == 0 and OH % (128 // OW) == 0


def nhwc_rows(x: torch.Tensor, kind: int) -> K.Packed:
    """``x [N, C, H, W]`` (NCHW or channels_last) -> 16-bit hi/lo rows ``[(n,h,w), C]``."""
    N, C, H, W = x.shape
    if x.stride(1) == 1 and x.permute(0, 2, 3, 1).is_contiguous():
        return K.pack_cast(x.permute(0, 2, 3, 1).reshape(N * H * W, C), kind)
    return K.pack_nchw_rows(x.contiguous().reshape(N, C, H * W), kind)


def _implicit_rows(X: K.Packed, N: int, H: int, W: int, mod: nn.Conv2d, which: str, n_out: int, sgn: int) -> torch.Tensor:
    kh, kw = mod.kernel_size
    Wt = _CACHE.get(mod, which)
    out = torch.empty(N * H * W, n_out, device=X.hi.device, dtype=torch.float32)
    ph, pw = mod.padding
    K.conv_nhwc(X, N, H, W, Wt, n_out, kh, kw, -sgn * ph, -sgn * pw, sgn, out, alpha=getattr(Wt, "inv_scale", 1.0))
    return out.view(N, H, W, n_out).permute(0, 3, 1, 2)     # channels_last view, no copy


def _implicit(x: torch.Tensor, mod: nn.Conv2d, which: str, n_out: int, sgn: int, X: K.Packed | None = None) -> torch.Tensor:
    N, _, H, W = x.shape
    if X is None:
        X = nhwc_rows(x, KIND_FWD if sgn > 0 else KIND)
    if sgn > 0:
        # forward: the packed NHWC input rows double as the operand of the implicit A-factor SYRK (no im2col at all)
        STASH.setdefault(id(mod), {})["X"] = (X, N, H, W)
    return _implicit_rows(X, N, H, W, mod, which, n_out, sgn)


def conv_forward(x: torch.Tensor, mod: nn.Conv2d) -> torch.Tensor:
    if implicit_ok(mod, x.shape[2], x.shape[3]):
        out = _implicit(x, mod, "fwd_taps", mod.out_channels, +1)
        return out if mod.bias is None else out + mod.bias.detach().view(1, -1, 1, 1)
    x = x.contiguous()
    N = x.shape[0]
    Co = mod.out_channels
    OH, OW = K.conv_out_hw(x.shape, mod)
    P = K.pack_conv_rows(x, mod, KIND_FWD)                   # [(n,t), d_in]
    STASH.setdefault(id(mod), {})["P"] = P
    Wk = _CACHE.get(mod, "fwd")                              # [Co, d_in]
    out = torch.empty(N * OH * OW, Co, device=x.device, dtype=torch.float32)
    K.gemm_nt(P, Wk, out, getattr(Wk, "inv_scale", 1.0), accumulate=False)
    out = out.view(N, OH, OW, Co).permute(0, 3, 1, 2)        # channels-last view like the implicit path
    if mod.bias is not None:
        out = out + mod.bias.detach().view(1, -1, 1, 1)
    return out


def _backward_from_rows(G: K.Packed, Q: int, T: int, mod: nn.Conv2d, in_shape, need_dx: bool):
    """Input gradient from the packed output-gradient rows ``G [(q,t), Co]`` (stashed for the B-factor SYRK)."""
    st = STASH.setdefault(id(mod), {})
    if st.get("G_pass") == PASS_ID[0]:
        raise FusionConflict("convolution engine: the output of a fused convolution is consumed by more than one "
                             "operation (two reverse passes reached the same layer); use fuse_elementwise=False")
    st["G"], st["G_pass"] = G, PASS_ID[0]
    if not need_dx:
        return None
    Gd = K.hi_only(G) if (0 < LEAN_BWD_MIN_ROWS <= G.rows and G.kind == K.BF16X3) else G
    if implicit_ok(mod, in_shape[2], in_shape[3]):
        return _implicit_rows(Gd, Q, in_shape[2], in_shape[3], mod, "bwd_taps", mod.in_channels, -1)
    if strided_ok(mod, in_shape[2], in_shape[3]):
        OH, OW = K.conv_out_hw(in_shape, mod)
        return K.conv_bwd_strided(Gd, Q, OH, OW, _CACHE.get(mod, "bwd_taps"), mod, (Q,) + tuple(in_shape[1:]))
    Wt = _CACHE.get(mod, "bwd_taps")                         # [(kh,kw,ci), Co]
    Dc = torch.empty(Q * T, Wt.rows, device=G.hi.device, dtype=torch.float32)
    K.gemm_nt(G, Wt, Dc, 1.0, accumulate=False)              # [(q,t), (kh,kw,ci)]
    return K.col2im_nhwc(Dc, (Q,) + tuple(in_shape[1:]), mod)   # channels-last view: every gradient stays NHWC


def conv_backward_data(g: torch.Tensor, mod: nn.Conv2d, in_shape, need_dx: bool = True):
    """Packs the output-gradient rows (stashed for the B-factor SYRK) and, if ``need_dx``, returns the input gradient."""
    if STASH.get(id(mod), {}).get("fused"):
        raise FusionConflict("convolution engine: the output of a fused convolution is also consumed outside the fused "
                             "BatchNorm/ReLU chain; use fuse_elementwise=False")
    Q = g.shape[0]
    T = g.shape[2] * g.shape[3]
    G = nhwc_rows(g, KIND)                                   # [(q,t), Co]
    return _backward_from_rows(G, Q, T, mod, in_shape, need_dx)


def _is_nhwc(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.stride(1) == 1 and t.permute(0, 2, 3, 1).is_contiguous()


def conv_backward_fused(g: torch.Tensor, mod: nn.Conv2d, in_shape, need_dx: bool, scale, y, reps: int):
    """Reverse pass of ``relu?(affine?(conv(x)))`` w.r.t. the convolution: the ReLU mask (``y > 0``, shared by the
    ``reps`` folded columns), the frozen-BN scale and the 16-bit hi/lo operand split happen in ONE pass over the
    gradient (``pack_cast_fused``); the result is both the B-factor operand and the A operand of the backward-data
    convolution.  The fp32 gradients between the three modules are never materialised."""
    Q, Co = g.shape[0], g.shape[1]
    T = g.shape[2] * g.shape[3]
    pre = STASH.get(id(mod), {}).pop("G_pre", None)
    if pre is not None:
        # the producer of this gradient (a fused max-pool reverse map) already emitted the masked / scaled operand rows and
        # handed autograd a storage-less placeholder.  If anything else contributed to the gradient, autograd has summed it
        # into a NEW tensor -- the placeholder is gone and the pre-packed rows would miss that contribution.
        G, pass_id, ptr = pre
        if pass_id == PASS_ID[0] and g.data_ptr() == ptr and all(st == 0 for st in g.stride()) and G.rows == Q * T:
            return _backward_from_rows(G, Q, T, mod, in_shape, need_dx)
        raise FusionConflict("convolution engine: the output of a fused conv/BN/ReLU chain feeds a fused max-pool AND another "
                             "operation; use fuse_elementwise=False")
    if _is_nhwc(g) and (y is None or _is_nhwc(y)):
        y2 = None if y is None else y.permute(0, 2, 3, 1).reshape(-1, Co)
        G = K.pack_cast_fused(g.permute(0, 2, 3, 1).reshape(Q * T, Co), KIND, scale, y2)
    else:   # unusual layouts: the separate kernels
        if y is not None:
            g = _ReluBwd.apply(g, y, reps)
        if scale is not None:
            g = K.scale_channels(g, scale)
        G = nhwc_rows(g, KIND)
    return _backward_from_rows(G, Q, T, mod, in_shape, need_dx)


class _ConvBwdData(torch.autograd.Function):
    @staticmethod
    def forward(g, mod, in_shape, need_dx):
        out = conv_backward_data(g if g.dtype == torch.float32 else g.float(), mod, in_shape, need_dx)
        return out if out is not None else g.new_empty(0)

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, *grads):  # pragma: no cover
        raise NotImplementedError("double backward through the convolution engine is not supported")

    @staticmethod
    def vmap(info, in_dims, g, mod, in_shape, need_dx):
        g = g.movedim(in_dims[0], 0)
        nb, B = g.shape[0], g.shape[1]
        out = _ConvBwdData.apply(g.reshape(nb * B, *g.shape[2:]), mod, (nb * B,) + tuple(in_shape[1:]), need_dx)
        if out.numel() == 0:
            return out, None
        return out.view(nb, B, *out.shape[1:]), 0


class _Conv(torch.autograd.Function):
    @staticmethod
    def forward(x, weight, mod):
        return conv_forward(x, mod)

    @staticmethod
    def setup_context(ctx, inputs, output):
        x, weight, mod = inputs
        ctx.mod = mod
        ctx.in_shape = tuple(x.shape)

    @staticmethod
    def backward(ctx, g):
        gx = _ConvBwdData.apply(g, ctx.mod, ctx.in_shape, bool(ctx.needs_input_grad[0]))
        return (gx if ctx.needs_input_grad[0] else None), None, None


# ------------------------------------------------------------------------------------------ nn.Linear
def linear_forward(x2: torch.Tensor, mod: nn.Linear) -> torch.Tensor:
    """``x2 [rows, d_in] @ W^T`` on the tensor cores (fp16 hi/lo operands); stashes the input rows for the A factor."""
    X = K.pack_cast(x2, KIND_FWD)
    STASH.setdefault(id(mod), {})["P"] = X
    Wk = _CACHE.get(mod, "fwd")                                  # [d_out, d_in], pre-scaled
    out = torch.empty(x2.shape[0], mod.out_features, device=x2.device, dtype=torch.float32)
    K.gemm_nt(X, Wk, out, getattr(Wk, "inv_scale", 1.0), accumulate=False)
    return out


def linear_backward_data(g2: torch.Tensor, mod: nn.Linear, need_dx: bool):
    G = K.pack_cast(g2, KIND)                                    # [rows, d_out] bf16 hi/lo
    STASH.setdefault(id(mod), {})["G"] = G
    if not need_dx:
        return None
    Wt = _CACHE.get(mod, "bwd")                                  # [d_in, d_out]
    out = torch.empty(g2.shape[0], mod.in_features, device=g2.device, dtype=torch.float32)
    K.gemm_nt(G, Wt, out, 1.0, accumulate=False)
    return out


class _LinearBwdData(torch.autograd.Function):
    @staticmethod
    def forward(g, mod, need_dx):
        g = g if g.dtype == torch.float32 else g.float()
        out = linear_backward_data(g.reshape(-1, g.shape[-1]).contiguous(), mod, need_dx)
        return out.view(*g.shape[:-1], mod.in_features) if out is not None else g.new_empty(0)

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, *grads):  # pragma: no cover
        raise NotImplementedError

    @staticmethod
    def vmap(info, in_dims, g, mod, need_dx):
        g2, nb, B = _fold(g, in_dims[0])
        out = _LinearBwdData.apply(g2, mod, need_dx)
        if out.numel() == 0:
            return out, None
        return out.view(nb, B, *out.shape[1:]), 0


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(x, weight, mod):
        out = linear_forward(x.reshape(-1, x.shape[-1]).contiguous(), mod).view(*x.shape[:-1], mod.out_features)
        return out if mod.bias is None else out + mod.bias.detach()

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.mod = inputs[2]

    @staticmethod
    def backward(ctx, g):
        gx = _LinearBwdData.apply(g, ctx.mod, bool(ctx.needs_input_grad[0]))
        return (gx if ctx.needs_input_grad[0] else None), None, None


class _ConvFusedBwd(torch.autograd.Function):
    @staticmethod
    def forward(g, mod, in_shape, need_dx, scale, y, reps):
        out = conv_backward_fused(g if g.dtype == torch.float32 else g.float(), mod, in_shape, need_dx, scale, y, reps)
        return out if out is not None else g.new_empty(0)

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, *grads):  # pragma: no cover
        raise NotImplementedError("double backward through the convolution engine is not supported")

    @staticmethod
    def vmap(info, in_dims, g, mod, in_shape, need_dx, scale, y, reps):
        g = g.movedim(in_dims[0], 0)
        nb, B = g.shape[0], g.shape[1]
        out = _ConvFusedBwd.apply(g.reshape(nb * B, *g.shape[2:]), mod, (nb * B,) + tuple(in_shape[1:]), need_dx, scale, y,
                                  nb * reps)
        if out.numel() == 0:
            return out, None
        return out.view(nb, B, *out.shape[1:]), 0


class _ConvAffine(torch.autograd.Function):
    """``affine(conv(x))`` given the convolution output ``t1`` (already computed by ``_Conv``): forward is the affine map
    alone, the reverse pass runs scale + operand split + backward-data convolution as one fused chain."""

    @staticmethod
    def forward(x, weight, mod, t1, scale, shift):
        return t1 * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)

    @staticmethod
    def setup_context(ctx, inputs, output):
        x, weight, mod, t1, scale, shift = inputs
        ctx.mod, ctx.in_shape = mod, tuple(x.shape)
        ctx.save_for_backward(scale)

    @staticmethod
    def backward(ctx, g):
        need = bool(ctx.needs_input_grad[0])
        gx = _ConvFusedBwd.apply(g, ctx.mod, ctx.in_shape, need, ctx.saved_tensors[0], None, 1)
        return (gx if need else None), None, None, None, None, None


class _ConvAffineRelu(torch.autograd.Function):
    """``relu(affine?(conv(x)))`` given the pre-activation ``t2``; ``scale`` is ``None`` without a BatchNorm."""

    @staticmethod
    def forward(x, weight, mod, t2, scale):
        return torch.relu(t2)

    @staticmethod
    def setup_context(ctx, inputs, output):
        x, weight, mod, t2, scale = inputs
        ctx.mod, ctx.in_shape, ctx.has_scale = mod, tuple(x.shape), scale is not None
        if scale is not None:
            ctx.save_for_backward(output, scale)
        else:
            ctx.save_for_backward(output)

    @staticmethod
    def backward(ctx, g):
        need = bool(ctx.needs_input_grad[0])
        y = ctx.saved_tensors[0]
        scale = ctx.saved_tensors[1] if ctx.has_scale else None
        gx = _ConvFusedBwd.apply(g, ctx.mod, ctx.in_shape, need, scale, y, 1)
        return (gx if need else None), None, None, None, None


def _tag(t: torch.Tensor, info: tuple) -> torch.Tensor:
    t._lpb_tag = info
    return t


def supported(mod: nn.Module) -> bool:
    return (isinstance(mod, nn.Conv2d) and mod.groups == 1 and not isinstance(mod.padding, str)
            and mod.padding_mode == "zeros")


def _fold(g, in_dim):
    """vmap helper: move the column dimension first and fold it into the batch.  Returns (folded, nb, B)."""
    g = g.movedim(in_dim, 0)
    nb, B = g.shape[0], g.shape[1]
    return g.reshape(nb * B, *g.shape[2:]), nb, B


class _AffineBwd(torch.autograd.Function):
    """``g * scale[c]`` -- reverse pass of a frozen BatchNorm2d."""

    @staticmethod
    def forward(g, scale):
        return K.scale_channels(g if g.dtype == torch.float32 else g.float(), scale)

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, *grads):  # pragma: no cover
        raise NotImplementedError

    @staticmethod
    def vmap(info, in_dims, g, scale):
        g2, nb, B = _fold(g, in_dims[0])
        out = _AffineBwd.apply(g2, scale)
        return out.view(nb, B, *out.shape[1:]), 0


class _Affine(torch.autograd.Function):
    @staticmethod
    def forward(x, scale, shift):
        return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.save_for_backward(inputs[1])

    @staticmethod
    def backward(ctx, g):
        return _AffineBwd.apply(g, ctx.saved_tensors[0]), None, None


class _ReluBwd(torch.autograd.Function):
    @staticmethod
    def forward(g, y, reps):
        g = g if g.dtype == torch.float32 else g.float()
        B = y.shape[0]
        dense = y.is_contiguous() or (y.dim() == 4 and y.is_contiguous(memory_format=torch.channels_last))
        if dense and g.shape[0] == reps * B and tuple(g.stride()) == tuple(y.stride()):
            return K.relu_bwd(g, y, reps)
        return (g.reshape(reps, B, *y.shape[1:]) * (y > 0)).reshape(g.shape)

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, *grads):  # pragma: no cover
        raise NotImplementedError

    @staticmethod
    def vmap(info, in_dims, g, y, reps):
        g2, nb, B = _fold(g, in_dims[0])
        out = _ReluBwd.apply(g2, y, nb * reps)
        return out.view(nb, B, *out.shape[1:]), 0


class _Relu(torch.autograd.Function):
    @staticmethod
    def forward(x):
        return torch.relu(x)

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.save_for_backward(output)

    @staticmethod
    def backward(ctx, g):
        return _ReluBwd.apply(g, ctx.saved_tensors[0], 1)


class _MaxPoolBwd(torch.autograd.Function):
    """``chain`` = ``None`` or ``(conv module, BN scale or None, pool input y)`` when the pool's input is the output of a
    fused conv -> BN -> ReLU chain: the un-pooled gradient is then emitted as that chain's packed operand rows
    (``maxpool2d_bwd_pack``) and autograd only carries a storage-less placeholder (see ``conv_backward_fused``)."""

    @staticmethod
    def forward(g, idx, in_shape, k, s, p, chain):
        g = g if g.dtype == torch.float32 else g.float()
        cl = torch.channels_last
        if (chain is not None and g.shape[1] % 8 == 0 and g.is_contiguous(memory_format=cl) and not g.is_contiguous()
                and idx.is_contiguous(memory_format=cl) and chain[2].is_contiguous(memory_format=cl)
                and idx.shape[0] * in_shape[1] * in_shape[2] * in_shape[3] < (1 << 31)):
            cm, scale, y = chain
            G = K.maxpool2d_bwd_pack(g, idx, in_shape, k, s, p, scale, y)
            dummy = g.new_zeros(1).expand(g.shape[0], g.shape[1], in_shape[2], in_shape[3])
            STASH.setdefault(id(cm), {})["G_pre"] = (G, PASS_ID[0], dummy.data_ptr())
            return dummy
        return K.maxpool2d_bwd(g, idx, in_shape, k, s, p)

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, *grads):  # pragma: no cover
        raise NotImplementedError

    @staticmethod
    def vmap(info, in_dims, g, idx, in_shape, k, s, p, chain):
        g2, nb, B = _fold(g, in_dims[0])
        out = _MaxPoolBwd.apply(g2, idx, in_shape, k, s, p, chain)
        return out.view(nb, B, *out.shape[1:]), 0


class _MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(x, k, s, p, chain):
        out, idx = torch.nn.functional.max_pool2d(x, k, s, p, return_indices=True)
        return out, idx

    @staticmethod
    def setup_context(ctx, inputs, output):
        x, k, s, p, chain = inputs
        ctx.save_for_backward(output[1])
        ctx.mark_non_differentiable(output[1])
        ctx.geom = (tuple(x.shape), k, s, p)
        # the chain's ReLU mask is the sign of THIS input: keep it (detached) for the fused reverse map
        ctx.chain = None if chain is None else (chain[0], chain[1], x.detach())

    @staticmethod
    def backward(ctx, g, _gidx):
        in_shape, k, s, p = ctx.geom
        return _MaxPoolBwd.apply(g, ctx.saved_tensors[0], in_shape, k, s, p, ctx.chain), None, None, None, None


def _pool_geom(m: nn.Module):
    """(k, s, p) of a square, undilated, floor-mode ``nn.MaxPool2d`` -- else ``None`` (left to PyTorch)."""
    if not isinstance(m, nn.MaxPool2d) or m.ceil_mode or m.return_indices:
        return None
    def one(v):
        if isinstance(v, (tuple, list)):
            return v[0] if len(set(v)) == 1 else None
        return v
    k, s, p, d = one(m.kernel_size), one(m.stride if m.stride is not None else m.kernel_size), one(m.padding), one(m.dilation)
    if None in (k, s, p, d) or d != 1:
        return None
    return int(k), int(s), int(p)


def _frozen_eval_bn(m: nn.Module) -> bool:
    """BatchNorm2d in eval mode with running statistics and no trainable affine: a fixed per-channel affine map."""
    return (isinstance(m, nn.BatchNorm2d) and not m.training and m.track_running_stats and m.running_mean is not None
            and not any(p.requires_grad for p in m.parameters(recurse=False)))


# The custom element-wise Functions trade ~0.15 ms of host time per call (functorch dispatch of a Python
# autograd.Function) for 2-4x less device time on the folded C x B gradients: worth it once a step is device-bound.
ELEMENTWISE_MIN_BATCH = 1024
ELEMENTWISE_MIN_NUMEL = 1 << 23   # ... or a single activation of >= 8 M elements (wide layers at small batch)


def _device_bound(x: torch.Tensor) -> bool:
    return x.shape[0] >= ELEMENTWISE_MIN_BATCH or x.numel() >= ELEMENTWISE_MIN_NUMEL

_BN_CACHE = weakref.WeakKeyDictionary()   # module -> (tag, (scale, shift)); dies with the module


def _bn_affine(m: nn.BatchNorm2d):
    tag = (m.running_var.data_ptr(), m.running_var._version, m.running_mean.data_ptr(), m.running_mean._version,
           None if m.weight is None else (m.weight.data_ptr(), m.weight._version),
           None if m.bias is None else (m.bias.data_ptr(), m.bias._version), m.running_var.device, m.eps)
    hit = _BN_CACHE.get(m)
    if hit is not None and hit[0] == tag:
        return hit[1]
    out = _bn_affine_compute(m)
    _BN_CACHE[m] = (tag, out)
    return out


def _bn_affine_compute(m: nn.BatchNorm2d):
    invstd = torch.rsqrt(m.running_var + m.eps)
    scale = invstd if m.weight is None else m.weight.detach() * invstd
    shift = -m.running_mean * scale
    if m.bias is not None:
        shift = shift + m.bias.detach()
    return scale.float().contiguous(), shift.float().contiguous()


class patched_convs:
    """Context manager: route the forward (and thereby the reverse pass) of every supported ``nn.Conv2d`` of
    ``model`` through the engine.  ``weight`` is passed to the Function so that the output joins the autograd
    graph even when the input does not require grad (first layer).

    Frozen eval-mode ``BatchNorm2d`` layers are evaluated as the per-channel affine map they are
    (``x * scale + shift``): identical values, but the reverse pass becomes one broadcast multiply instead of
    ``native_batch_norm_backward``, whose vmap rule folds the column dimension into channels with two physical
    copies of the C-times-batched gradient (19 % of a step in profiles/r01_launches_conv_engine_implicit.md)."""

    def __init__(self, model: nn.Module, fuse: bool = False):
        # fuse: chain conv -> frozen BN -> ReLU (each optional after the conv) into one reverse-pass node per
        # convolution (``_ConvAffine`` / ``_ConvAffineRelu``).  The per-layer output gradients are then not available
        # as tensors -- only as the packed rows in ``STASH`` -- so only the KFAC path of the backend asks for it.
        self.fuse = fuse
        self.mods = [m for m in model.modules() if supported(m)]
        self.bns = [m for m in model.modules() if _frozen_eval_bn(m)]
        self.linears = [m for m in model.modules() if type(m) is nn.Linear and m.in_features >= 16 and m.out_features >= 16]
        self.relus = [m for m in model.modules() if type(m) is nn.ReLU]
        self.pools = [(m, _pool_geom(m)) for m in model.modules() if _pool_geom(m) is not None]

    def __enter__(self):
        for m in self.mods:
            def fwd(x, m=m):
                if x.dtype != torch.float32 or not x.is_cuda and not _ALLOW_CPU:
                    return nn.Conv2d.forward(m, x)
                out = _Conv.apply(x, m.weight, m)
                if self.fuse and _device_bound(x):
                    _tag(out, ("conv", m, x))
                return out
            m.forward = fwd
        def usable(x):
            return x.dtype == torch.float32 and (x.is_cuda or _ALLOW_CPU)

        for m in self.bns:
            def bn_fwd(x, m=m):
                if x.dim() != 4 or not usable(x):
                    return nn.BatchNorm2d.forward(m, x)
                scale, shift = _bn_affine(m)
                if not _device_bound(x):
                    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
                tag = getattr(x, "_lpb_tag", None)
                if self.fuse and tag is not None and tag[0] == "conv":
                    _, cm, xin = tag
                    STASH.setdefault(id(cm), {})["fused"] = True
                    out = _ConvAffine.apply(xin, cm.weight, cm, x.detach(), scale, shift)
                    return _tag(out, ("conv_affine", cm, xin, scale))
                return _Affine.apply(x, scale, shift)
            m.forward = bn_fwd
        for m in self.linears:
            def lin_fwd(x, m=m):
                return _Linear.apply(x, m.weight, m) if usable(x) else nn.Linear.forward(m, x)
            m.forward = lin_fwd
        for m in self.relus:
            def relu_fwd(x, m=m):
                if not (usable(x) and _device_bound(x)):
                    return nn.ReLU.forward(m, x)
                tag = getattr(x, "_lpb_tag", None)
                if self.fuse and tag is not None and tag[0] in ("conv", "conv_affine") and x.dim() == 4:
                    cm, xin = tag[1], tag[2]
                    STASH.setdefault(id(cm), {})["fused"] = True
                    sc = tag[3] if tag[0] == "conv_affine" else None
                    return _tag(_ConvAffineRelu.apply(xin, cm.weight, cm, x.detach(), sc), ("chain_out", cm, sc))
                return _Relu.apply(x)
            m.forward = relu_fwd
        for m, geom in self.pools:
            def pool_fwd(x, m=m, geom=geom):
                if x.dim() != 4 or not usable(x) or not _device_bound(x) or x.shape[2] * x.shape[3] > 1024:
                    return nn.MaxPool2d.forward(m, x)
                tag = getattr(x, "_lpb_tag", None)
                chain = (tag[1], tag[2]) if (self.fuse and FUSE_POOL and tag is not None and tag[0] == "chain_out") else None
                return _MaxPool.apply(x, *geom, chain)[0]
            m.forward = pool_fwd
        return self

    def __exit__(self, *exc):
        for m in self.mods + self.bns + self.linears + self.relus + [p[0] for p in self.pools]:
            m.__dict__.pop("forward", None)
        return False


_ALLOW_CPU = False  # tests flip this together with the CPU kernel emulation
