"""``B200GGN`` / ``B200EF`` -- curvature backends for the Laplace library whose per-batch hot path
runs on hand-written sm_100a kernels (``liblaplace_b200.so``).

Plug-in surface = reference ``laplace.curvature.CurvatureInterface`` (curvature/curvature.py:12-291):
``jacobians``, ``last_layer_jacobians``, ``gradients``, ``full``, ``kron``, ``diag``.  Semantics of every
method follow the reference's default backend, the curvlinops adapter (curvature/curvlinops.py:46-188)
for ``kron`` and the in-tree GGN/EF (curvature/curvature.py:375-505) for ``full`` / ``diag``.

How a batch is processed:

1. one forward pass of the model with hooks on every ``nn.Linear`` / ``nn.Conv2d`` that owns a trainable
   parameter captures the layer input ``a`` and the layer output tensor;
2. one *batched* reverse pass (``autograd.grad(..., is_grads_batched=True)``) propagates the ``ncols``
   columns that define the curvature (C columns of the loss-Hessian square root for the GGN, the
   loss gradient for the EF, sampled gradients for the MC Fisher) down to every captured layer
   output -- this is the network's own forward/backward (cuDNN/cuBLAS through autograd) and is *below*
   the plug-in boundary, exactly as in the reference;
3. everything above it -- packing ``a`` / ``g`` into K-major bf16 (hi/lo) or fp32 staging, the
   ``X^T X`` factor contractions, diagonal reductions, Jacobian materialisation, the last-layer
   structured GGN -- runs in the native kernels.
"""
from __future__ import annotations

import math
from collections.abc import MutableMapping
from typing import Any

import torch
from torch import nn

from . import kernels as K
from .interface import CurvatureInterface, EFInterface, GGNInterface
from .matrix import B200Kron, JacobianFactors

SUPPORTED = (nn.Linear, nn.Conv2d)
PRECISIONS = ("auto", "fp32", "bf16", "bf16x3")
# precision="auto": input (A) factors use ONE fp16 product instead of three when they are summed over at least
# max(A_SINGLE_PRODUCT_MIN_ROWS, A_SINGLE_PRODUCT_ROWS_PER_DIM * d_in) sample rows.  fp16 rounds every activation by at most
# 2^-12 relative, independently across rows; for a factor over K rows of d features the relative Frobenius error is at most
# ~ 3.4e-4 * sqrt(d / K) (white inputs: measured 2.1e-5 on the 147-feature stem at K = 16 384, predicted 3.2e-5) and an order
# of magnitude less on post-ReLU activations, whose factors are dominated by the mean (measured <= 1e-5 on every such layer
# of ResNet-18).  128 rows per feature bounds the white-input case by 3e-5, inside the 1e-4 gate next to the 2.5e-5 of the
# reverse pass.  Output-gradient (B) factors and the reverse pass itself keep three products: a two-product backward-data
# convolution (gradient rows rounded to bf16, exact weights: conv_engine.LEAN_BWD_MIN_ROWS) was measured at 1.1e-4 .. 1.8e-4
# on the B factors -- the rounding of a gradient element does not average out inside its own dot product -- and stays off.
A_SINGLE_PRODUCT_MIN_ROWS = 16384
A_SINGLE_PRODUCT_ROWS_PER_DIM = 128
LEAN_BACKWARD_MIN_ROWS = 0     # > 0: two-product backward-data convolutions in kron() from this many gradient rows (off)


class _Layer:
    __slots__ = ("name", "mod", "has_w", "has_b", "d_in", "d_out", "is_conv")

    def __init__(self, name, mod, has_w, has_b):
        self.name, self.mod, self.has_w, self.has_b = name, mod, has_w, has_b
        self.is_conv = isinstance(mod, nn.Conv2d)
        if self.is_conv:
            if mod.groups != 1:
                raise ValueError(f"{name}: grouped convolutions are not supported by the B200 backend")
            if isinstance(mod.padding, str) or mod.padding_mode != "zeros":
                raise ValueError(f"{name}: only zero padding given as ints is supported")
            self.d_in = mod.in_channels * mod.kernel_size[0] * mod.kernel_size[1]
            self.d_out = mod.out_channels
        else:
            self.d_in, self.d_out = mod.in_features, mod.out_features


class _Fp32Model:
    """Context manager: disable TF32 in cuDNN convolutions and cuBLAS matmuls for the enclosed model passes."""

    def __enter__(self):
        self._c, self._m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self._c, self._m
        return False


def _inert_hook(mod, inp, out):
    """What a pickled ``_CaptureHook`` comes back as: its backend did not travel with it (the unpickled backend registers its
    own hooks)."""
    return None


class _CaptureHook:
    """Forward hook of one Linear / Conv2d layer: records ``(input, output)`` while its backend runs a captured forward pass.
    An object rather than a closure so that a model carrying it can be pickled (see ``_B200Mixin.__getstate__``)."""

    def __init__(self, backend, name):
        import weakref

        self.ref = weakref.ref(backend)     # the module must not keep the backend (and its captured graphs / buffers) alive
        self.name = name

    def __call__(self, mod, inp, out):
        be, name = self.ref(), self.name
        if be is not None and be._capturing:
            if name in be._outs:
                # the reference's hook-based KFAC sums the contributions of every call; this backend captures one
                # (input, output) pair per module, so say so instead of silently keeping the last call only
                raise ValueError(f"B200 backend: module {name!r} is called more than once per forward pass (weight "
                                 "sharing across calls); give every call its own module")
            be._acts[name] = inp[0].detach()
            be._outs[name] = out
            be._out_state[name] = (out._version, out.grad_fn)

    def __reduce__(self):
        return (_inert_hook_factory, ())


def _inert_hook_factory():
    return _inert_hook


class _B200Mixin:
    """Shared machinery of the GGN and EF flavours."""

    def _b200_init(self, precision: str = "auto", batched_backward: bool = True, model_tf32: bool = False,
                   conv_engine: bool = True, fuse_elementwise: bool = True, cuda_graph: bool = False):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        self.precision = precision
        self.batched_backward = batched_backward
        self.model_tf32 = model_tf32
        self.conv_engine = conv_engine and not model_tf32
        # KFAC path only: reverse pass of conv -> frozen BN -> ReLU chains as one node per convolution (conv_engine.py)
        self.fuse_elementwise = fuse_elementwise
        self._fused = False
        # kron(): replay a captured CUDA graph of the whole step once the same (shapes, N, kwargs, parameter versions) has
        # run eagerly twice (opt-in: a model whose forward branches on data would be frozen on the captured branch)
        self.cuda_graph = bool(cuda_graph)
        self._graphs: dict = {}
        # KFAC path: input (A) factors on a side stream, concurrent with the reverse pass
        import os

        self.overlap_factors = os.environ.get("LPB_NO_OVERLAP") != "1"
        self._side = None
        self._layers: list[_Layer] | None = None
        self._unsupported: list[str] = []
        self._hooks = []
        self._capturing = False
        self._acts: dict[str, torch.Tensor] = {}
        self._outs: dict[str, torch.Tensor] = {}
        self._out_state: dict[str, tuple] = {}

    # ------------------------------------------------------------------ layer plan
    def _plan(self) -> list[_Layer]:
        """Layers in ``named_modules()`` order that own parameters of ``self.params`` (the walk order of
        ``CurvlinopsInterface._get_kron_factors``, curvature/curvlinops.py:55-75)."""
        if self._layers is not None:
            return self._layers
        ids = {id(p) for p in self.params}
        layers, unsupported, covered = [], [], set()
        for name, mod in self.model.named_modules():
            own = [p for p in mod.parameters(recurse=False) if id(p) in ids]
            if not own:
                continue
            if isinstance(mod, SUPPORTED):
                has_w = id(mod.weight) in ids
                has_b = mod.bias is not None and id(mod.bias) in ids
                layers.append(_Layer(name, mod, has_w, has_b))
                covered.update(id(p) for p in own)
            else:
                unsupported.append(f"{name} ({type(mod).__name__})")
        order = {id(p): i for i, p in enumerate(self.params)}
        pos = []
        for L in layers:
            if L.has_w:
                pos.append(order[id(L.mod.weight)])
            if L.has_b:
                pos.append(order[id(L.mod.bias)])
        if pos != sorted(pos) or len(set(pos)) != len(pos):
            raise ValueError("module order and parameters() order disagree (shared or re-used parameters?)")
        self._layers, self._unsupported = layers, unsupported
        for L in layers:
            self._hooks.append(L.mod.register_forward_hook(self._make_hook(L.name)))
        return layers

    def _require_supported(self, what: str):
        self._plan()
        if self._unsupported:
            raise ValueError(
                f"B200 backend: {what} supports trainable parameters in nn.Linear / nn.Conv2d only; found "
                + ", ".join(self._unsupported) + ". Freeze them (requires_grad=False) as for the reference KFAC path."
            )

    def _make_hook(self, name):
        return _CaptureHook(self, name)

    # ``torch.save(la, path)`` (reference tests/test_serialization.py:295-336) pickles the posterior together with its backend
    # and the model -- forward hooks included.  Streams, captured graphs, hook handles and captured tensors are per-process
    # state: they are dropped here and rebuilt lazily (``_plan`` registers fresh hooks on the unpickled model).
    _TRANSIENT = {"_graphs": dict, "_side": lambda: None, "_layers": lambda: None, "_hooks": list, "_acts": dict,
                  "_outs": dict, "_out_state": dict, "_unsupported": list, "_jac_cache": lambda: None}

    def __getstate__(self):
        state = dict(self.__dict__)
        for k, fresh in self._TRANSIENT.items():
            if k in state:
                state[k] = fresh()
        state["_capturing"] = False
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)

    def __del__(self):
        for h in getattr(self, "_hooks", ()):
            try:
                h.remove()
            except Exception:  # noqa: BLE001 -- interpreter shutdown
                pass

    # in-place ops that leave d(result)/d(tensor) = identity: the gradient w.r.t. the modified tensor IS the gradient
    # w.r.t. the layer output (e.g. torchvision's ``out += identity``)
    _INPLACE_IDENTITY = ("AddBackward", "SubBackward")

    def _check_outputs_intact(self):
        """``autograd.grad(f, out)`` differentiates w.r.t. the CURRENT version of ``out``: if an in-place activation ran
        on the captured layer output (``F.relu_``, ``x.clamp_()`` ...) the result would silently lack that activation's
        mask.  Module-level ``inplace=True`` flags are switched off while capturing (``_no_inplace``); anything that
        still modified a captured output in place is reported."""
        for name, out in self._outs.items():
            ver, fn = self._out_state.get(name, (out._version, out.grad_fn))
            if out._version == ver:
                continue
            new = type(out.grad_fn).__name__ if out.grad_fn is not None else "None"
            if out.grad_fn is not fn and new.startswith(self._INPLACE_IDENTITY):
                continue
            raise RuntimeError(f"B200 backend: the output of layer {name!r} was modified in place after the layer ran "
                               f"(autograd node {new}); its gradient would miss that operation. Use the out-of-place form.")

    class _no_inplace:
        """Switch off ``inplace=True`` on activation / dropout modules for the duration of a captured forward pass."""

        def __init__(self, model):
            self.mods = [m for m in model.modules() if getattr(m, "inplace", False) is True]

        def __enter__(self):
            for m in self.mods:
                m.inplace = False

        def __exit__(self, *exc):
            for m in self.mods:
                m.inplace = True
            return False

    # ------------------------------------------------------------------ forward / backward
    def _device_check(self, t: torch.Tensor):
        if not t.is_cuda:
            raise RuntimeError("the B200 curvature backend runs on CUDA only (model/input on %s); there is no "
                               "CPU fallback" % t.device)

    def _forward(self, x, fuse: bool = False):
        self._plan()
        self._acts, self._outs, self._out_state = {}, {}, {}
        self._fused = bool(fuse and self.conv_engine)
        if self.conv_engine:
            from . import conv_engine

            conv_engine.STASH.clear()
        self._capturing = True
        try:
            with torch.enable_grad(), self._model_numerics(), self._no_inplace(self.model), self._conv_patch(self._fused):
                f = self.model(x)
        finally:
            self._capturing = False
        self._check_outputs_intact()
        if f.ndim != 2:
            raise ValueError(f"the B200 backend supports (batch, outputs) model outputs, got shape {tuple(f.shape)}")
        self._device_check(f)
        self._n_outputs = int(f.shape[1])
        if self.conv_engine and f.is_cuda:
            # the engine's forward uses fp16 hi/lo operands: an activation beyond +-65504 would surface here
            torch._assert_async(torch.isfinite(f).all())
        return f

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream()
        return self._side

    def _conv_patch(self, fuse: bool = False):
        """fp32-accurate convolution passes on the tensor cores (laplace_b200/conv_engine.py) instead of cuDNN's
        fp32 fallback kernels; ``conv_engine=False`` keeps the model's own convolution implementation."""
        import contextlib

        if not self.conv_engine:
            return contextlib.nullcontext()
        from .conv_engine import patched_convs

        return patched_convs(self.model, fuse=fuse)

    def _model_numerics(self):
        """The network's own forward/reverse passes run in true fp32 unless ``model_tf32=True``: PyTorch's default
        lets cuDNN use TF32 for convolutions, which perturbs Jacobians by ~1e-3 -- far above the 1e-4 parity gate."""
        import contextlib

        if self.model_tf32:
            return contextlib.nullcontext()
        return _Fp32Model()

    def _backward(self, f: torch.Tensor, cols: torch.Tensor) -> list[torch.Tensor]:
        with self._model_numerics():
            return self._backward_impl(f, cols)

    def _backward_impl(self, f: torch.Tensor, cols: torch.Tensor) -> list[torch.Tensor]:
        """Gradients of ``sum_n <cols[j, n], f[n]>`` w.r.t. every captured layer output, for all ``j`` at once.
        Returns one fp32 tensor ``[ncols, M, ...]`` per planned layer."""
        outs = [self._outs[L.name] for L in self._layers]
        n_layers = len(outs)
        bump = lambda: None
        if self.conv_engine:
            from . import conv_engine

            def bump():
                conv_engine.PASS_ID[0] += 1

            if self._fused:
                # fused chains hang off the layer INPUT and weight, not off the captured layer output: ask for the
                # weights of fused layers too so that their reverse node runs (and packs its gradient rows) even when
                # nothing upstream needs an input gradient (first layer).  The weight "gradient" itself is None.
                outs = outs + [L.mod.weight for L in self._layers
                               if conv_engine.STASH.get(id(L.mod), {}).get("fused") and L.mod.weight.requires_grad]
        cols = cols.to(f.dtype)
        grads = None
        if self.batched_backward and cols.shape[0] > 1:
            try:
                if self.conv_engine:
                    # functorch vmap over autograd.grad: custom Functions fold the column dim into the batch
                    missing = []

                    def one(col):
                        gs = torch.autograd.grad(f, outs, grad_outputs=col, retain_graph=True, allow_unused=True)
                        missing[:] = [g is None for g in gs]
                        return tuple(col.new_zeros(()) if g is None else g for g in gs)

                    bump()
                    grads = list(torch.func.vmap(one)(cols))
                    grads = [None if miss else g for g, miss in zip(grads, missing)]
                else:
                    grads = torch.autograd.grad(f, outs, grad_outputs=cols, is_grads_batched=True, retain_graph=True,
                                                allow_unused=True)
                self.last_backward_mode = "batched"
            except (RuntimeError, NotImplementedError) as e:
                if self.conv_engine and isinstance(e, conv_engine.FusionConflict):
                    raise
                grads = None  # an op without a batching rule: fall back to one reverse pass per column
                self.last_backward_mode = f"loop ({type(e).__name__}: {str(e)[:120]})"
        if grads is None:
            per = []
            for j in range(cols.shape[0]):
                bump()
                per.append(torch.autograd.grad(f, outs, grad_outputs=cols[j], retain_graph=j + 1 < cols.shape[0],
                                               allow_unused=True))
            grads = [None if per[0][i] is None else torch.stack([p[i] for p in per]) for i in range(len(outs))]
        res = []
        for L, g, o in zip(self._layers, grads[:n_layers], outs[:n_layers]):
            if g is None and self._fused:
                res.append(None)   # fused chain: the gradient exists only as packed rows in conv_engine.STASH
                continue
            if g is None:
                g = torch.zeros((cols.shape[0],) + tuple(o.shape), device=o.device, dtype=torch.float32)
            res.append(g.detach().float())   # layout fixed lazily by the consumers (no copy when unused)
        self._outs = {}
        return res

    # ------------------------------------------------------------------ columns that define the curvature
    def _hessian_sqrt_cols(self, f: torch.Tensor) -> torch.Tensor:
        """GGN: columns of ``S_n`` with ``S_n S_n^T`` = functional Hessian of the reference
        (``diag(p) - p p^T`` / identity, curvature/curvature.py:366-373).  Shape ``[C, M, C]``."""
        M, C = f.shape
        if self.likelihood == "regression":
            return torch.eye(C, device=f.device, dtype=f.dtype).unsqueeze(1).expand(C, M, C).contiguous()
        p = torch.softmax(f, dim=-1)
        sp = p.sqrt()
        S = torch.diag_embed(sp) - p.unsqueeze(2) * sp.unsqueeze(1)   # [M, C(row), C(col)]
        return S.permute(2, 0, 1).contiguous()

    def _mc_cols(self, f: torch.Tensor, n_samples: int) -> torch.Tensor:
        """MC Fisher (``_get_mc_functional_fisher``, curvature/curvature.py:341-364): ``n_samples`` sampled
        functional gradients scaled by ``1/sqrt(n_samples)``.  Shape ``[n_samples, M, C]``."""
        cols = []
        for _ in range(n_samples):
            if self.likelihood == "regression":
                cols.append(-torch.randn_like(f))
            else:
                p = torch.softmax(f, dim=-1)
                ys = torch.multinomial(p, 1).squeeze(1)
                cols.append(p - torch.nn.functional.one_hot(ys, f.shape[-1]).to(f.dtype))
        return torch.stack(cols) / math.sqrt(n_samples)

    def _loss_grad_cols(self, f: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """d(sum-reduced torch loss)/df: ``p - onehot(y)`` (CE) / ``2 (f - y)`` (MSE).  Shape ``[1, M, C]``."""
        if self.likelihood == "regression":
            return (2.0 * (f - y)).unsqueeze(0)
        p = torch.softmax(f, dim=-1)
        return (p - torch.nn.functional.one_hot(y, f.shape[-1]).to(f.dtype)).unsqueeze(0)

    def _diag_conv_tc(self, L, a, g, M, ncols, out, scale) -> bool:
        """Diagonal of a stride-1 convolution weight on the tensor cores (``K.diag_conv_sq``): per-sample weight
        gradients formed in TMEM from the engine's gradient rows and the NHWC input rows, squared and summed --
        instead of the fp32 SIMT contraction over a 9x-inflated im2col.  Returns False when the layer does not qualify."""
        if not (self.conv_engine and L.is_conv and self.precision in ("auto", "bf16x3") and a.dim() == 4 and g.dim() == 5):
            return False
        from . import conv_engine

        mod = L.mod
        H, W = a.shape[2], a.shape[3]
        if not (conv_engine.implicit_ok(mod, H, W) and tuple(g.shape[3:]) == (H, W)
                and K.diag_conv_ok(mod.in_channels, H, W, *mod.kernel_size)):
            return False
        G = conv_engine.STASH.get(id(mod), {}).get("G")
        if (G is None or G.rows != ncols * M * H * W or G.kind != K.BF16X3
                or not (ncols == 1 or getattr(self, "last_backward_mode", "") == "batched")):
            g4 = g.reshape(ncols * M, *g.shape[2:])
            G = conv_engine.nhwc_rows(g4 if g4.dtype == torch.float32 else g4.float(), K.BF16X3)
        X = conv_engine.nhwc_rows(a, K.BF16X3)
        K.diag_conv_sq(G, X, M, H, W, mod, out, alpha=scale)
        return True

    # ------------------------------------------------------------------ precision policy
    def _kind(self, d: int, k: int) -> int:
        """Operand format of one contraction with output dim ``d`` and reduction length ``k``."""
        if self.precision == "fp32":
            return K.F32
        if self.precision == "bf16":
            return K.BF16
        if self.precision == "bf16x3":
            return K.BF16X3
        return K.BF16X3 if (d >= 64 and k >= 256) else K.F32

    # ------------------------------------------------------------------ per-layer operand packing
    def _pack_act(self, L: _Layer, a: torch.Tensor, kind: int, reduce: bool = False):
        """K-major layer-input rows ``[d_in, M*T]`` (expand) / ``[d_in, M]`` (reduce).  Returns (packed, T)."""
        a = a.float() if a.dtype != torch.float32 else a
        if L.is_conv:
            return K.pack_conv(a, L.mod, kind, reduce_mean=reduce)
        M = a.shape[0]
        rows = a.reshape(M, -1, a.shape[-1])
        T = rows.shape[1]
        if reduce and T > 1:
            rows = rows.mean(1, keepdim=True)
        rows = rows.reshape(-1, a.shape[-1]).contiguous()
        return K.pack_rows(rows, kind), T

    def _pack_grad(self, L: _Layer, g: torch.Tensor, kind: int, reduce: bool = False):
        """K-major output-gradient rows ``[d_out, ncols*M*T]`` (column-major over ``(col, n, t)``)."""
        if L.is_conv:
            nc, M, Co, OH, OW = g.shape
            if not reduce and g.stride(2) == 1 and g.permute(0, 1, 3, 4, 2).is_contiguous():
                # channels_last gradients (convolution engine): rows [(col,n,h,w), Co] -> tiled transpose
                return K.pack_rows(g.permute(0, 1, 3, 4, 2).reshape(nc * M * OH * OW, Co), kind)
            return K.pack_nchw(g.reshape(nc * M, Co, OH * OW).contiguous(), kind, reduce_sum=reduce)
        nc, M = g.shape[:2]
        rows = g.reshape(nc * M, -1, g.shape[-1])
        if reduce and rows.shape[1] > 1:
            rows = rows.sum(1, keepdim=True)
        return K.pack_rows(rows.reshape(-1, g.shape[-1]).contiguous(), kind)

    # ------------------------------------------------------------------ KFAC through a captured CUDA graph
    GRAPH_WARMUP_CALLS = 2

    def _state_key(self):
        """Versions of everything the captured step reads through cached host-side state (packed weights, folded BN
        affines): an optimiser step between two ``kron()`` calls (``marglik_training``) must trigger a fresh capture."""
        return tuple((t.data_ptr(), t._version) for t in list(self.model.parameters()) + list(self.model.buffers()))

    def _kron_graphed(self, x, y, N, cols_fn, weight, kfac_approx):
        """One ``kron()`` step at B <= 1024 is bound by the host (~1500 kernel launches plus the functorch dispatch of the
        custom autograd Functions: 13 ms at B = 512 against 4 ms of device time).  The step has static shapes and no host
        synchronisation, so after ``GRAPH_WARMUP_CALLS`` eager calls with the same key it is captured once (forward, column-
        batched reverse pass, packs, SYRKs, side stream included) and replayed: inputs are copied into the graph's static
        buffers, the returned loss / ``B200Kron`` ARE the graph's static outputs -- consume them (``la.H += H_batch``,
        baselaplace.py:985) before the next call.  Any failure to capture falls back to eager for good (and restores torch's
        CUDA generator, which an aborted capture leaves in capture mode).  Known case (r02, tools/gpu_graph_diag.py): a second,
        eager, fusing backend on the SAME model instance called in strict alternation with this one invalidates the capture
        (cudaErrorStreamCaptureInvalidated; cause not identified) -- any other interleaving, separate model instances, or
        one backend per model (what ``Laplace(...)`` does) capture fine."""
        if not (self.cuda_graph and torch.is_tensor(x) and x.is_cuda and torch.is_tensor(y)):
            return self._kron_impl(x, y, N, cols_fn, weight, kfac_approx)
        key = (tuple(x.shape), x.dtype, tuple(y.shape), y.dtype, float(N), kfac_approx, float(weight), self.precision,
               self.fuse_elementwise, self._state_key())
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) > 8:
                self._graphs.clear()       # parameter versions moved on (training): drop stale captures
            ent = self._graphs[key] = {"calls": 0, "graph": None}
        if ent["graph"] is False:
            return self._kron_impl(x, y, N, cols_fn, weight, kfac_approx)
        if ent["graph"] is None:
            ent["calls"] += 1
            if ent["calls"] <= self.GRAPH_WARMUP_CALLS:
                return self._kron_impl(x, y, N, cols_fn, weight, kfac_approx)
            try:
                from .matrix import wait_prewarm

                wait_prewarm(x.device)     # library start-up on another thread (cudaMalloc ...) would invalidate the capture
                if self.conv_engine:
                    from . import conv_engine as _ceq

                    _ceq.STASH.clear()     # operands another backend's step left behind are released OUTSIDE the capture
                import gc

                gc.collect()
                rng_state = torch.cuda.get_rng_state(x.device)
                sx, sy = torch.empty_like(x), torch.empty_like(y.to(x.device))
                sx.copy_(x), sy.copy_(y)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                n0 = K.LAUNCHES
                # thread_local: loader / prefetch threads of the host program may keep calling CUDA while we capture
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    out = self._kron_impl(sx, sy, N, cols_fn, weight, kfac_approx)
                ent.update(graph=g, x=sx, y=sy, out=out, launches=K.LAUNCHES - n0)
                K._bump(-ent["launches"])       # capturing launched nothing; every replay launches them all
            except Exception as e:  # noqa: BLE001 -- capture is an optimisation; say why it is off and carry on eagerly
                import warnings

                torch.cuda.synchronize()
                ent["graph"] = False
                try:    # an aborted capture leaves torch's CUDA generator in capture mode ("Offset increment outside graph
                    torch.cuda.manual_seed(0)                      # capture"): re-seeding clears it, then the state is restored
                    torch.cuda.set_rng_state(rng_state, x.device)
                except Exception:  # noqa: BLE001
                    pass
                warnings.warn(f"laplace_b200: CUDA-graph capture of kron() failed ({type(e).__name__}: {str(e)[:200]}); "
                              "running eagerly")
                return self._kron_impl(x, y, N, cols_fn, weight, kfac_approx)
        ent["x"].copy_(x, non_blocking=True)
        ent["y"].copy_(y, non_blocking=True)
        ent["graph"].replay()
        K._bump(ent["launches"])                # the graph's kernel nodes (bench.py reports launches of native kernels)
        return ent["out"]

    # ------------------------------------------------------------------ KFAC
    def _kron_impl(self, x, y, N, cols_fn, weight: float, kfac_approx: str = "expand"):
        if kfac_approx not in ("expand", "reduce"):
            raise ValueError(f"kfac_approx must be 'expand' or 'reduce', got {kfac_approx!r}")
        self._require_supported("kron()")
        reduce = kfac_approx == "reduce"
        fuse = (self.conv_engine and self.fuse_elementwise and not reduce and self.precision in ("auto", "bf16x3")
                and self.batched_backward)
        f = self._forward(x, fuse=fuse)
        fd = f.detach()
        M = fd.shape[0]
        y = y.to(fd.device)
        loss = self.factor * self.lossfunc(fd, y)
        cols = cols_fn(fd, y)
        acts = self._acts
        out_t = {L.name: (math.prod(self._outs[L.name].shape[2:]) if L.is_conv else 1) for L in self._layers}
        dims = []
        for L in self._layers:
            if L.has_w:
                dims.append([L.d_out, L.d_in])
            if L.has_b:
                dims.append([L.d_out])
        kron = B200Kron.zeros(dims, fd.device, torch.float32)
        if fd.is_cuda:
            from .matrix import prewarm_eigensolver

            prewarm_eigensolver(fd.device)     # library start-up overlaps the data pass (once per process)
        sq = math.sqrt(self.factor)
        slots, idx = {}, 0   # layer name -> index of its first Kron block
        for L in self._layers:
            slots[L.name] = idx
            idx += (1 if L.has_w else 0) + (1 if L.has_b else 0)
        fwd_stash = {}
        if self.conv_engine and not reduce and self.precision in ("auto", "bf16x3"):
            from . import conv_engine

            fwd_stash = conv_engine.STASH   # operands the forward packed (patch rows "P", NHWC input rows "X")

        def input_factors():
            """A factors: they need the layer inputs only, so they are issued right after the forward -- on a side
            stream, where the tensor-bound SYRKs overlap the HBM-bound stretches of the reverse pass."""
            for L in self._layers:
                if not L.has_w:
                    continue
                a = acts[L.name]
                Af = kron.kfacs[slots[L.name]][1]
                rows = fwd_stash.get(id(L.mod), {})
                Prows, Xs = rows.get("P"), rows.get("X")

                def lean(P):   # see A_SINGLE_PRODUCT_MIN_ROWS (fp16 operands only: a bf16 half keeps 8 bits, not 11)
                    ok = (self.precision == "auto" and P.kind == K.F16X3
                          and P.rows >= max(A_SINGLE_PRODUCT_MIN_ROWS, A_SINGLE_PRODUCT_ROWS_PER_DIM * L.d_in))
                    return K.hi_only(P) if ok else P

                if (Xs is not None and Xs[1] == M and K.conv_patches_ok(Xs[0].K, Xs[2], Xs[3], *L.mod.kernel_size)):
                    # implicit-path convolution: the A factor straight from the NHWC input rows the forward packed
                    X, _, Hh, Ww = Xs
                    K.syrk_conv_patches(lean(X), M, Hh, Ww, L.mod, Af, alpha=sq / (N * Hh * Ww))
                    continue
                if Prows is None and L.is_conv and len(fwd_stash) > 0:
                    # build the patch rows once (the row-major pack is ~2x cheaper than the transposing K-major one)
                    af = a.float() if a.dtype != torch.float32 else a
                    Prows = K.pack_conv_rows(af, L.mod, K.F16X3)   # the engine's forward operand format
                if Prows is not None and Prows.rows % M != 0:
                    Prows = None
                if Prows is not None:
                    T = Prows.rows // M
                    Pl = lean(Prows)
                    K.gemm_tn(Pl, Pl, Af, alpha=sq / (N * T), accumulate=True, symmetric=True)
                else:
                    k_rows = M * out_t[L.name] if L.is_conv else (a.numel() // a.shape[-1])
                    ak = self._kind(L.d_in, M if reduce else k_rows)
                    A, T = self._pack_act(L, a, ak, reduce)
                    Teff = 1 if reduce else T
                    # A = factor^(1/2) * (M/N) * 1/(M*T) * sum a a^T   (curvlinops.py:46-53, matrix.py:116-118)
                    K.gemm_nt(A, A, Af, alpha=sq / (N * Teff), accumulate=True, symmetric=True)

        side = main = None
        if fd.is_cuda and self.overlap_factors:
            main = torch.cuda.current_stream()
            side = self._side_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                input_factors()
        else:
            input_factors()

        def join():
            if side is not None:
                main.wait_stream(side)

        if self.conv_engine and self.precision == "auto" and LEAN_BACKWARD_MIN_ROWS > 0:
            from . import conv_engine as _ce0

            _ce0.LEAN_BWD_MIN_ROWS = LEAN_BACKWARD_MIN_ROWS     # opt-in experiment, reset in the ``finally`` below
        try:
            try:
                grads = self._backward(f, cols)
            except RuntimeError as e:
                from . import conv_engine as _ce

                if not (self._fused and isinstance(e, _ce.FusionConflict)):
                    raise
                # a fused intermediate has a second consumer in this model: keep the chains unfused from now on
                self.fuse_elementwise = False
                join()   # the re-run forward clears the operands the side stream may still be reading
                f = self._forward(x, fuse=False)
                grads = self._backward(f, cols)
            if self._fused and not (cols.shape[0] == 1 or getattr(self, "last_backward_mode", "") == "batched"):
                # the column-batched reverse pass was not available: the fused chains keep gradients only as packed
                # rows of ONE pass, so redo the passes unfused (one reverse pass per column, gradients as tensors)
                join()
                f = self._forward(x, fuse=False)
                grads = self._backward(f, cols)
        finally:
            join()
            if self.conv_engine:
                from . import conv_engine as _ce1

                _ce1.LEAN_BWD_MIN_ROWS = 0
        self._acts = {}
        stash = {}
        if (self.conv_engine and not reduce and self.precision in ("auto", "bf16x3")
                and (cols.shape[0] == 1 or getattr(self, "last_backward_mode", "") == "batched")):
            from . import conv_engine

            stash = conv_engine.STASH   # gradient rows "G" the engine packed during the reverse pass
        for L, g in zip(self._layers, grads):
            idx = slots[L.name]
            rows = stash.get(id(L.mod), {})
            Grows = rows.get("G")
            if g is None:
                if Grows is None:
                    raise RuntimeError(f"{L.name}: fused reverse chain left no gradient rows (internal error)")
            elif Grows is not None and Grows.rows != g.numel() // L.d_out:
                Grows = None
            if Grows is None and L.is_conv and stash is not None and len(stash) > 0 and g.dim() == 5:
                # a convolution whose reverse node never ran (the first layer: nothing upstream needs its input
                # gradient): pack the output-gradient rows here, channels-last, for the same MN-major SYRK
                g4 = g.reshape(g.shape[0] * g.shape[1], *g.shape[2:])
                Grows = conv_engine.nhwc_rows(g4 if g4.dtype == torch.float32 else g4.float(), K.BF16X3)
            if Grows is None:
                gk = self._kind(L.d_out, g.numel() // L.d_out)
                G = self._pack_grad(L, g, gk, reduce)

            def syrk_B(out, alpha):
                if Grows is not None:   # MN-major SYRK straight on the engine's gradient rows
                    K.gemm_tn(Grows, Grows, out, alpha=alpha, accumulate=True, symmetric=True)
                else:
                    K.gemm_nt(G, G, out, alpha=alpha, accumulate=True, symmetric=True)

            if L.has_w:
                Bf = kron.kfacs[idx][0]
                syrk_B(Bf, sq * weight)
                if L.has_b:
                    kron.kfacs[idx + 1][0].copy_(Bf).mul_(sq)  # bias block: factor * B   (len(F) == 1)
            elif L.has_b:
                syrk_B(kron.kfacs[idx][0], self.factor * weight)
        dtype = next(self.model.parameters()).dtype
        if dtype != torch.float32:
            kron = B200Kron([[H.to(dtype) for H in F] for F in kron.kfacs])
        return loss.detach(), kron

    # ------------------------------------------------------------------ Jacobian rows (dense) + factors
    def _rows(self, x, cols_fn, y=None, n_major: bool = False, want_factors: bool = False):
        """Dense rows ``Z[(col, n), :] = sum_c cols[col, n, c] * d f_c(x_n)/d theta`` for all trainable params
        (``parameters()`` order, row-major flattening; curvature/curvature.py:115-124).
        Returns ``(Z [ncols, M, P] or [M, ncols, P] if n_major, f, factors)``."""
        self._require_supported("jacobians()/full()/diag()")
        f = self._forward(x)
        fd = f.detach()
        M, C = fd.shape
        cols = cols_fn(fd, y)
        ncols = cols.shape[0]
        acts = self._acts
        grads = self._backward(f, cols)
        self._acts = {}
        sizes = []
        for L in self._layers:
            if L.has_w:
                sizes.append(L.d_out * L.d_in)
            if L.has_b:
                sizes.append(L.d_out)
        P = sum(sizes)
        Z = torch.empty((M, ncols, P) if n_major else (ncols, M, P), device=fd.device, dtype=torch.float32)
        sn, sc = (ncols * P, P) if n_major else (P, M * P)
        blocks, off = [], 0
        for L, g in zip(self._layers, grads):
            a = acts[L.name]
            a = a.float() if a.dtype != torch.float32 else a
            shared = L.is_conv or a.dim() > 2
            off_w = off if L.has_w else -1
            off_b = (off + (L.d_out * L.d_in if L.has_w else 0)) if L.has_b else -1
            if not shared:
                g = g.contiguous()
                K.jac_linear_write(g, a.contiguous(), Z, sn, sc, off_w, off_b)
                if L.has_w:
                    blocks.append(("outer", g, a.contiguous()))
                if L.has_b:
                    blocks.append(("vec", g))
            else:
                if L.has_w:
                    A, T = self._pack_act(L, a, K.F32)
                    G = self._pack_grad(L, g, K.F32)
                    K.shared_weight_contract(1, G, A, L.d_out, L.d_in, T, M, ncols, Z[..., off_w:], js_stride_n=sn,
                                             js_stride_c=sc)
                    if want_factors:
                        blocks.append(("dense", None))
                if L.has_b:
                    gb = g.reshape(ncols, M, L.d_out, -1).sum(-1) if L.is_conv else g.reshape(ncols, M, -1, L.d_out).sum(2)
                    gb = gb.contiguous()
                    dummy = torch.zeros(M, 1, device=fd.device, dtype=torch.float32)
                    K.jac_linear_write(gb, dummy, Z, sn, sc, -1, off_b)
                    if want_factors:
                        blocks.append(("vec", gb))
            off += (L.d_out * L.d_in if L.has_w else 0) + (L.d_out if L.has_b else 0)
        factors = None
        if want_factors and n_major:
            # fill dense blocks with views of the materialised rows
            fixed, o = [], 0
            for blk, p in zip(blocks, sizes):
                fixed.append(("dense", Z[:, :, o:o + p].contiguous()) if blk[0] == "dense" else blk)
                o += p
            factors = JacobianFactors(fixed, M, ncols, sizes)
        return Z, fd, factors

    # ------------------------------------------------------------------ Jacobian factors without the dense tensor
    def _factor_blocks(self, x, cols_fn):
        """Per-parameter-block structure of ``d f / d theta`` for a batch WITHOUT forming the dense ``(M, C, P)`` tensor
        (447 MB per sample for ResNet-18): ``("outer", g, a)`` / ``("vec", g)`` for layers without weight sharing,
        ``("conv", G_rows, A_rows, T)`` for convolutions / token-shared linears -- ``J_{n,c} = G_{n,c}^T A_n`` with
        ``G_rows [(c,n,t), d_out]`` the output-gradient rows and ``A_rows [(n,t), d_in]`` the unfolded input rows
        (SURVEY App. A).  Returns ``(f, JacobianFactors)``."""
        self._require_supported("jacobians()")
        f = self._forward(x)
        fd = f.detach()
        M, C = fd.shape
        cols = cols_fn(fd, None)
        ncols = cols.shape[0]
        acts = self._acts
        grads = self._backward(f, cols)
        self._acts = {}
        blocks, sizes = [], []
        for L, g in zip(self._layers, grads):
            a = acts[L.name]
            a = a.float() if a.dtype != torch.float32 else a
            shared = L.is_conv or a.dim() > 2
            if not shared:
                g = g.contiguous()
                if L.has_w:
                    blocks.append(("outer", g, a.contiguous()))
                    sizes.append(L.d_out * L.d_in)
                if L.has_b:
                    blocks.append(("vec", g))
                    sizes.append(L.d_out)
                continue
            if L.is_conv:
                T = g.shape[3] * g.shape[4]
                Grows = g.permute(0, 1, 3, 4, 2).reshape(ncols * M * T, L.d_out).contiguous()
                gb = g.reshape(ncols, M, L.d_out, -1).sum(-1).contiguous() if L.has_b else None
            else:
                T = a.numel() // (M * L.d_in)
                Grows = g.reshape(ncols * M * T, L.d_out).contiguous()
                gb = g.reshape(ncols, M, -1, L.d_out).sum(2).contiguous() if L.has_b else None
            if L.has_w:
                if L.is_conv:
                    P = K.pack_conv_rows(a, L.mod, K.F32)
                    Arows = P.hi if P.hi.shape[1] == L.d_in else P.hi[:, :L.d_in].contiguous()   # drop the row padding
                else:
                    Arows = a.reshape(M * T, L.d_in).contiguous()
                blocks.append(("conv", Grows, Arows, T))
                sizes.append(L.d_out * L.d_in)
            if L.has_b:
                blocks.append(("vec", gb))
                sizes.append(L.d_out)
        return fd, JacobianFactors(blocks, M, ncols, sizes)

    def _identity_cols(self, f, y=None):
        M, C = f.shape
        return torch.eye(C, device=f.device, dtype=f.dtype).unsqueeze(1).expand(C, M, C).contiguous()

    # ------------------------------------------------------------------ public: jacobians
    def cached_jacobians(self, max_batches: int = 64):
        """Context manager: memoise ``jacobians(x)`` by input content for the duration of the block.

        The reference's prior-precision grid search (``optimize_prior_precision(method="gridsearch")``,
        baselaplace.py:516-561 -> utils/utils.py:39-101) re-runs the whole GLM predictive over the validation set for each
        of its 100 grid values, Jacobians included, although only ``deltas`` changes.  Inside this block the second and
        later passes over the same batches return the first pass' Jacobians (and, through them, their cached eigenbasis
        projections, ``JacobianFactors.projections``): ``with la.backend.cached_jacobians(): la.optimize_prior_precision(
        pred_type="glm", method="gridsearch", val_loader=...)`` -- host code untouched, 100x fewer Jacobian passes.
        Keyed by shape + two checksums of the batch (one small device->host read per call); tensor inputs only."""
        backend = self

        class _Ctx:
            def __enter__(self_ctx):
                backend._jac_cache = {}
                backend._jac_cache_max = max_batches
                return backend

            def __exit__(self_ctx, *exc):
                backend._jac_cache = None
                return False

        return _Ctx()

    @staticmethod
    def _content_key(x: torch.Tensor):
        xf = x.detach().reshape(-1)
        xf = xf.float() if xf.dtype not in (torch.float32, torch.float64) else xf
        n = xf.numel()
        w = torch.arange(1, 1 + min(n, 4096), device=x.device, dtype=xf.dtype)
        sig = torch.stack([xf.sum(), (xf * xf).sum(), (xf[:w.numel()] * w).sum()]).tolist()
        return (tuple(x.shape), x.dtype, str(x.device)) + tuple(sig)

    def jacobians(self, x, enable_backprop: bool = False):
        """``CurvatureInterface.jacobians`` (curvature/curvature.py:88-129): ``Js (B, C, P)``, ``f (B, C)``."""
        if enable_backprop:
            return self._reference_fallback("jacobians", x, enable_backprop=True)
        cache = getattr(self, "_jac_cache", None)
        key = None
        if cache is not None and torch.is_tensor(x):
            key = self._content_key(x)
            if key in cache:
                return cache[key]
        out = self._jacobians_impl(x)
        if key is not None and len(cache) < self._jac_cache_max:
            fac = getattr(out[0], "_lpb_factors", None)
            if fac is not None:
                fac.keep_projections = True
            cache[key] = out
        return out

    def _jacobians_impl(self, x):
        dtype = next(self.model.parameters()).dtype
        if self.subnetwork_indices is None and dtype == torch.float32 and self._lazy_jacobian(x):
            from .predictive import LazyJacobian

            f, factors = self._factor_blocks(x, self._identity_cols)
            return LazyJacobian.wrap(factors, f.device), f
        Z, f, factors = self._rows(x, self._identity_cols, n_major=True, want_factors=self.subnetwork_indices is None)
        dtype = next(self.model.parameters()).dtype
        Js = Z.to(dtype)
        if self.subnetwork_indices is not None:
            Js = Js[:, :, self.subnetwork_indices]
        elif factors is not None and dtype == torch.float32:
            Js._lpb_factors = factors
        return Js, f.to(dtype)

    functorch_jacobians = jacobians

    # dense (B, C, P) Jacobians above this many bytes are returned as a ``LazyJacobian`` (factors only; materialised on
    # demand by any operation that is not one of the structured predictive contractions)
    LAZY_JACOBIAN_BYTES = 1 << 30

    def _lazy_jacobian(self, x) -> bool:
        if getattr(self, "lazy_jacobians", None) is not None:
            return bool(self.lazy_jacobians)
        xb = x[self.dict_key_x] if isinstance(x, MutableMapping) else x
        M = int(xb.shape[0])
        P = sum(p.numel() for p in self.params)
        C = getattr(self.model, "output_size", None) or getattr(self, "_n_outputs", None)   # baselaplace.py:956-962
        if C is None:
            return False
        return 4 * M * int(C) * P > self.LAZY_JACOBIAN_BYTES

    def last_layer_jacobians(self, x, enable_backprop: bool = False):
        """``CurvatureInterface.last_layer_jacobians`` (curvature/curvature.py:131-167)."""
        if enable_backprop:
            return self._reference_fallback("last_layer_jacobians", x, enable_backprop=True)
        with torch.no_grad(), self._model_numerics(), self._conv_patch():
            f, phi = self.model.forward_with_features(x)
        self._device_check(f)
        C = int(f.numel() / phi.shape[0])
        has_bias = self.model.last_layer.bias is not None
        Js = K.ll_jacobian_write(phi.detach().float(), C, has_bias)
        dtype = next(self.model.parameters()).dtype
        from .predictive import StructuredJacobian

        Js = StructuredJacobian.wrap(Js.to(dtype), (phi.detach().float().contiguous(), C, has_bias))
        return Js, f.detach()

    def gradients(self, x, y):
        """``CurvatureInterface.gradients`` (curvature/curvature.py:169-210): per-sample loss gradients ``(B, P)``."""
        y = y.to(next(self.model.parameters()).device)
        Z, f, _ = self._rows(x, self._loss_grad_cols, y=y)
        Gs = Z[0]
        if self.subnetwork_indices is not None:
            Gs = Gs[:, self.subnetwork_indices]
        dtype = next(self.model.parameters()).dtype
        return Gs.to(dtype), self.lossfunc(f, y).detach()

    def _reference_fallback(self, name, *args, **kwargs):
        from .interface import HAVE_REFERENCE

        if not HAVE_REFERENCE:
            raise NotImplementedError(f"{name} with enable_backprop=True needs the reference's torch.func path "
                                      "(the B200 kernels are not differentiable)")
        return getattr(CurvatureInterface, name)(self, *args, **kwargs)

    # ------------------------------------------------------------------ dense SYRK of rows
    def _rows_syrk(self, Z2d: torch.Tensor, alpha: float) -> torch.Tensor:
        Kr, P = Z2d.shape
        H = torch.zeros(P, P, device=Z2d.device, dtype=torch.float32)
        Zt = K.pack_rows(Z2d, self._kind(P, Kr))
        K.gemm_nt(Zt, Zt, H, alpha=alpha, accumulate=True, symmetric=True)
        return H

    # ------------------------------------------------------------------ last-layer structured curvature
    def _ll_forward(self, x):
        with torch.no_grad(), self._model_numerics(), self._conv_patch():
            f, phi = self.model.forward_with_features(x)
        self._device_check(f)
        if f.ndim != 2 or phi.ndim != 2:
            raise ValueError("last-layer curvature needs (batch, outputs) logits and (batch, features) features")
        return f.detach(), phi.detach().float().contiguous()

    def _ll_full(self, phi: torch.Tensor, Lam: torch.Tensor, scale: float) -> torch.Tensor:
        """``sum_n Lam_n (x) [phi;1][phi;1]^T`` in the reference's last-layer ordering ``[vec(W); b]``
        (structured form of curvature/curvature.py:398-408 with last_layer=True)."""
        M, D = phi.shape
        C = Lam.shape[1]
        has_bias = self.model.last_layer.bias is not None
        phit = torch.cat([phi, torch.ones(M, 1, device=phi.device, dtype=phi.dtype)], 1).contiguous() if has_bias else phi
        Dt = phit.shape[1]
        iu = torch.triu_indices(C, C, device=phi.device)
        w = Lam[:, iu[0], iu[1]].t().contiguous().float()            # [npairs, M]
        npairs = w.shape[0]
        kind = self._kind(Dt, M)
        A = K.pack_rows(phit, kind)
        Bw = K.pack_rows(phit, kind, row_scale=w.reshape(-1), nrep=npairs)
        G = torch.zeros(Dt, npairs * Dt, device=phi.device, dtype=torch.float32)
        K.gemm_nt(A, Bw, G, alpha=scale, accumulate=True)
        P = C * D + (C if has_bias else 0)
        H = torch.empty(P, P, device=phi.device, dtype=torch.float32)
        K.ll_ggn_expand(G, C, D, has_bias, H, accumulate=False)
        return H

    def _ll_diag(self, phi: torch.Tensor, lam_diag: torch.Tensor, scale: float) -> torch.Tensor:
        """diag of the above: ``sum_n Lam_n[c,c] * [phi;1]^2`` -> ``[C*D (+C)]``."""
        has_bias = self.model.last_layer.bias is not None
        M = phi.shape[0]
        phit = torch.cat([phi, torch.ones(M, 1, device=phi.device, dtype=phi.dtype)], 1).contiguous() if has_bias else phi
        lam_diag = lam_diag.float().contiguous()                       # [C, M] is already K-major
        A = K.Packed(lam_diag, None, K.F32, lam_diag.shape[0], lam_diag.shape[1])
        Bq = K.pack_rows(phit, K.F32, square=True)                     # [Dt, M]
        out = torch.zeros(lam_diag.shape[0], phit.shape[1], device=phi.device, dtype=torch.float32)
        K.gemm_nt(A, Bq, out, alpha=scale, accumulate=True)
        D = phi.shape[1]
        return torch.cat([out[:, :D].reshape(-1), out[:, D]]) if has_bias else out.reshape(-1)

    def _out_dtype(self, t: torch.Tensor) -> torch.Tensor:
        dtype = next(self.model.parameters()).dtype
        return t if t.dtype == dtype else t.to(dtype)


# =========================================================================================
class B200GGN(_B200Mixin, GGNInterface):
    """GGN / Fisher curvature on B200 (drop-in for ``CurvlinopsGGN``, curvature/curvlinops.py:144-168).

    Parameters beyond the reference's: ``precision`` in {"auto", "fp32", "bf16", "bf16x3"} selects the
    operand format of the tensor-core contractions (fp32 accumulation always); ``batched_backward=False``
    falls back to one reverse pass per column for models with ops that lack a vmap rule; ``fuse_elementwise=False``
    keeps ``conv -> frozen BatchNorm -> ReLU`` chains as separate reverse-pass kernels in ``kron()`` (required only if a
    convolution output feeds a BatchNorm/ReLU *and* another operation, which the backend detects and reports).
    """

    def __init__(self, model, likelihood, last_layer=False, subnetwork_indices=None, dict_key_x="input_ids",
                 dict_key_y="labels", stochastic=False, num_samples=1, precision="auto", batched_backward=True,
                 model_tf32=False, conv_engine=True, fuse_elementwise=True, cuda_graph=False):
        GGNInterface.__init__(self, model, likelihood, last_layer, subnetwork_indices, dict_key_x, dict_key_y,
                              stochastic, num_samples)
        self._b200_init(precision, batched_backward, model_tf32, conv_engine, fuse_elementwise, cuda_graph)

    def _ggn_cols(self, f, y=None, mc_samples=None):
        return self._mc_cols(f, mc_samples or self.num_samples) if self.stochastic else self._hessian_sqrt_cols(f)

    def _cols_fn(self, kwargs):
        """``mc_samples=S`` passed per call (as to ``CurvlinopsGGN.full`` / ``.kron``, curvlinops.py:92-98, 121) overrides the
        constructor's ``num_samples`` for a stochastic backend."""
        S = kwargs.get("mc_samples")
        if not self.stochastic or S is None:
            return self._ggn_cols
        return lambda f, y=None: self._ggn_cols(f, y, int(S))

    def _functional_hessian(self, f, mc_samples=None):
        if self.stochastic:
            cols = self._mc_cols(f, mc_samples or self.num_samples)               # [S, M, C]
            return torch.einsum("smc,smk->mck", cols, cols)
        if self.likelihood == "regression":
            return torch.eye(f.shape[1], device=f.device, dtype=f.dtype).expand(f.shape[0], -1, -1)
        p = torch.softmax(f, dim=-1)
        return torch.diag_embed(p) - p.unsqueeze(2) * p.unsqueeze(1)

    def kron(self, x, y, N, **kwargs: Any):
        """``CurvlinopsInterface.kron`` (curvature/curvlinops.py:77-108) with ``FisherType.TYPE2`` /
        ``FisherType.MC`` (:162-164): returns ``(factor * loss, Kron)``."""
        approx = kwargs.get("kfac_approx", "expand")
        if self.stochastic:
            S = int(kwargs.get("mc_samples", 1))
            sq2 = 2.0 if self.likelihood == "regression" else 1.0   # MSE-sum Hessian is 2 I
            return self._kron_impl(x, y, N, lambda f, yy: self._mc_cols(f, S) * math.sqrt(S * sq2), 1.0 / S, approx)
        sq2 = math.sqrt(2.0) if self.likelihood == "regression" else 1.0
        return self._kron_graphed(x, y, N, lambda f, yy: self._hessian_sqrt_cols(f) * sq2, 1.0, approx)

    def full(self, x, y, **kwargs: Any):
        """``GGNInterface.full`` (curvature/curvature.py:375-411): ``H = sum_n J_n^T L_n J_n``; loss = factor * loss;
        no ``factor`` on ``H`` (identical in value to ``CurvlinopsGGN.full``, curvlinops.py:110-141)."""
        if self.last_layer:
            f, phi = self._ll_forward(x)
            y = y.to(f.device)
            H = self._ll_full(phi, self._functional_hessian(f, kwargs.get("mc_samples")), 1.0)
            return (self.factor * self.lossfunc(f, y)).detach(), self._out_dtype(H)
        y = y.to(next(self.model.parameters()).device)
        Z, f, _ = self._rows(x, self._cols_fn(kwargs))
        P = Z.shape[-1]
        Z2 = Z.reshape(-1, P)
        if self.subnetwork_indices is not None:
            Z2 = Z2[:, self.subnetwork_indices].contiguous()
        H = self._rows_syrk(Z2, 1.0)
        return (self.factor * self.lossfunc(f, y)).detach(), self._out_dtype(H)

    def diag(self, x, y, **kwargs: Any):
        """``GGNInterface.diag`` (curvature/curvature.py:413-433)."""
        if self.last_layer:
            f, phi = self._ll_forward(x)
            y = y.to(f.device)
            lam = torch.diagonal(self._functional_hessian(f, kwargs.get("mc_samples")), dim1=1, dim2=2).t()   # [C, M]
            return (self.factor * self.lossfunc(f, y)).detach(), self._out_dtype(self._ll_diag(phi, lam, 1.0))
        y = y.to(next(self.model.parameters()).device)
        loss_f, d = self._diag_impl(x, self._cols_fn(kwargs), None, 1.0)
        if self.subnetwork_indices is not None:
            d = d[self.subnetwork_indices]
        return (self.factor * self.lossfunc(loss_f, y)).detach(), self._out_dtype(d)

    # structured diagonal: never materialises (B, C, P)
    def _diag_impl(self, x, cols_fn, y, scale):
        self._require_supported("diag()")
        f = self._forward(x)
        fd = f.detach()
        M = fd.shape[0]
        cols = cols_fn(fd, y)
        ncols = cols.shape[0]
        acts = self._acts
        grads = self._backward(f, cols)
        self._acts = {}
        parts = []
        for L, g in zip(self._layers, grads):
            a = acts[L.name]
            a = a.float() if a.dtype != torch.float32 else a
            shared = L.is_conv or a.dim() > 2
            if L.has_w:
                out = torch.zeros(L.d_out, L.d_in, device=fd.device, dtype=torch.float32)
                if not shared:
                    # sum_n (sum_col g^2)[n, i] * a[n, j]^2   -- one GEMM on element-wise squares
                    g2 = K.pack_rows((g * g).sum(0).contiguous(), K.F32)
                    a2 = K.pack_rows(a.contiguous(), K.F32, square=True)
                    K.gemm_nt(g2, a2, out, alpha=scale, accumulate=True)
                elif self._diag_conv_tc(L, a, g, M, ncols, out, scale):
                    pass
                else:
                    A, T = self._pack_act(L, a, K.F32)
                    G = self._pack_grad(L, g, K.F32)
                    K.shared_weight_contract(0, G, A, L.d_out, L.d_in, T, M, ncols, out, scale=scale, out_ld=L.d_in)
                parts.append(out.reshape(-1))
            if L.has_b:
                gb = g if not shared else (g.reshape(ncols, M, L.d_out, -1).sum(-1) if L.is_conv
                                           else g.reshape(ncols, M, -1, L.d_out).sum(2))
                parts.append(scale * (gb * gb).sum((0, 1)))
        return fd, torch.cat(parts)


class B200EF(_B200Mixin, EFInterface):
    """Empirical Fisher on B200 (drop-in for ``CurvlinopsEF``, curvature/curvlinops.py:171-180)."""

    def __init__(self, model, likelihood, last_layer=False, subnetwork_indices=None, dict_key_x="input_ids",
                 dict_key_y="labels", precision="auto", batched_backward=True, model_tf32=False, conv_engine=True,
                 fuse_elementwise=True, cuda_graph=False):
        EFInterface.__init__(self, model, likelihood, last_layer, subnetwork_indices, dict_key_x, dict_key_y)
        self._b200_init(precision, batched_backward, model_tf32, conv_engine, fuse_elementwise, cuda_graph)

    def kron(self, x, y, N, **kwargs: Any):
        """``CurvlinopsInterface.kron`` with ``FisherType.EMPIRICAL`` (curvature/curvlinops.py:174-176)."""
        return self._kron_graphed(x, y, N, self._loss_grad_cols, 1.0, kwargs.get("kfac_approx", "expand"))

    def full(self, x, y, **kwargs: Any):
        """``EFInterface.full`` (curvature/curvature.py:467-493): ``factor * sum_n g_n g_n^T``."""
        if self.last_layer:
            f, phi = self._ll_forward(x)
            y = y.to(f.device)
            r = self._loss_grad_cols(f, y)[0]
            H = self._ll_full(phi, r.unsqueeze(2) * r.unsqueeze(1), self.factor)
            return (self.factor * self.lossfunc(f, y)).detach(), self._out_dtype(H)
        y = y.to(next(self.model.parameters()).device)
        Z, f, _ = self._rows(x, self._loss_grad_cols, y=y)
        Z2 = Z[0]
        if self.subnetwork_indices is not None:
            Z2 = Z2[:, self.subnetwork_indices].contiguous()
        H = self._rows_syrk(Z2, self.factor)
        return (self.factor * self.lossfunc(f, y)).detach(), self._out_dtype(H)

    def diag(self, x, y, **kwargs: Any):
        """``EFInterface.diag`` (curvature/curvature.py:495-505): ``factor * sum_n g_n^2``."""
        if self.last_layer:
            f, phi = self._ll_forward(x)
            y = y.to(f.device)
            r = self._loss_grad_cols(f, y)[0]
            d = self._ll_diag(phi, (r * r).t().contiguous(), self.factor)
            return (self.factor * self.lossfunc(f, y)).detach(), self._out_dtype(d)
        y = y.to(next(self.model.parameters()).device)
        f, d = B200GGN._diag_impl(self, x, self._loss_grad_cols, y, self.factor)
        if self.subnetwork_indices is not None:
            d = d[self.subnetwork_indices]
        return (self.factor * self.lossfunc(f, y)).detach(), self._out_dtype(d)
