"""Posterior-predictive contractions with last-layer structure, and the tensor subclass that routes the reference's
own ``functional_variance`` einsums (baselaplace.py:1683-1684, :2113-2115) into them without touching host code.

``B200GGN.last_layer_jacobians`` returns the dense ``(B, C, P)`` Jacobian the reference expects, as a
``StructuredJacobian``: a ``torch.Tensor`` subclass that remembers ``phi`` (the features) and recognises

    torch.einsum("ncp,pq,nkq->nck", Js, Sigma, Js)      FullLaplace.functional_variance
    torch.einsum("ncp,p,nkp->nck",  Js, var,   Js)      DiagLaplace.functional_variance

For ``J_n = [I_C (x) phi_n^T, I_C]`` (curvature/curvature.py:157-165) these are ``C^2`` quadratic forms of size
``D + 1`` per sample instead of a dense ``C x P x P`` contraction on a 90 %-zero Jacobian.  Every other operation on
the tensor behaves like a plain tensor.
"""
from __future__ import annotations

import torch

from . import kernels as K


def _phit(phi: torch.Tensor, has_bias: bool) -> torch.Tensor:
    M = phi.shape[0]
    return torch.cat([phi, torch.ones(M, 1, device=phi.device, dtype=phi.dtype)], 1).contiguous() if has_bias else phi


def ll_full_variance(phi: torch.Tensor, C: int, has_bias: bool, Sigma: torch.Tensor, gathered: torch.Tensor | None = None):
    """``f_var[n,c,k] = [phi_n;1]^T Sigma_{(c,.),(k,.)} [phi_n;1]``: one GEMM-NT against the gathered covariance blocks
    (``ll_sigma_gather``) + a per-sample reduction.  Returns ``(f_var [M,C,C], gathered)`` (cache ``gathered``)."""
    M, D = phi.shape
    phit = _phit(phi, has_bias)
    Dt = phit.shape[1]
    if gathered is None:
        gathered = K.ll_sigma_gather(Sigma.float(), C, D, has_bias)            # [(c,k,et), dt]
    Y = torch.empty(M, C * C * Dt, device=phi.device, dtype=torch.float32)
    K.gemm_nt(K.Packed(phit, None, K.F32, M, Dt), K.Packed(gathered, None, K.F32, gathered.shape[0], Dt), Y, 1.0,
              accumulate=False)
    out = torch.empty(M, C * C, 1, device=phi.device, dtype=torch.float32)
    K.batched_pair_dot(Y.view(M, C * C, Dt), phit.view(M, 1, Dt), None, out)
    return out.view(M, C, C), gathered


def ll_diag_variance(phi: torch.Tensor, C: int, has_bias: bool, var: torch.Tensor) -> torch.Tensor:
    """Diagonal posterior: ``f_var[n,c,k] = delta_ck * sum_d [phi;1]_d^2 var[idx(c,d)]`` (one small GEMM on squares)."""
    M, D = phi.shape
    phit = _phit(phi, has_bias)
    var = var.float()
    V = var[:C * D].reshape(C, D)
    if has_bias:
        V = torch.cat([V, var[C * D:].reshape(C, 1)], 1)
    V = V.contiguous()
    out = torch.zeros(M, C, device=phi.device, dtype=torch.float32)
    # out[n, c] = sum_d phit[n,d]^2 V[c,d]  -> A = phit^2 rows [M, K=Dt], B = V [C, K=Dt]
    K.gemm_nt(K.Packed((phit * phit).contiguous(), None, K.F32, M, phit.shape[1]), K.Packed(V, None, K.F32, C, V.shape[1]),
              out, 1.0, accumulate=False)
    return torch.diag_embed(out)


class StructuredJacobian(torch.Tensor):
    """Dense last-layer Jacobian ``(B, C, P)`` that carries ``(phi, C, has_bias)`` and intercepts the reference's
    functional-variance einsums (see module docstring)."""

    @staticmethod
    def wrap(dense: torch.Tensor, ll) -> "StructuredJacobian":
        t = torch.Tensor._make_subclass(StructuredJacobian, dense)
        t._lpb_ll = ll
        return t

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.einsum and len(args) == 4 and isinstance(args[0], str) and not kwargs:
            eq = args[0].replace(" ", "")
            J1, Mid, J2 = args[1], args[2], args[3]
            ll = getattr(J1, "_lpb_ll", None)
            if J1 is J2 and ll is not None and isinstance(Mid, torch.Tensor) and not Mid.requires_grad and Mid.is_cuda == ll[0].is_cuda:
                phi, C, has_bias = ll
                if eq == "ncp,pq,nkq->nck" and Mid.dim() == 2:
                    return ll_full_variance(phi, C, has_bias, Mid)[0].to(J1.dtype)
                if eq == "ncp,p,nkp->nck" and Mid.dim() == 1:
                    return ll_diag_variance(phi, C, has_bias, Mid).to(J1.dtype)
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **kwargs)


class LazyJacobian(torch.Tensor):
    """``(B, C, P)`` Jacobian that exists only as its per-layer factors (``laplace_b200.matrix.JacobianFactors``).

    ``B200GGN.jacobians`` returns it when the dense tensor would exceed ``B200GGN.LAZY_JACOBIAN_BYTES`` (ResNet-18: 447 MB per
    sample).  ``KronLaplace.functional_variance(Js)`` = ``posterior_precision.inv_square_form(Js)`` (baselaplace.py:1834-1835)
    reads the factors and never touches the data; shape / dtype / device queries answer from the metadata; ANY other
    operation first materialises the dense tensor (``dense()``) and proceeds on it, so the object is always correct to
    use -- just as memory-hungry as the reference's Jacobian when used outside the structured path."""

    @staticmethod
    def wrap(factors, device) -> "LazyJacobian":
        P = sum(factors.sizes)
        base = torch.zeros(1, device=device, dtype=torch.float32).expand(factors.n_batch, factors.n_out, P)   # no storage
        t = torch.Tensor._make_subclass(LazyJacobian, base)
        t._lpb_factors = factors
        t._lpb_dense = None
        return t

    def dense(self) -> torch.Tensor:
        if self._lpb_dense is None:
            from .matrix import materialize_block

            fac = self._lpb_factors
            parts = [materialize_block(blk, fac.n_batch, fac.n_out) for blk in fac.blocks]
            self._lpb_dense = torch.cat(parts, dim=2)
        return self._lpb_dense

    _PASSIVE = {"__get__", "size", "dim", "numel", "stride", "is_contiguous", "__repr__", "__str__", "__format__", "data_ptr",
                "is_floating_point", "element_size", "_is_view", "untyped_storage"}

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if name in cls._PASSIVE:
            if name in ("__repr__", "__str__", "__format__"):
                lz = next(a for a in args if isinstance(a, LazyJacobian))
                return f"LazyJacobian(shape={tuple(lz.shape)}, device={lz.device}, blocks={len(lz._lpb_factors.blocks)})"
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if name == "detach" and len(args) == 1:
            return args[0]

        def unwrap(o):
            if isinstance(o, LazyJacobian):
                return o.dense()
            if isinstance(o, (tuple, list)):
                return type(o)(unwrap(v) for v in o)
            return o

        with torch._C.DisableTorchFunctionSubclass():
            return func(*unwrap(args), **{k: unwrap(v) for k, v in kwargs.items()})
