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
