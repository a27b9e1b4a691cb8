"""CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT) for the per-batch curvature path.

Plain torch on CPU; every function states the reference lines it restates.
Straightforward loops are used on purpose -- this code is the checker, it is
never the thing that is measured or shipped (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
import torch.nn.functional as F
from torch import nn

SUPPORTED = (nn.Linear, nn.Conv2d)


# --------------------------------------------------------------------------
# likelihood pieces
# --------------------------------------------------------------------------
def likelihood_factor(likelihood: str) -> float:
    """``CurvatureInterface.__init__`` (curvature/curvature.py:63-72): 0.5 for the
    MSE-sum loss, 1.0 for the CE-sum loss."""
    return 0.5 if likelihood == "regression" else 1.0


def sum_loss(f: torch.Tensor, y: torch.Tensor, likelihood: str) -> torch.Tensor:
    """``MSELoss(reduction='sum')`` / ``CrossEntropyLoss(reduction='sum')``
    (curvature/curvature.py:63-72)."""
    if likelihood == "regression":
        return ((f - y) ** 2).sum()
    return F.cross_entropy(f, y, reduction="sum")


def functional_hessian(f: torch.Tensor, likelihood: str) -> torch.Tensor | None:
    """``GGNInterface._get_functional_hessian`` (curvature/curvature.py:366-373):
    ``diag(p) - p p^T`` for softmax-CE, identity (returned as ``None``) for regression."""
    if likelihood == "regression":
        return None
    p = torch.softmax(f, dim=-1)
    return torch.diag_embed(p) - p.unsqueeze(2) * p.unsqueeze(1)


def loss_hessian_sqrt(f: torch.Tensor, likelihood: str) -> torch.Tensor:
    """A square root ``S_n`` (``S_n S_n^T = d^2 loss / df^2``) of the *sum-reduced torch
    loss* Hessian: CE -> ``diag(sqrt p) - p sqrt(p)^T``; MSE-sum -> ``sqrt(2) I``.
    KFAC type-2 back-propagates the columns of ``S_n`` (curvlinops
    2.0.0 ``KFACLinearOperator`` with ``FisherType.TYPE2``, call site
    curvature/curvlinops.py:87-100).  ``sum_c g_c g_c^T`` only depends on ``S S^T``."""
    B, C = f.shape
    if likelihood == "regression":
        return math.sqrt(2.0) * torch.eye(C, dtype=f.dtype).expand(B, C, C).clone()
    p = torch.softmax(f, dim=-1)
    sp = p.sqrt()
    return torch.diag_embed(sp) - p.unsqueeze(2) * sp.unsqueeze(1)


def loss_gradient(f: torch.Tensor, y: torch.Tensor, likelihood: str) -> torch.Tensor:
    """d(sum loss)/df per sample: CE -> ``p - onehot(y)``, MSE-sum -> ``2 (f - y)``."""
    if likelihood == "regression":
        return 2.0 * (f - y)
    p = torch.softmax(f, dim=-1)
    return p - F.one_hot(y, f.shape[-1]).to(f.dtype)


# --------------------------------------------------------------------------
# Jacobians / gradients (curvature/curvature.py:88-210)
# --------------------------------------------------------------------------
def jacobians(model: nn.Module, x: torch.Tensor, params: Sequence[nn.Parameter] | None = None):
    """``CurvatureInterface.jacobians`` (curvature/curvature.py:88-129) restated as the
    naive loop of the reference's own test oracle (tests/utils.py:85-106): one reverse
    pass per (sample, output).  Returns ``Js (B, C, P)``, ``f (B, C)``; parameters are
    concatenated in ``parameters()`` order, each flattened row-major."""
    if params is None:
        params = [p for p in model.parameters() if p.requires_grad]
    f = model(x)
    B, C = f.shape
    rows = []
    for n in range(B):
        per_out = []
        for c in range(C):
            gs = torch.autograd.grad(f[n, c], params, retain_graph=True, allow_unused=True)
            per_out.append(
                torch.cat([
                    (g if g is not None else torch.zeros_like(p)).reshape(-1)
                    for g, p in zip(gs, params)
                ])
            )
        rows.append(torch.stack(per_out))
    return torch.stack(rows).detach(), f.detach()


def last_layer_jacobians(phi: torch.Tensor, n_outputs: int, has_bias: bool) -> torch.Tensor:
    """``CurvatureInterface.last_layer_jacobians`` (curvature/curvature.py:131-167):
    ``J_n = [I_C (x) phi_n^T , I_C]`` with the weight flattened row-major ``(C, D)``."""
    B, D = phi.shape
    C = n_outputs
    Js = torch.zeros(B, C, C * D + (C if has_bias else 0), dtype=phi.dtype)
    for c in range(C):
        Js[:, c, c * D:(c + 1) * D] = phi
        if has_bias:
            Js[:, c, C * D + c] = 1.0
    return Js


def gradients(Js: torch.Tensor, f: torch.Tensor, y: torch.Tensor, likelihood: str):
    """``CurvatureInterface.gradients`` (curvature/curvature.py:169-210): per-sample
    gradients of the sum loss, ``G_n = J_n^T dl/df_n``; loss = sum over the batch."""
    r = loss_gradient(f, y, likelihood)
    Gs = (Js * r.unsqueeze(2)).sum(1)
    return Gs, sum_loss(f, y, likelihood)


# --------------------------------------------------------------------------
# dense / diagonal GGN and EF (curvature/curvature.py:375-505)
# --------------------------------------------------------------------------
def ggn_full(Js, f, y, likelihood, H_lik=None):
    """``GGNInterface.full`` (curvature/curvature.py:375-411): ``H = sum_n J_n^T L_n J_n``
    (``L_n`` = functional Hessian, identity for regression); loss = factor * sum-loss;
    no ``factor`` on ``H``."""
    if H_lik is None:
        H_lik = functional_hessian(f, likelihood)
    P = Js.shape[-1]
    H = torch.zeros(P, P, dtype=Js.dtype)
    for n in range(Js.shape[0]):
        Jn = Js[n]
        H += Jn.T @ Jn if H_lik is None else Jn.T @ H_lik[n] @ Jn
    return likelihood_factor(likelihood) * sum_loss(f, y, likelihood), H


def ggn_diag(Js, f, y, likelihood, H_lik=None):
    """``GGNInterface.diag`` (curvature/curvature.py:413-433)."""
    if H_lik is None:
        H_lik = functional_hessian(f, likelihood)
    d = torch.zeros(Js.shape[-1], dtype=Js.dtype)
    for n in range(Js.shape[0]):
        Jn = Js[n]
        d += (Jn * Jn).sum(0) if H_lik is None else ((H_lik[n] @ Jn) * Jn).sum(0)
    return likelihood_factor(likelihood) * sum_loss(f, y, likelihood), d


def ef_full(Js, f, y, likelihood):
    """``EFInterface.full`` (curvature/curvature.py:467-493): ``factor * sum_n g_n g_n^T``."""
    Gs, loss = gradients(Js, f, y, likelihood)
    fac = likelihood_factor(likelihood)
    return fac * loss, fac * (Gs.T @ Gs)


def ef_diag(Js, f, y, likelihood):
    """``EFInterface.diag`` (curvature/curvature.py:495-505)."""
    Gs, loss = gradients(Js, f, y, likelihood)
    fac = likelihood_factor(likelihood)
    return fac * loss, fac * (Gs * Gs).sum(0)


# --------------------------------------------------------------------------
# KFAC (curvature/curvlinops.py:46-108 + curvlinops 2.0.0 semantics)
# --------------------------------------------------------------------------
def kfac_layers(model: nn.Module, params: Sequence[nn.Parameter]):
    """Modules curvlinops' KFAC maps (``nn.Linear`` / ``nn.Conv2d`` owning at least one
    of ``params``), in ``named_modules()`` order -- the order
    ``CurvlinopsInterface._get_kron_factors`` walks (curvature/curvlinops.py:55-75)."""
    ids = {id(p) for p in params}
    out = []
    for name, mod in model.named_modules():
        own = [p for p in mod.parameters(recurse=False) if id(p) in ids]
        if not own:
            continue
        if not isinstance(mod, SUPPORTED):
            raise ValueError(f"KFAC supports nn.Linear / nn.Conv2d only, got {type(mod).__name__} ({name})")
        out.append((name, mod))
    return out


def layer_input_rows(mod: nn.Module, a: torch.Tensor) -> torch.Tensor:
    """``(B, T, d_in)`` rows of a layer input: conv -> unfolded ``C_in*kh*kw`` patches at
    every output position; linear -> tokens (``T = 1`` for 2-D inputs)."""
    if isinstance(mod, nn.Conv2d):
        if mod.groups != 1:
            raise ValueError("grouped convolutions are not supported")
        cols = F.unfold(a, mod.kernel_size, dilation=mod.dilation, padding=mod.padding, stride=mod.stride)
        return cols.transpose(1, 2)  # (B, T, C_in*kh*kw)
    return a.reshape(a.shape[0], -1, a.shape[-1])


def layer_output_rows(mod: nn.Module, g: torch.Tensor) -> torch.Tensor:
    """``(B, T, d_out)`` rows of a layer-output gradient."""
    if isinstance(mod, nn.Conv2d):
        return g.flatten(2).transpose(1, 2)
    return g.reshape(g.shape[0], -1, g.shape[-1])


def kfac_factors(
    model: nn.Module,
    likelihood: str,
    x: torch.Tensor,
    y: torch.Tensor,
    N: int,
    fisher: str = "type2",
    kfac_approx: str = "expand",
    mc_samples: int = 1,
    params: Sequence[nn.Parameter] | None = None,
    generator: torch.Generator | None = None,
):
    """``CurvlinopsInterface.kron`` (curvature/curvlinops.py:77-108) restated.

    Per mapped layer with input rows ``a_{n,t}`` and output-gradient rows ``g_{n,t,c}``:

    * expand: ``A = 1/(M T) sum a a^T``,       ``B = sum_{n,t,c} g g^T``
    * reduce: ``A = 1/M sum_n abar abar^T`` (``abar = mean_t a``), ``B = sum_{n,c} (sum_t g)(sum_t g)^T``
    * adapter: blocks ``[B, A]`` (weight) then ``[B]`` (bias) (:55-75); ``A *= M/N`` (:46-53,
      :103); ``kron *= factor`` splits the scalar as ``factor**(1/len(F))`` over the factors
      of a block (utils/matrix.py:116-118); loss = ``factor * lossfunc(model(x), y)`` (:106-108).

    Returns ``(loss, kfacs)`` with ``kfacs`` a list (one entry per trainable parameter,
    ``parameters()`` order) of lists of tensors.
    """
    if params is None:
        params = [p for p in model.parameters() if p.requires_grad]
    ids = {id(p) for p in params}
    layers = kfac_layers(model, params)
    acts, outs, hooks = {}, {}, []
    for name, mod in layers:
        def hook(m, inp, out, name=name):
            acts[name] = inp[0].detach()
            outs[name] = out
        hooks.append(mod.register_forward_hook(hook))
    try:
        f = model(x)
    finally:
        for h in hooks:
            h.remove()
    M, C = f.shape
    fd = f.detach()
    if fisher == "type2":
        S = loss_hessian_sqrt(fd, likelihood)
        cols = [S[:, :, c] for c in range(C)]
        weight = 1.0
    elif fisher == "empirical":
        cols = [loss_gradient(fd, y, likelihood)]
        weight = 1.0
    elif fisher == "mc":
        cols = []
        for _ in range(mc_samples):
            if likelihood == "regression":
                eps = torch.randn(fd.shape, generator=generator, dtype=fd.dtype) * math.sqrt(0.5)
                cols.append(-2.0 * eps)
            else:
                p = torch.softmax(fd, -1)
                ys = torch.multinomial(p, 1, generator=generator).squeeze(1)
                cols.append(p - F.one_hot(ys, C).to(fd.dtype))
        weight = 1.0 / mc_samples
    else:
        raise ValueError(fisher)

    names = [n for n, _ in layers]
    Bs = {n: None for n in names}
    for col in cols:
        gs = torch.autograd.grad(f, [outs[n] for n in names], grad_outputs=col, retain_graph=True)
        for (name, mod), g in zip(layers, gs):
            rows = layer_output_rows(mod, g.detach())  # (M, T, d_out)
            rows = rows.sum(1) if kfac_approx == "reduce" else rows.reshape(-1, rows.shape[-1])
            cov = weight * rows.T @ rows
            Bs[name] = cov if Bs[name] is None else Bs[name] + cov

    fac = likelihood_factor(likelihood)
    kfacs = []
    for name, mod in layers:
        rows = layer_input_rows(mod, acts[name])  # (M, T, d_in)
        T = rows.shape[1]
        if kfac_approx == "reduce":
            rows = rows.mean(1)
            A = rows.T @ rows / M
        else:
            rows = rows.reshape(-1, rows.shape[-1])
            A = rows.T @ rows / (M * T)
        A = A * (M / N)
        Bf = Bs[name]
        if id(mod.weight) in ids:
            s = fac ** 0.5
            kfacs.append([s * Bf, s * A])
        if mod.bias is not None and id(mod.bias) in ids:
            kfacs.append([fac * Bf])
    loss = fac * sum_loss(fd, y, likelihood)
    return loss, kfacs


def kfacs_to_matrix(kfacs) -> torch.Tensor:
    """``Kron.to_matrix`` (utils/matrix.py:258-275): block-diag of ``F0 (x) F1`` / ``F0``."""
    return torch.block_diag(*[torch.kron(F[0], F[1]) if len(F) == 2 else F[0] for F in kfacs])


def kfacs_diag(kfacs) -> torch.Tensor:
    """``Kron.diag`` (utils/matrix.py:241-256)."""
    return torch.cat([
        torch.outer(F[0].diag(), F[1].diag()).reshape(-1) if len(F) == 2 else F[0].diag() for F in kfacs
    ])
