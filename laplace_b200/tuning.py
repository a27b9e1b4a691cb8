"""Prior-precision grid search on cached eigenbasis projections.

``BaseLaplace.optimize_prior_precision(method="gridsearch")`` (baselaplace.py:483-561) evaluates, for each of
``grid_size`` prior precisions, ``utils.validate`` (utils/utils.py:39-101): the full GLM predictive over the validation
loader -- a Jacobian pass and the eigenbasis rotations per batch PER grid value -- and keeps the value with the smallest
loss.  Only ``deltas`` changes between grid values.  Here every validation batch pays for its Jacobians and for their
projections into the Kronecker eigenbasis once (``JacobianFactors.projections``); a grid value then costs one small GEMM
and one row-pair reduction per parameter block.

Works on a fitted reference ``KronLaplace`` / ``FullLaplace`` / ``DiagLaplace`` driven by a B200 backend and on the
stand-alone ``B200Laplace``.  Classification with the probit link and regression are covered (the reference's defaults);
anything else should go through the reference's own loop inside ``backend.cached_jacobians()``.

Reference quirk reproduced on request: with the default metric (``RunningNLLMetric``, utils/metrics.py) the reference
never resets the metric between grid values, so the score of grid value ``i`` is the running mean over values ``0..i``
(baselaplace.py:528-559 + utils/utils.py:99-101).  ``running_metric=True`` mirrors that; the default scores every value
on its own.
"""
from __future__ import annotations

import math
from collections.abc import MutableMapping

import torch


def _predict(la, Js, f_mu, likelihood: str):
    f_var = la.functional_variance(Js)
    if likelihood == "regression":
        return f_mu, f_var
    kappa = 1.0 / torch.sqrt(1.0 + math.pi / 8.0 * torch.diagonal(f_var, dim1=1, dim2=2))   # probit, baselaplace.py:662-664
    return torch.softmax(kappa * f_mu, dim=-1), None


@torch.no_grad()
def gridsearch_prior_precision(la, val_loader, interval: torch.Tensor | None = None, log_prior_prec_min: float = -4,
                               log_prior_prec_max: float = 4, grid_size: int = 100, running_metric: bool = False,
                               set_result: bool = True, dict_key_y: str = "labels"):
    """Returns ``(best_prior_precision, losses [grid_size])``; NLL for classification (probit GLM predictive), summed MSE
    of the predictive mean for regression (the reference's default metrics).  Sets ``la.prior_precision`` to the winner
    unless ``set_result=False``.  Mapping batches (Hugging Face style) are passed to the model whole and their targets
    read from ``dict_key_y``, as ``utils.validate`` does (utils/utils.py:60-65)."""
    likelihood = str(getattr(la.likelihood, "value", la.likelihood))
    if likelihood == "reward_modeling":
        likelihood = "classification"
    dev = getattr(la, "_device", None) or next(la.model.parameters()).device
    if interval is None:
        interval = torch.logspace(log_prior_prec_min, log_prior_prec_max, grid_size)
    last_layer = type(la).__name__.endswith("LLLaplace") or getattr(la, "last_layer", False)
    batches = []
    la.model.eval()
    for data in val_loader:
        if isinstance(data, MutableMapping):
            X, y = data, data[dict_key_y]
        else:
            X, y = data
            X = X.to(dev)
        y = y.to(dev)
        with torch.enable_grad():
            Js, f_mu = la.backend.last_layer_jacobians(X) if last_layer else la.backend.jacobians(X)
        fac = getattr(Js, "_lpb_factors", None)
        if fac is not None:
            fac.keep_projections = True
        batches.append((Js, f_mu.detach(), y))
    keep = la.prior_precision
    losses = []
    tot_sum, tot_n = 0.0, 0.0
    for pp in interval:
        la.prior_precision = pp
        s = torch.zeros((), device=dev, dtype=torch.float64)
        n = 0
        for Js, f_mu, y in batches:
            pred, _ = _predict(la, Js, f_mu, likelihood)
            if likelihood == "regression":
                s += ((pred - y) ** 2).sum().double()
            else:
                s += torch.nn.functional.nll_loss(pred.log(), y, reduction="sum").double()
            n += len(y)
        s = float(s)
        if running_metric:
            tot_sum, tot_n = tot_sum + s, tot_n + n
            losses.append(tot_sum / tot_n)
        else:
            losses.append(s / n)
    losses_t = torch.tensor(losses, dtype=torch.float64)
    finite = torch.where(torch.isfinite(losses_t), losses_t, torch.full_like(losses_t, float("inf")))
    best = interval[int(torch.argmin(finite))]
    la.prior_precision = best if set_result else keep
    return best, losses_t
