"""Stand-alone host driver (laplace_b200.posterior) on CPU with emulated kernels vs golden vectors / oracle."""
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

from laplace_b200 import B200GGN
from laplace_b200.posterior import B200Laplace
from oracle import curvature_oracle as co
from oracle import kron_oracle as ko
from tests.fixtures import load, rel_fro


@pytest.mark.parametrize("hs", ["full", "diag"])
@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_full_diag_vs_golden(golden, cpu_kernels, hs, lik):
    model, X, y, rec = load(golden, "mlp", lik)
    la = B200Laplace(model, lik, "all", hs, prior_precision=0.7).fit(DataLoader(TensorDataset(X, y), batch_size=4))
    f_mu, f_var = la.glm_predictive_distribution(X)
    assert torch.allclose(f_mu, rec[f"la_{hs}_f_mu"], atol=1e-6)
    assert torch.allclose(f_var, rec[f"la_{hs}_f_var"], rtol=1e-4, atol=1e-7)
    if lik == "classification":
        assert torch.allclose(la(X), rec[f"la_{hs}_probit"], atol=1e-5)


@pytest.mark.parametrize("kind", ["mlp", "conv"])
def test_kron_vs_oracle(golden, cpu_kernels, kind):
    model, X, y, _ = load(golden, kind, "classification", dtype=torch.float32)
    la = B200Laplace(model, "classification", "all", "kron", prior_precision=0.7).fit(
        DataLoader(TensorDataset(X, y), batch_size=5))
    f_mu, f_var = la.glm_predictive_distribution(X)
    md = model.double()
    kfs = None
    for i in range(0, len(X), 5):
        _, kf = co.kfac_factors(md, "classification", X[i:i + 5].double(), y[i:i + 5], N=len(X))
        kfs = kf if kfs is None else [[a + b for a, b in zip(Fa, Fb)] for Fa, Fb in zip(kfs, kf)]
    Qs, ls = ko.decompose(kfs)
    Js, _ = co.jacobians(md, X.double())
    ref = ko.kron_inv_square_form(Qs, ls, torch.tensor(0.7, dtype=torch.float64), Js)
    assert torch.allclose(f_var.double(), ref, rtol=1e-3, atol=1e-6)
    assert torch.allclose(la.log_det_posterior_precision.double(), ko.kron_logdet(ls, torch.tensor(0.7, dtype=torch.float64)), rtol=1e-4)


def test_last_layer_full_vs_golden(golden, cpu_kernels):
    model, X, y, rec = load(golden, "mlp", "classification")
    la = B200Laplace(model, "classification", "last_layer", "full", prior_precision=0.7).fit(
        DataLoader(TensorDataset(X, y), batch_size=4))
    assert rel_fro(la.H, rec["ll_ggn_full"]) < 1e-5
    f_mu, f_var = la.glm_predictive_distribution(X)
    Sigma = ko.full_posterior_covariance(rec["ll_ggn_full"], torch.full((la.n_params,), 0.7, dtype=torch.float64))
    ref = ko.full_functional_variance(rec["ll_Js"], Sigma)
    assert torch.allclose(f_var.double(), ref, rtol=1e-4, atol=1e-8)
    for hs in ("kron", "diag"):
        B200Laplace(model, "classification", "last_layer", hs).fit(DataLoader(TensorDataset(X, y), batch_size=4))(X)
