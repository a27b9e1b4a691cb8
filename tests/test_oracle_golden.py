"""The CPU oracle against the committed golden vectors (generated from the unmodified
reference by tests/golden/make_golden.py) -- this is what pins the oracle."""
import pytest
import torch

from oracle import curvature_oracle as co
from oracle import kron_oracle as ko
from tests.fixtures import load, rel_fro

CASES = [(k, l) for k in ("mlp", "conv") for l in ("classification", "regression")]


@pytest.mark.parametrize("kind,lik", CASES)
def test_jacobians_ggn_ef(golden, kind, lik):
    model, X, y, rec = load(golden, kind, lik)
    Js, f = co.jacobians(model, X)
    assert torch.allclose(Js, rec["Js"], atol=1e-10)
    assert torch.allclose(f, rec["f"], atol=1e-12)
    loss, H = co.ggn_full(Js, f, y, lik)
    assert torch.allclose(loss, rec["ggn_loss"], rtol=1e-10)
    assert rel_fro(H, rec["ggn_full"]) < (1e-6 if kind == "conv" else 1e-12)
    _, d = co.ggn_diag(Js, f, y, lik)
    assert torch.allclose(d, rec["ggn_diag"], rtol=1e-9, atol=1e-14)
    Gs, gl = co.gradients(Js, f, y, lik)
    assert torch.allclose(Gs, rec["Gs"], atol=1e-10)
    assert torch.allclose(gl, rec["grad_loss"], rtol=1e-10)
    loss, Hef = co.ef_full(Js, f, y, lik)
    assert torch.allclose(loss, rec["ef_loss"], rtol=1e-10)
    if "ef_full" in rec:
        assert rel_fro(Hef, rec["ef_full"]) < 1e-12
    _, def_ = co.ef_diag(Js, f, y, lik)
    assert torch.allclose(def_, rec["ef_diag"], rtol=1e-9, atol=1e-14)


@pytest.mark.parametrize("kind,lik", CASES)
def test_last_layer(golden, kind, lik):
    model, X, y, rec = load(golden, kind, lik)
    feats = {}
    last = model[-1]
    h = last.register_forward_hook(lambda m, i, o: feats.__setitem__("phi", i[0].detach()))
    f = model(X).detach()
    h.remove()
    Js = co.last_layer_jacobians(feats["phi"], f.shape[-1], last.bias is not None)
    assert torch.allclose(Js, rec["ll_Js"], atol=1e-12)
    loss, H = co.ggn_full(Js, f, y, lik)
    assert torch.allclose(loss, rec["ll_ggn_loss"], rtol=1e-10)
    assert rel_fro(H, rec["ll_ggn_full"]) < 1e-12


def test_kron_algebra(golden):
    rec = golden["kron_algebra"]
    kfacs, W = rec["kfacs"], rec["W"]
    Qs, ls = ko.decompose(kfacs)
    for damping, tag in ((False, "plain"), (True, "damp")):
        for a, b in zip(ls, rec[f"{tag}_eigvals"]):
            for x, y in zip(a, b):
                assert torch.allclose(x, y, atol=1e-12)
        for name in ("scalar", "layer"):
            delta = rec[f"{tag}_{name}_delta"]
            lsc = ko.scale_eigenvalues(ls, 1.7)
            isf = ko.kron_inv_square_form(Qs, lsc, delta, W, damping)
            assert torch.allclose(isf, rec[f"{tag}_{name}_isf"], rtol=1e-9, atol=1e-12)
            assert torch.allclose(ko.kron_logdet(lsc, delta, damping), rec[f"{tag}_{name}_logdet"], rtol=1e-10)
            m05 = ko.kron_bmm(Qs, lsc, delta, W, -0.5, damping)
            assert torch.allclose(m05, rec[f"{tag}_{name}_bmm_m05"], rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("kind,lik", CASES)
def test_full_and_diag_predictive(golden, kind, lik):
    model, X, y, rec = load(golden, kind, lik)
    Js, f = co.jacobians(model, X)
    _, H = co.ggn_full(Js, f, y, lik)
    prior = torch.full((H.shape[0],), 0.7, dtype=H.dtype)
    Sigma = ko.full_posterior_covariance(H, prior)
    fv = ko.full_functional_variance(Js, Sigma)
    assert torch.allclose(fv, rec["la_full_f_var"], rtol=1e-7, atol=1e-12)
    _, d = co.ggn_diag(Js, f, y, lik)
    assert torch.allclose(d, rec["la_diag_H"], rtol=1e-9, atol=1e-14)
    fvd = ko.diag_functional_variance(Js, 1.0 / (d + 0.7))
    assert torch.allclose(fvd, rec["la_diag_f_var"], rtol=1e-8, atol=1e-13)
    if lik == "classification":
        assert torch.allclose(ko.probit_predictive(f, fv), rec["la_full_probit"], atol=1e-10)
