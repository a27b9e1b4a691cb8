"""The UNMODIFIED reference front end (``laplace.Laplace(...)``, installed under ``baseline/_ref`` by
``tools/install_reference.sh``) driving the B200 backend ON THE GPU: ``fit`` -> ``la.H += H_batch`` dispatch ->
``decompose`` -> ``posterior_precision`` -> ``la(x, pred_type="glm")`` / ``_glm_predictive_distribution`` /
``log_marginal_likelihood`` / ``optimize_prior_precision`` / ``sample``.  The CPU twin of
this file (``test_dropin_reference_cpu.py``) pins the wiring with emulated kernels; here the real kernels run behind
the real host code.  Skipped when the reference is not importable."""
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

from oracle import curvature_oracle as co
from oracle import kron_oracle as ko
from tests.fixtures import load, rel_fro

pytestmark = pytest.mark.gpu
DEV = "cuda"
FACTOR_TOL, VAR_TOL = 1e-4, 1e-5


@pytest.fixture(scope="module")
def lap():
    from laplace_b200.interface import HAVE_REFERENCE

    if not HAVE_REFERENCE:
        pytest.skip("reference front end not importable (no baseline/_ref, no /root/reference)")
    import laplace

    return laplace


def var_err(f_var, ref):
    return float((f_var.cpu().double() - ref).abs().max() / ref.abs().max())


def _loader(X, y, bs):
    return DataLoader(TensorDataset(X, y), batch_size=bs)


@pytest.mark.parametrize("kind", ["mlp", "conv"])
@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_kron_laplace_fit_predict_marglik(golden, lap, kind, lik):
    from laplace_b200 import B200GGN, B200Kron, B200KronDecomposed

    model, X, y, _ = load(golden, kind, lik)
    kfs = None
    for i in range(0, len(X), 5):
        _, kf = co.kfac_factors(model, lik, X[i:i + 5], y[i:i + 5], N=len(X))
        kfs = kf if kfs is None else [[a + b for a, b in zip(Fa, Fb)] for Fa, Fb in zip(kfs, kf)]
    Qs, ls = ko.decompose(kfs)
    Js, f = co.jacobians(model, X)
    delta = torch.tensor(0.7, dtype=torch.float64)
    ref = ko.kron_inv_square_form(Qs, ls, delta, Js)
    m32 = model.float().to(DEV)
    y32 = y if y.dtype == torch.long else y.float()
    la = lap.Laplace(m32, lik, "all", "kron", backend=B200GGN, prior_precision=0.7)
    la.fit(_loader(X.float(), y32, 5))          # host batches: the reference moves them (baselaplace.py:974)
    assert isinstance(la.H_facs, B200Kron) and isinstance(la.H, B200KronDecomposed)
    assert la.H_facs._flat is not None, "first `la.H += H_batch` must keep the flat factor buffer"
    for F, Fo in zip(la.H_facs.kfacs, kfs):
        for H, Ho in zip(F, Fo):
            assert H.is_cuda and rel_fro(H.cpu(), Ho) < FACTOR_TOL
    f_mu, f_var = la._glm_predictive_distribution(X.float().to(DEV))
    assert torch.allclose(f_mu.cpu().double(), f, atol=1e-5)
    assert var_err(f_var, ref) < VAR_TOL, var_err(f_var, ref)
    assert torch.allclose(la.log_det_posterior_precision.cpu().double(), ko.kron_logdet(ls, delta), rtol=1e-5)
    if lik == "classification":
        probs = la(X.float().to(DEV), pred_type="glm", link_approx="probit")
        assert torch.allclose(probs.sum(-1), torch.ones(len(X), device=DEV), atol=1e-5)
    lml = la.log_marginal_likelihood()
    assert torch.isfinite(lml)
    # differentiable in the prior precision on the device (marglik optimisation, baselaplace.py:363-470)
    pp = torch.tensor([0.7], device=DEV, requires_grad=True)
    g, = torch.autograd.grad(la.log_marginal_likelihood(prior_precision=pp), pp)
    d_fd = 1e-3
    num = (la.log_marginal_likelihood(prior_precision=torch.tensor([0.7 + d_fd], device=DEV))
           - la.log_marginal_likelihood(prior_precision=torch.tensor([0.7 - d_fd], device=DEV))) / (2 * d_fd)
    assert torch.allclose(g.squeeze(), num.detach(), rtol=2e-2)
    la.prior_precision = 0.7
    # sampling: bmm(exponent=-1/2) on the device (baselaplace.py:1845-1879, utils/matrix.py:463-488)
    torch.manual_seed(0)
    S = la.sample(4096)
    assert S.shape == (4096, la.n_params) and S.is_cuda
    var_mc = (S - la.mean).pow(2).mean(0).cpu().double()                          # Monte-Carlo marginal variances
    assert rel_fro(var_mc, ko.kron_dense(Qs, ls, delta, exponent=-1.0).diagonal()) < 0.1
    W = torch.randn(3, 2, la.n_params, dtype=torch.float64)
    half = la.posterior_precision.bmm(W.float().to(DEV), exponent=-0.5)          # the map sample() applies
    assert rel_fro(half.cpu(), ko.kron_bmm(Qs, ls, delta, W, -0.5)) < 1e-5
    # state_dict round trip re-decomposes on load (baselaplace.py:1793-1809 / serialization)
    sd = la.state_dict()
    la2 = lap.Laplace(m32, lik, "all", "kron", backend=B200GGN, prior_precision=0.7)
    la2.load_state_dict(sd)
    from laplace_b200 import adopt

    assert isinstance(adopt(la2).H, B200KronDecomposed) and la2.H_facs._flat is not None
    _, f_var2 = la2._glm_predictive_distribution(X.float().to(DEV))
    assert var_err(f_var2, ref) < VAR_TOL


@pytest.mark.parametrize("hs", ["full", "diag"])
@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_full_diag_laplace_vs_golden(golden, lap, hs, lik):
    from laplace_b200 import B200GGN

    model, X, y, rec = load(golden, "mlp", lik, dtype=torch.float32)
    la = lap.Laplace(model.to(DEV), lik, "all", hs, backend=B200GGN, prior_precision=0.7)
    la.fit(_loader(X, y, 4))
    f_mu, f_var = la._glm_predictive_distribution(X.to(DEV))
    assert torch.allclose(f_mu.cpu().double(), rec[f"la_{hs}_f_mu"], atol=1e-5)
    assert var_err(f_var, rec[f"la_{hs}_f_var"]) < VAR_TOL
    assert torch.allclose(la.log_marginal_likelihood().cpu().double(), rec[f"la_{hs}_logmarglik"], rtol=1e-5)
    if lik == "classification":
        assert torch.allclose(la(X.to(DEV)).cpu().double(), rec[f"la_{hs}_probit"], atol=1e-5)


@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_last_layer_full_laplace_routes_einsums(golden, lap, lik):
    """FullLLLaplace through the reference's own ``functional_variance`` einsum (baselaplace.py:1683-1684), which the
    ``StructuredJacobian`` returned by ``last_layer_jacobians`` routes into the structured kernels."""
    from laplace_b200 import B200GGN, predictive

    calls = []
    orig = predictive.ll_full_variance
    predictive.ll_full_variance = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        model, X, y, rec = load(golden, "mlp", lik, dtype=torch.float32)
        la = lap.Laplace(model.to(DEV), lik, "last_layer", "full", backend=B200GGN, prior_precision=0.7)
        la.fit(_loader(X, y, 4))
        assert rel_fro(la.H.cpu(), rec["ll_ggn_full"]) < FACTOR_TOL
        f_mu, f_var = la._glm_predictive_distribution(X.to(DEV))
    finally:
        predictive.ll_full_variance = orig
    Sigma = ko.full_posterior_covariance(rec["ll_ggn_full"], torch.full((la.n_params,), 0.7, dtype=torch.float64))
    ref = ko.full_functional_variance(rec["ll_Js"], Sigma)
    assert var_err(f_var, ref) < VAR_TOL, var_err(f_var, ref)
    assert calls, "the reference's einsum was not routed into the structured last-layer kernel"


def test_kron_laplace_resnet_through_reference_front_end(lap):
    """Reduced-width ResNet-18 (fused conv->BN->ReLU chains forced on) behind ``Laplace(...).fit`` on the GPU."""
    from laplace_b200 import B200GGN, B200Kron, conv_engine, models

    model = models.make("resnet18", width=16)
    torch.manual_seed(5)
    X, y = torch.randn(128, 3, 32, 32), torch.randint(10, (128,))
    md = model.double()
    kfs = None
    for i in (0, 64):
        _, kf = co.kfac_factors(md, "classification", X[i:i + 64].double(), y[i:i + 64], N=128)
        kfs = kf if kfs is None else [[a + b for a, b in zip(Fa, Fb)] for Fa, Fb in zip(kfs, kf)]
    model = model.float().to(DEV)
    keep = conv_engine.ELEMENTWISE_MIN_BATCH
    conv_engine.ELEMENTWISE_MIN_BATCH = 0
    try:
        la = lap.Laplace(model, "classification", "all", "kron", backend=B200GGN, prior_precision=1.0)
        la.fit(_loader(X, y, 64))
        assert la.backend._fused
    finally:
        conv_engine.ELEMENTWISE_MIN_BATCH = keep
    assert isinstance(la.H_facs, B200Kron)
    worst = max(rel_fro(h.cpu(), ho) for F, Fo in zip(la.H_facs.kfacs, kfs) for h, ho in zip(F, Fo))
    assert worst < FACTOR_TOL, worst
    probs = la(X[:8].to(DEV), pred_type="glm", link_approx="probit")
    assert probs.shape == (8, 10) and torch.allclose(probs.sum(-1), torch.ones(8, device=DEV), atol=1e-5)


def test_prior_precision_gridsearch_gpu(lap):
    """SURVEY 8(f)1 on the device: the reference's own grid search (one Jacobian pass per grid value), the same call inside
    ``backend.cached_jacobians()`` and ``laplace_b200.tuning.gridsearch_prior_precision`` (one pass per validation batch)
    pick the same prior precision, and the loss curve matches the fp64 restatement of the reference loop to 1e-5."""
    from laplace_b200 import B200GGN
    from laplace_b200.tuning import gridsearch_prior_precision

    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(20, 64), torch.nn.ReLU(), torch.nn.Linear(64, 5))
    X, y = torch.randn(256, 20), torch.randint(5, (256,))
    Xv, yv = torch.randn(96, 20), torch.randint(5, (96,))
    md = torch.nn.Sequential(torch.nn.Linear(20, 64), torch.nn.ReLU(), torch.nn.Linear(64, 5)).double()
    md.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    kfs = None
    for i in range(0, 256, 64):
        _, kf = co.kfac_factors(md, "classification", X[i:i + 64].double(), y[i:i + 64], N=256)
        kfs = kf if kfs is None else [[a + b for a, b in zip(Fa, Fb)] for Fa, Fb in zip(kfs, kf)]
    Qs, ls = ko.decompose(kfs)
    Jb, fb, yb = [], [], []
    for i in range(0, 96, 48):
        J, f = co.jacobians(md, Xv[i:i + 48].double())
        Jb.append(J), fb.append(f), yb.append(yv[i:i + 48])
    G = 16
    la = lap.Laplace(model.to(DEV), "classification", "all", "kron", backend=B200GGN)
    la.fit(_loader(X, y, 64))
    vl = _loader(Xv, yv, 48)
    calls = {"n": 0}
    orig = la.backend._jacobians_impl
    la.backend._jacobians_impl = lambda x: (calls.__setitem__("n", calls["n"] + 1), orig(x))[1]
    with pytest.warns(UserWarning):
        la.optimize_prior_precision(pred_type="glm", method="gridsearch", val_loader=vl, grid_size=G)
    ref_pp, n_ref = la.prior_precision.clone(), calls["n"]
    calls["n"] = 0
    with la.backend.cached_jacobians(), pytest.warns(UserWarning):
        la.optimize_prior_precision(pred_type="glm", method="gridsearch", val_loader=vl, grid_size=G)
    assert torch.equal(la.prior_precision, ref_pp) and n_ref == 2 * G and calls["n"] == 2
    for running in (True, False):
        bo, lo = ko.gridsearch_prior_precision(Qs, ls, Jb, fb, yb, torch.logspace(-4, 4, G).double(), running_metric=running)
        b2, l2 = gridsearch_prior_precision(la, vl, grid_size=G, running_metric=running, set_result=False)
        assert torch.allclose(l2, lo, atol=1e-5, rtol=1e-5), (l2 - lo).abs().max()
        assert torch.allclose(b2.double().cpu(), bo)
    assert torch.allclose(ref_pp.cpu().double().squeeze(), ko.gridsearch_prior_precision(
        Qs, ls, Jb, fb, yb, torch.logspace(-4, 4, G).double(), running_metric=True)[0])
    # method="marglik": Adam on the device-side, differentiable log marginal likelihood (baselaplace.py:430-484)
    la.optimize_prior_precision(pred_type="glm", method="marglik", n_steps=20, prior_structure="layerwise")
    assert la.prior_precision.shape == (la.n_layers,) and torch.isfinite(la.prior_precision).all()


@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_marglik_training_gpu(lap, lik):
    """the reference's ``marglik_training`` loop with ``backend=B200GGN`` on the device."""
    from laplace.marglik_training import marglik_training

    from laplace_b200 import B200GGN, B200Kron

    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 32), torch.nn.Tanh(), torch.nn.Linear(32, 3)).to(DEV)
    X = torch.randn(128, 5)
    y = torch.randint(3, (128,)) if lik == "classification" else torch.randn(128, 3)
    la, model, margliks, losses = marglik_training(model, _loader(X, y, 32), lik, "kron", backend=B200GGN, n_epochs=4,
                                                   marglik_frequency=2, n_hypersteps=3)
    assert isinstance(la.H_facs, B200Kron) and la.H_facs.kfacs[0][0].is_cuda
    assert len(margliks) == 6 and margliks[-1] < margliks[0]
    if lik == "regression":
        assert float(la.sigma_noise) != 1.0
    # the marginal likelihood the loop optimised equals the fp64 restatement at the trained weights
    md = torch.nn.Sequential(torch.nn.Linear(5, 32), torch.nn.Tanh(), torch.nn.Linear(32, 3)).double()
    md.load_state_dict({k: v.detach().cpu().double() for k, v in model.state_dict().items()})
    yd = y if lik == "classification" else y.double()
    kfs = None
    for i in range(0, 128, 32):
        _, kf = co.kfac_factors(md, lik, X[i:i + 32].double(), yd[i:i + 32], N=128)
        kfs = kf if kfs is None else [[a + b for a, b in zip(Fa, Fb)] for Fa, Fb in zip(kfs, kf)]
    for F, Fo in zip(la.H_facs.kfacs, kfs):
        for H, Ho in zip(F, Fo):
            assert rel_fro(H.cpu(), Ho) < FACTOR_TOL
    _, ls = ko.decompose(kfs)
    h = 1.0 / float(la.sigma_noise) ** 2 / float(la.temperature)
    ld_ref = ko.kron_logdet(ko.scale_eigenvalues(ls, h), la.prior_precision.detach().cpu().double())
    assert torch.allclose(la.log_det_posterior_precision.detach().cpu().double(), ld_ref, rtol=1e-5)
