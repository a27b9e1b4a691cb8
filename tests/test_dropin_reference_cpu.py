"""Drop-in check: the UNMODIFIED reference front end (``Laplace(...).fit`` / ``la(x, pred_type="glm")``)
driven with ``backend=B200GGN`` / ``B200EF``.  Needs the reference tree (build container only); kernels
are replaced by the CPU emulation, so this pins the *boundary wiring* (SURVEY 8(b)): construction through
``BaseLaplace.backend``, ``la.H += H_batch`` dispatch, ``decompose``, ``posterior_precision`` algebra,
``functional_variance``."""
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

from oracle import ref_shim
from tests.fixtures import load, rel_fro

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


@pytest.fixture
def laplace_mod():
    try:
        import laplace
    except ImportError:
        pytest.skip("reference package not importable (LPB_NO_REFERENCE=1 or stubs not installed)")
    from laplace_b200.interface import HAVE_REFERENCE

    if not HAVE_REFERENCE:
        pytest.skip("laplace_b200 was imported before the reference became importable")
    return laplace


@pytest.mark.parametrize("kind", ["mlp", "conv"])
@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_kron_laplace_fit_and_glm(golden, cpu_kernels, laplace_mod, kind, lik):
    from laplace.utils.matrix import Kron, KronDecomposed

    from laplace_b200 import B200GGN, B200Kron, B200KronDecomposed
    from oracle import curvature_oracle as co

    model, X, y, _ = load(golden, kind, lik, dtype=torch.float32)
    loader = DataLoader(TensorDataset(X, y), batch_size=4)
    la = laplace_mod.Laplace(model, lik, "all", "kron", backend=B200GGN, prior_precision=0.7)
    la.fit(loader)
    assert isinstance(la.H_facs, B200Kron) and isinstance(la.H, B200KronDecomposed)
    assert isinstance(la.posterior_precision, B200KronDecomposed)
    # reference algebra on the oracle factors
    kfs = None
    for i in range(0, len(X), 4):
        _, kf = co.kfac_factors(model.double(), lik, X[i:i + 4].double(), y[i:i + 4] if lik == "classification" else y[i:i + 4].double(), N=len(X))
        kfs = kf if kfs is None else [[a + b for a, b in zip(Fa, Fb)] for Fa, Fb in zip(kfs, kf)]
    model.float()
    for F, Fo in zip(la.H_facs.kfacs, kfs):
        for H, Ho in zip(F, Fo):
            assert rel_fro(H, Ho) < 1e-4
    ref = Kron([[H.float() for H in F] for F in kfs]).decompose() * 1.0 + la.prior_precision
    assert type(ref) is KronDecomposed
    f_mu, f_var = la._glm_predictive_distribution(X)
    Js, _ = la.backend.jacobians(X)
    f_var_ref = ref.inv_square_form(Js.clone())
    assert torch.allclose(f_var, f_var_ref, rtol=2e-3, atol=1e-6)
    assert torch.allclose(la.log_det_posterior_precision, ref.logdet(), rtol=1e-4)
    if lik == "classification":
        probs = la(X, pred_type="glm", link_approx="probit")
        assert torch.allclose(probs.sum(-1), torch.ones(len(X)), atol=1e-5)
    la.log_marginal_likelihood()
    sd = la.state_dict()
    assert all(isinstance(h, torch.Tensor) for F in sd["H"] for h in F)


@pytest.mark.parametrize("hs", ["full", "diag"])
@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_full_and_diag_laplace_match_golden(golden, cpu_kernels, laplace_mod, hs, lik):
    from laplace_b200 import B200GGN

    model, X, y, rec = load(golden, "mlp", lik)
    loader = DataLoader(TensorDataset(X, y), batch_size=4)
    la = laplace_mod.Laplace(model, lik, "all", hs, backend=B200GGN, prior_precision=0.7)
    la.fit(loader)
    f_mu, f_var = la._glm_predictive_distribution(X)
    assert torch.allclose(f_mu, rec[f"la_{hs}_f_mu"], atol=1e-6)
    assert torch.allclose(f_var, rec[f"la_{hs}_f_var"], rtol=1e-4, atol=1e-7)
    assert torch.allclose(la.log_marginal_likelihood(), rec[f"la_{hs}_logmarglik"], rtol=1e-5)


def test_last_layer_full_laplace(golden, cpu_kernels, laplace_mod):
    from laplace.curvature import GGNInterface

    from laplace_b200 import B200EF, B200GGN

    from laplace_b200 import predictive

    calls = {"full": 0, "diag": 0}
    orig_full, orig_diag = predictive.ll_full_variance, predictive.ll_diag_variance
    predictive.ll_full_variance = lambda *a, **k: (calls.__setitem__("full", calls["full"] + 1), orig_full(*a, **k))[1]
    predictive.ll_diag_variance = lambda *a, **k: (calls.__setitem__("diag", calls["diag"] + 1), orig_diag(*a, **k))[1]
    model, X, y, rec = load(golden, "mlp", "classification")
    loader = DataLoader(TensorDataset(X, y), batch_size=4)
    for be, ref_be in ((B200GGN, GGNInterface),):
        la = laplace_mod.Laplace(model, "classification", "last_layer", "full", backend=be, prior_precision=0.7)
        la.fit(loader)
        lr = laplace_mod.Laplace(model, "classification", "last_layer", "full", backend=ref_be, prior_precision=0.7)
        lr.fit(loader)
        assert rel_fro(la.H, lr.H) < 1e-5
        fm, fv = la._glm_predictive_distribution(X)
        fm_r, fv_r = lr._glm_predictive_distribution(X)
        assert torch.allclose(fv, fv_r, rtol=1e-4, atol=1e-8)
    for hs in ("diag", "kron"):
        la = laplace_mod.Laplace(model, "classification", "last_layer", hs, backend=B200GGN, prior_precision=0.7)
        la.fit(loader)
        _, fv = la._glm_predictive_distribution(X)
        if hs == "diag":
            lr = laplace_mod.Laplace(model, "classification", "last_layer", hs, backend=GGNInterface, prior_precision=0.7)
            lr.fit(loader)
            assert torch.allclose(fv, lr._glm_predictive_distribution(X)[1], rtol=1e-4, atol=1e-8)
    la = laplace_mod.Laplace(model, "classification", "last_layer", "full", backend=B200EF)
    la.fit(loader)
    predictive.ll_full_variance, predictive.ll_diag_variance = orig_full, orig_diag
    # the reference's own einsums (baselaplace.py:1683-1684, :2115) were routed into the structured kernels
    assert calls["full"] >= 1 and calls["diag"] >= 1, calls


def test_subnetwork_laplace(golden, cpu_kernels, laplace_mod):
    from laplace.curvature import GGNInterface

    from laplace_b200 import B200GGN

    model, X, y, _ = load(golden, "mlp", "classification")
    loader = DataLoader(TensorDataset(X, y), batch_size=5)
    idx = torch.tensor([0, 3, 17, 60, 61, 100, 121])
    la = laplace_mod.Laplace(model, "classification", "subnetwork", "full", subnetwork_indices=idx, backend=B200GGN)
    la.fit(loader)
    lr = laplace_mod.Laplace(model, "classification", "subnetwork", "full", subnetwork_indices=idx, backend=GGNInterface)
    lr.fit(loader)
    assert rel_fro(la.H, lr.H) < 1e-5


def test_kron_laplace_with_fused_conv_chains(cpu_kernels, laplace_mod, monkeypatch):
    """The large-batch code paths (fused conv -> frozen BN -> ReLU reverse chains, strided / implicit convolutions,
    im2col-free input factors, side-stream bookkeeping) behind the UNMODIFIED reference front end: ``Laplace(...).fit``
    accumulates the same Kronecker factors as the oracle and the GLM predictive runs."""
    from laplace_b200 import B200GGN, B200Kron, conv_engine
    from oracle import curvature_oracle as co

    monkeypatch.setattr(conv_engine, "ELEMENTWISE_MIN_BATCH", 0)
    torch.manual_seed(3)
    bn = torch.nn.BatchNorm2d(64)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 64, 3, 1, 1, bias=False), bn, torch.nn.ReLU(),
                                torch.nn.Conv2d(64, 64, 3, 2, 1, bias=False), torch.nn.ReLU(), torch.nn.Conv2d(64, 8, 3, 1, 1),
                                torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(), torch.nn.Linear(8, 3)).eval()
    bn.running_mean.normal_(0, 0.1), bn.running_var.uniform_(0.5, 1.5)
    for p in bn.parameters():
        p.requires_grad_(False)
    X, y = torch.randn(12, 3, 8, 8), torch.randint(3, (12,))
    la = laplace_mod.Laplace(model, "classification", "all", "kron", backend=B200GGN, prior_precision=1.3)
    la.fit(DataLoader(TensorDataset(X, y), batch_size=6))
    assert isinstance(la.H_facs, B200Kron) and la.backend._fused and la.backend.fuse_elementwise
    kfs = None
    md = model.double()
    for i in (0, 6):
        _, kf = co.kfac_factors(md, "classification", X[i:i + 6].double(), y[i:i + 6], N=12)
        kfs = kf if kfs is None else [[a + b for a, b in zip(Fa, Fb)] for Fa, Fb in zip(kfs, kf)]
    model.float()
    worst = max(rel_fro(H, Ho) for F, Fo in zip(la.H_facs.kfacs, kfs) for H, Ho in zip(F, Fo))
    assert worst < 1e-4, worst
    probs = la(X[:5], pred_type="glm", link_approx="probit")
    assert probs.shape == (5, 3) and torch.allclose(probs.sum(-1), torch.ones(5), atol=1e-5)


def _small_problem(seed=0):
    torch.manual_seed(seed)
    model = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
    X, y = torch.randn(64, 5), torch.randint(3, (64,))
    Xv, yv = torch.randn(40, 5), torch.randint(3, (40,))
    return model, X, y, Xv, yv


def test_prior_precision_gridsearch_on_cached_projections(cpu_kernels, laplace_mod):
    """the reference's grid search (baselaplace.py:516-561) re-derives Jacobians for each grid value; inside
    ``backend.cached_jacobians()`` -- and in ``laplace_b200.tuning`` -- every validation batch pays once.  Same winner,
    same loss curve as the fp64 restatement of the reference loop (running, never-reset metric included)."""
    from laplace_b200 import B200GGN
    from laplace_b200.tuning import gridsearch_prior_precision
    from oracle import curvature_oracle as co
    from oracle import kron_oracle as ko

    model, X, y, Xv, yv = _small_problem()
    la = laplace_mod.Laplace(model, "classification", "all", "kron", backend=B200GGN)
    la.fit(DataLoader(TensorDataset(X, y), batch_size=16))
    vl = DataLoader(TensorDataset(Xv, yv), batch_size=20)
    calls = {"n": 0}
    orig = la.backend._jacobians_impl
    la.backend._jacobians_impl = lambda x: (calls.__setitem__("n", calls["n"] + 1), orig(x))[1]
    G = 12
    with pytest.warns(UserWarning):
        la.optimize_prior_precision(pred_type="glm", method="gridsearch", val_loader=vl, grid_size=G)
    ref_pp, n_ref = la.prior_precision.clone(), calls["n"]
    calls["n"] = 0
    with la.backend.cached_jacobians(), pytest.warns(UserWarning):
        la.optimize_prior_precision(pred_type="glm", method="gridsearch", val_loader=vl, grid_size=G)
    assert torch.equal(la.prior_precision, ref_pp) and n_ref == 2 * G and calls["n"] == 2
    calls["n"] = 0
    best, losses = gridsearch_prior_precision(la, vl, grid_size=G, running_metric=True)
    assert calls["n"] == 2 and torch.allclose(best, ref_pp.squeeze())
    md = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3)).double()
    md.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    kfs = None
    for i in range(0, 64, 16):
        _, kf = co.kfac_factors(md, "classification", X[i:i + 16].double(), y[i:i + 16], N=64)
        kfs = kf if kfs is None else [[a + b for a, b in zip(Fa, Fb)] for Fa, Fb in zip(kfs, kf)]
    Qs, ls = ko.decompose(kfs)
    Jb, fb, yb = [], [], []
    for i in range(0, 40, 20):
        J, f = co.jacobians(md, Xv[i:i + 20].double())
        Jb.append(J), fb.append(f), yb.append(yv[i:i + 20])
    for running in (True, False):
        bo, lo = ko.gridsearch_prior_precision(Qs, ls, Jb, fb, yb, torch.logspace(-4, 4, G).double(), running_metric=running)
        b2, l2 = gridsearch_prior_precision(la, vl, grid_size=G, running_metric=running, set_result=False)
        assert torch.allclose(l2, lo, atol=1e-5) and torch.allclose(b2.double(), bo)


@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_marglik_training_with_b200_backend(cpu_kernels, laplace_mod, lik):
    """the reference's ``marglik_training`` (marglik_training.py:34-58, 277-301) takes the backend as a
    plain class argument; with ``backend=B200GGN`` it fits every ``marglik_frequency`` epochs through our kernels and
    differentiates the marginal likelihood w.r.t. prior precision AND ``sigma_noise`` through ``B200KronDecomposed``."""
    from laplace.marglik_training import marglik_training

    from laplace_b200 import B200GGN, B200Kron

    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
    X = torch.randn(64, 5)
    y = torch.randint(3, (64,)) if lik == "classification" else torch.randn(64, 3)
    la, model, margliks, losses = marglik_training(model, DataLoader(TensorDataset(X, y), batch_size=16), lik, "kron",
                                                   backend=B200GGN, n_epochs=4, marglik_frequency=2, n_hypersteps=3)
    assert isinstance(la.H_facs, B200Kron) and len(margliks) == 6 and all(torch.isfinite(torch.tensor(margliks)))
    assert margliks[-1] < margliks[0]                      # the hyper-steps decrease the negative log marginal likelihood
    if lik == "regression":
        assert float(la.sigma_noise) != 1.0                # the gradient w.r.t. sigma_noise reached the optimiser
