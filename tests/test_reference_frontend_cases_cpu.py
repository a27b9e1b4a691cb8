"""The reference's own front-end test cases (tests/test_baselaplace.py:337-934) driven with ``backend=B200GGN`` /
``B200EF`` through the UNMODIFIED ``FullLaplace`` / ``KronLaplace`` / ``DiagLaplace`` classes -- same model
(``Linear(3, 20) -> Linear(20, 2)``), same loaders (10 samples, batch size 3), float64, same assertions and tolerances.
Kernels: the CPU emulation.  Everything the reference does with a fitted posterior must keep working when the curvature and
the Kronecker algebra come from this package: marginal likelihood, sampling, GLM / NN predictives with every link
approximation, joint covariances, predictive and functional samples, re-fitting, dict batches, reward modeling,
differentiable predictives, grid search, target-shape validation, dtypes."""
from copy import deepcopy
from math import prod, sqrt

import numpy as np
import pytest
import torch
from torch import nn
from torch.distributions import Categorical, MultivariateNormal, Normal
from torch.nn.utils import parameters_to_vector
from torch.utils.data import DataLoader, TensorDataset

from oracle import curvature_oracle as co
from oracle import ref_shim

D = torch.float64


def _usable():
    if not ref_shim.reference_available():
        return False
    try:
        import laplace  # noqa: F401
    except ImportError:
        return False
    from laplace_b200.interface import HAVE_REFERENCE

    return HAVE_REFERENCE


pytestmark = pytest.mark.skipif(not _usable(), reason="reference package not importable")

if _usable():
    from laplace.baselaplace import DiagLaplace, FullLaplace, KronLaplace
    from laplace.utils import KronDecomposed

    from laplace_b200 import B200EF, B200GGN

    FLAVORS = [FullLaplace, KronLaplace, DiagLaplace]
else:                                            # pragma: no cover
    FLAVORS = []
    B200EF = B200GGN = None


@pytest.fixture(autouse=True)
def _kernels(cpu_kernels):
    torch.manual_seed(240)
    yield


@pytest.fixture
def model():
    m = nn.Sequential(nn.Linear(3, 20), nn.Linear(20, 2)).to(D)
    m.output_size = 2
    m.n_layers = len(list(m.parameters()))
    m.n_params = len(parameters_to_vector(m.parameters()))
    return m


@pytest.fixture
def class_loader():
    return DataLoader(TensorDataset(torch.randn(10, 3, dtype=D), torch.randint(2, (10,))), batch_size=3)


@pytest.fixture
def reg_loader():
    return DataLoader(TensorDataset(torch.randn(10, 3, dtype=D), torch.randn(10, 2, dtype=D)), batch_size=3)


def make(laplace, model, lik, backend=None, **kw):
    return laplace(model, lik, backend=backend or B200GGN, **kw)


@pytest.mark.parametrize("lh", ["classification", "regression"])
@pytest.mark.parametrize("laplace", FLAVORS)
def test_laplace_functionality(laplace, lh, model, reg_loader, class_loader):
    """:337-410 -- log likelihood, log marginal likelihood from its definition, sampling mean, functional variance against
    the dense ``J Sigma J^T``."""
    loader, sigma_noise = (class_loader, 1.0) if lh == "classification" else (reg_loader, 0.3)
    lap = make(laplace, model, lh, sigma_noise=sigma_noise, prior_precision=0.7)
    lap.fit(loader)
    assert lap.n_data == len(loader.dataset) and lap.n_outputs == model.output_size
    X, y = loader.dataset.tensors
    f = model(X)
    if lh == "classification":
        log_lik_true = Categorical(logits=f).log_prob(y).sum()
        assert torch.allclose(lap.log_likelihood, log_lik_true)
    else:
        assert torch.allclose(lap.log_likelihood, Normal(loc=f, scale=sigma_noise).log_prob(y).sum())
        lap.sigma_noise = 0.72
        log_lik_true = Normal(loc=f, scale=0.72).log_prob(y).sum()
        assert torch.allclose(lap.log_likelihood, log_lik_true)
    theta = parameters_to_vector(model.parameters()).detach()
    assert torch.allclose(theta, lap.mean)
    prior_prec = torch.diag(lap.prior_precision_diag)
    lml = log_lik_true - 1 / 2 * theta @ prior_prec @ theta
    log_det_post = lap.posterior_precision.log().sum() if laplace == DiagLaplace else lap.posterior_precision.logdet()
    lml = lml + 1 / 2 * (prior_prec.logdet() - log_det_post)
    assert torch.allclose(lml, lap.log_marginal_likelihood())
    torch.manual_seed(61)
    assert lap.sample(n_samples=1).shape == (1, len(theta))
    samples = lap.sample(n_samples=200000)
    assert samples.shape == (200000, len(theta))
    assert torch.allclose(samples.mean(dim=0), lap.mean, atol=2e-2)          # the reference draws 1e6 and asserts 1e-2
    if laplace == FullLaplace:
        Sigma = lap.posterior_covariance
    elif laplace == KronLaplace:
        Sigma = lap.posterior_precision.to_matrix(exponent=-1)
    else:
        Sigma = torch.diag(lap.posterior_variance)
    Js, _ = co.jacobians(model, X)
    true_f_var = torch.einsum("mkp,pq,mcq->mkc", Js, Sigma, Js)
    assert torch.allclose(true_f_var, lap.functional_variance(Js), rtol=1e-4)
    # and with the Jacobians the backend itself hands to the predictive (structured route for Kron)
    Jb, _ = lap.backend.jacobians(X)
    assert torch.allclose(true_f_var, lap.functional_variance(Jb), rtol=1e-4, atol=1e-4 * float(true_f_var.abs().max()))


@pytest.mark.parametrize("laplace", FLAVORS)
def test_overriding_and_online_fit(laplace, model, reg_loader):
    """:413-452."""
    lap = make(laplace, model, "regression", sigma_noise=0.3, prior_precision=0.7)
    lap.fit(reg_loader)
    dense = (lambda P: P.to_matrix()) if laplace == KronLaplace else (lambda P: P.clone())
    P, m, marglik = dense(lap.posterior_precision), lap.mean.clone(), lap.log_marginal_likelihood().detach().clone()
    lap.fit(reg_loader, override=True)
    assert torch.allclose(lap.mean, m) and torch.allclose(dense(lap.posterior_precision), P)
    assert torch.allclose(marglik, lap.log_marginal_likelihood()) and lap.n_data == len(reg_loader.dataset)
    H1, loss, n_data = dense(lap.H), deepcopy(lap.loss.item()), deepcopy(lap.n_data)
    lap.fit(reg_loader, override=False)
    lap.fit(reg_loader, override=False)
    assert torch.allclose(3 * torch.tensor(loss, dtype=D), lap.loss) and 3 * n_data == lap.n_data
    assert torch.allclose(dense(lap.H), 3 * H1, rtol=1e-4, atol=1e-6 * float(H1.abs().max()))


def test_log_prob_full_and_kron(model, class_loader):
    """:455-478."""
    for laplace, pp in ((FullLaplace, 0.7), (KronLaplace, 0.24)):
        lap = make(laplace, model, "classification", prior_precision=pp)
        theta = torch.randn_like(parameters_to_vector(model.parameters()))
        prior = Normal(loc=lap.mean if laplace == KronLaplace else torch.zeros_like(theta), scale=sqrt(1 / pp))
        assert torch.allclose(lap.log_prob(theta), prior.log_prob(theta).sum())
        lap.fit(class_loader)
        P = lap.posterior_precision.to_matrix() if laplace == KronLaplace else lap.posterior_precision
        post = MultivariateNormal(loc=lap.mean, precision_matrix=P)
        assert torch.allclose(lap.log_prob(theta), post.log_prob(theta), rtol=1e-5)


@pytest.mark.parametrize("laplace", FLAVORS)
def test_regression_predictive(laplace, model, reg_loader):
    """:480-515 and :818-846 -- GLM, NN and joint predictives, ``diagonal_output``."""
    lap = make(laplace, model, "regression", sigma_noise=0.3, prior_precision=0.7)
    lap.fit(reg_loader)
    X, y = reg_loader.dataset.tensors
    f = model(X)
    with pytest.raises(ValueError):
        lap(X, pred_type="linear")
    f_mu, f_var = lap(X, pred_type="glm")
    assert torch.allclose(f_mu, f) and f_var.shape == (len(X), 2, 2)
    f_mu_nn, f_var_nn = lap(X, pred_type="nn", link_approx="mc")
    assert f_mu_nn.shape == f_var_nn.shape == (len(X), 2)
    f_mu_joint, f_cov_joint = lap(X, pred_type="glm", joint=True)
    assert f_mu_joint.shape == (prod(f_mu.shape),) and f_cov_joint.shape == (f_mu_joint.shape[0],) * 2
    b, k = y.shape
    f_var_joint = torch.einsum("bkbl->bkl", f_cov_joint.reshape(b, k, b, k))
    assert torch.allclose(f_var_joint, f_var, rtol=1e-4, atol=1e-4 * float(f_var.abs().max()))
    assert lap(X, pred_type="glm", joint=True, diagonal_output=True)[1].shape == (b * k, b * k)
    assert lap(X, pred_type="glm", joint=False, diagonal_output=True)[1].shape == (b, k)
    assert lap(X, pred_type="glm", joint=False, diagonal_output=False)[1].shape == (b, k, k)


@pytest.mark.parametrize("laplace", FLAVORS)
def test_classification_predictive(laplace, model, class_loader):
    """:518-556 -- every link approximation returns probabilities."""
    lap = make(laplace, model, "classification", prior_precision=0.7)
    lap.fit(class_loader)
    X, _ = class_loader.dataset.tensors
    f = torch.softmax(model(X), dim=-1)
    with pytest.raises(ValueError):
        lap(X, pred_type="linear")
    one = torch.tensor(len(f), dtype=D)
    for kw in (dict(pred_type="glm", link_approx="mc", n_samples=100), dict(pred_type="glm", link_approx="probit"),
               dict(pred_type="glm", link_approx="bridge"), dict(pred_type="glm", link_approx="bridge_norm"),
               dict(pred_type="nn", link_approx="mc", n_samples=100)):
        f_pred = lap(X, **kw)
        assert f_pred.shape == f.shape and torch.allclose(f_pred.sum(), one), kw


@pytest.mark.parametrize("laplace", FLAVORS)
def test_predictive_and_functional_samples(laplace, model, reg_loader, class_loader):
    """:559-633."""
    lap = make(laplace, model, "regression", sigma_noise=0.3, prior_precision=0.7)
    lap.fit(reg_loader)
    X, _ = reg_loader.dataset.tensors
    for pt in ("glm", "nn"):
        assert lap.predictive_samples(X, pred_type=pt, n_samples=100).shape == (100, len(X), 2)
    gen = torch.Generator()
    reg = {pt: lap.functional_samples(X, pred_type=pt, n_samples=100, generator=gen.manual_seed(123)) for pt in ("glm", "nn")}
    lap.likelihood = "classification"               # the samples do not depend on the likelihood
    for pt in ("glm", "nn"):
        again = lap.functional_samples(X, pred_type=pt, n_samples=100, generator=gen.manual_seed(123))
        assert again.shape == (100, len(X), 2) and torch.allclose(again, reg[pt])
    lapc = make(laplace, model, "classification", prior_precision=0.7)
    lapc.fit(class_loader)
    Xc, _ = class_loader.dataset.tensors
    for pt in ("glm", "nn"):
        s = lapc.predictive_samples(Xc, pred_type=pt, n_samples=100)
        assert s.shape == (100, len(Xc), 2) and np.allclose(s.sum().item(), len(Xc) * 100)


@pytest.mark.parametrize("laplace", [KronLaplace, DiagLaplace] if FLAVORS else [])
def test_reward_modeling(laplace):
    """:636-657 -- pairs ``(batch, 2, dim)`` at fit time, single inputs at test time."""
    class RewardModel(nn.Module):
        def __init__(self):
            super().__init__()
            self.net = nn.Sequential(nn.Linear(3, 100), nn.ReLU(), nn.Linear(100, 1))

        def forward(self, x):
            if len(x.shape) == 3:
                b, _, d = x.shape
                return self.net(x.reshape(-1, d)).reshape(b, 2)
            return self.net(x)

    rm = RewardModel().to(D)
    loader = DataLoader(TensorDataset(torch.randn(10, 2, 3, dtype=D), torch.randint(2, (10,))), batch_size=3)
    Xt = torch.randn(10, 3, dtype=D)
    lap = make(laplace, rm, "reward_modeling")
    lap.fit(loader)
    with pytest.raises(ValueError):
        lap(Xt, pred_type="linear")
    f_mu, f_var = lap(Xt, pred_type="glm")
    assert torch.allclose(f_mu, rm(Xt)) and f_var.shape == (10, 1, 1)
    f_mu, f_var = lap(Xt, pred_type="nn", link_approx="mc")
    assert f_mu.shape == f_var.shape == (10, 1)


@pytest.mark.parametrize("lik", ["classification", "regression", "reward_modeling"])
@pytest.mark.parametrize("backend_name", ["B200GGN", "B200EF"])
@pytest.mark.parametrize("laplace", [KronLaplace, DiagLaplace] if FLAVORS else [])
def test_dict_data(laplace, backend_name, lik):
    """:660-729 -- mapping batches under custom keys; the default key must fail loudly."""
    backend = {"B200GGN": B200GGN, "B200EF": B200EF}[backend_name]

    class CustomModel(nn.Module):
        def __init__(self):
            super().__init__()
            self.net = nn.Sequential(nn.Linear(5, 100), nn.ReLU(), nn.Linear(100, 2))

        def forward(self, data):
            x = data["test_input_key"] if isinstance(data, dict) else data
            return self.net(x)

    cm = CustomModel().to(D)
    n = 10
    Xs = torch.randn(n, 5, dtype=D)
    ys = torch.randn(n, 2, dtype=D) if lik == "regression" else torch.randint(2, (n,))

    class Loader(list):
        dataset = range(n)

    loader = Loader({"test_input_key": Xs[i:i + 3], "test_label_key": ys[i:i + 3]} for i in range(0, n, 3))
    with pytest.raises(KeyError):
        laplace(cm, lik, backend=backend).fit(loader)
    lap = laplace(cm, lik, backend=backend, dict_key_x="test_input_key", dict_key_y="test_label_key")
    lap.fit(loader)
    test_data = loader[0]
    f = cm(test_data)
    if lik == "classification":
        for kw in (dict(pred_type="glm"), dict(pred_type="nn", link_approx="mc")):
            f_pred = lap(test_data, **kw)
            assert f_pred.shape == f.shape and torch.allclose(f_pred.sum(), torch.tensor(len(f_pred), dtype=D))
    else:
        f_pred, f_var = lap(test_data, pred_type="glm")
        assert torch.allclose(f_pred, f) and f_var.shape == (len(f), 2, 2)
        f_pred, f_var = lap(test_data, pred_type="nn", link_approx="mc")
        assert f_pred.shape == f.shape == f_var.shape


@pytest.mark.parametrize("mode", ["glm", "glm_joint", "glm_mc", "nn"])
@pytest.mark.parametrize("backend_name", ["B200GGN", "B200EF"])
@pytest.mark.parametrize("laplace", FLAVORS)
def test_backprop(laplace, backend_name, mode, model, reg_loader):
    """:732-812 -- ``enable_backprop=True``: predictive mean and variance are differentiable w.r.t. the input."""
    backend = {"B200GGN": B200GGN, "B200EF": B200EF}[backend_name]
    X, _ = reg_loader.dataset.tensors
    X.requires_grad = True
    lap = laplace(model, "regression", enable_backprop=True, backend=backend)
    lap.fit(reg_loader)
    kw = {"glm": dict(pred_type="glm"), "glm_joint": dict(pred_type="glm", joint=True),
          "glm_mc": dict(pred_type="glm", link_approx="mc"), "nn": dict(pred_type="nn", link_approx="mc", n_samples=10)}[mode]
    f_mu, f_var = lap(X, **kw)
    assert torch.autograd.grad(f_mu.sum(), X, retain_graph=True)[0].shape == X.shape
    assert torch.autograd.grad(f_var.sum(), X)[0].shape == X.shape


@pytest.mark.parametrize("prior_prec_type", ["scalar", "layerwise", "diag"])
@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_gridsearch(model, lik, prior_prec_type, reg_loader, class_loader):
    """:858-879 -- runs for every prior-precision shape."""
    loader = reg_loader if lik == "regression" else class_loader
    pp = {"scalar": 1.0, "layerwise": torch.ones(model.n_layers, dtype=D), "diag": torch.ones(model.n_params, dtype=D)}[prior_prec_type]
    lap = make(DiagLaplace, model, lik, prior_precision=pp)
    lap.fit(loader)
    lap.optimize_prior_precision(method="gridsearch", val_loader=loader, n_steps=10)      # should not raise


@pytest.mark.parametrize("laplace", FLAVORS)
def test_parametric_fit_y_shape(laplace):
    """:882-890 -- a flat target against an ``(N, 1)`` output must raise, not broadcast."""
    torch.manual_seed(9999)
    m1 = nn.Sequential(nn.Linear(3, 20), nn.Linear(20, 1)).to(D)
    X = torch.randn(10, 3, dtype=D)
    ok = DataLoader(TensorDataset(X, torch.randn(10, 1, dtype=D)), batch_size=3)
    flat = DataLoader(TensorDataset(X, torch.randn(10, dtype=D)), batch_size=3)
    make(laplace, m1, "regression").fit(ok)
    with pytest.raises(ValueError):
        make(laplace, m1, "regression").fit(flat)


@pytest.mark.parametrize("lik", ["classification", "regression"])
@pytest.mark.parametrize("dtype", [torch.float, torch.double])
@pytest.mark.parametrize("backend_name", ["B200GGN", "B200EF"])
@pytest.mark.parametrize("laplace", FLAVORS)
def test_dtype(laplace, backend_name, dtype, lik):
    """:893-934 -- everything the posterior returns is in the model's dtype (fp16 models are outside the kernels' contract)."""
    backend = {"B200GGN": B200GGN, "B200EF": B200EF}[backend_name]
    X, Y = torch.randn(10, 3, dtype=dtype), torch.randn(10, 3, dtype=dtype)
    if lik == "classification":
        Y = torch.randint(3, (10,))
    loader = DataLoader(TensorDataset(X, Y), batch_size=10)
    m = nn.Linear(3, 3, dtype=dtype)
    la = laplace(m, lik, backend=backend)
    la.fit(loader)
    if isinstance(la.H, torch.Tensor):
        assert la.H.dtype == dtype
    else:
        assert isinstance(la.H, KronDecomposed)
        assert la.H.eigenvalues[0][0].dtype == dtype and la.H.eigenvectors[0][0].dtype == dtype
    assert la.log_marginal_likelihood().dtype == dtype
    out = la(X, pred_type="glm")
    for t in (out if isinstance(out, tuple) else (out,)):
        assert t.dtype == dtype
    out = la(X, pred_type="nn", link_approx="mc", n_samples=3)      # (the reference's call raises and is ignored by its try/except)
    for t in (out if isinstance(out, tuple) else (out,)):
        assert t.dtype == dtype
