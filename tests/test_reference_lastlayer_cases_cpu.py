"""The reference's own last-layer test cases (tests/test_lllaplace.py:361-706) driven with ``backend=B200GGN`` / ``B200EF``
through the UNMODIFIED ``FullLLLaplace`` / ``KronLLLaplace`` / ``DiagLLLaplace`` -- same models (``Linear(3, 20) ->
Linear(20, 2)``; the ``relu -> mean(1)`` model on ``(10, 6, 3)`` inputs for the feature-reduction cases), float64, same
assertions and tolerances.  Kernels: the CPU emulation."""
from itertools import product

import numpy as np
import pytest
import torch
from torch import nn
from torch.distributions import Categorical, Normal
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
    from laplace.lllaplace import DiagLLLaplace, FullLLLaplace, KronLLLaplace
    from laplace.utils import FeatureExtractor
    from laplace.utils.feature_extractor import FeatureReduction

    from laplace_b200 import B200EF, B200GGN

    FLAVORS = [FullLLLaplace, KronLLLaplace, DiagLLLaplace]
    REDUCTIONS = [f.value for f in FeatureReduction] + [None]
else:                                            # pragma: no cover
    FLAVORS, REDUCTIONS = [], []


@pytest.fixture(autouse=True)
def _kernels(cpu_kernels):
    torch.manual_seed(240)
    yield


@pytest.fixture
def model():
    m = nn.Sequential(nn.Linear(3, 20), nn.Linear(20, 2)).to(D)
    m.output_size = 2
    return m


class ReductionModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1, self.fc2, self.output_size = nn.Linear(3, 20), nn.Linear(20, 2), 2

    def forward(self, x):
        return self.fc2(nn.functional.relu(self.fc1(x)).mean(1))


def loader(lik, multidim=False):
    X = torch.randn(10, 6, 3, dtype=D) if multidim else torch.randn(10, 3, dtype=D)
    y = torch.randint(2, (10,)) if lik == "classification" else torch.randn(10, 2, dtype=D)
    return DataLoader(TensorDataset(X, y), batch_size=3)


@pytest.mark.parametrize("reduction", REDUCTIONS)
@pytest.mark.parametrize("multidim", [False, True])
@pytest.mark.parametrize("laplace,lh", list(product(FLAVORS, ["classification", "regression"])))
def test_laplace_functionality(laplace, lh, multidim, reduction, model):
    """:361-468 -- likelihood, marginal likelihood from its definition, sampling mean, last-layer Jacobians against the naive
    ones, functional variance against the dense ``J Sigma J^T``."""
    sigma_noise = 1.0 if lh == "classification" else 0.3
    dl = loader(lh, multidim)
    last_layer_name = "1"
    if multidim:
        model, last_layer_name = ReductionModel().to(D), "fc2"
    lap = laplace(model, lh, sigma_noise=sigma_noise, prior_precision=0.7, feature_reduction=reduction, backend=B200GGN)
    lap.fit(dl)
    assert lap.n_data == len(dl.dataset) and lap.n_outputs == model.output_size
    X, y = dl.dataset.tensors
    f = model(X)
    assert f.shape == (10, 2)
    if lh == "classification":
        log_lik_true = Categorical(logits=f).log_prob(y).sum()
        assert torch.allclose(lap.log_likelihood, log_lik_true)
    else:
        assert torch.allclose(lap.log_likelihood, Normal(loc=f, scale=sigma_noise).log_prob(y).sum())
        lap.sigma_noise = 0.72
        log_lik_true = Normal(loc=f, scale=0.72).log_prob(y).sum()
        assert torch.allclose(lap.log_likelihood, log_lik_true)
    fe = FeatureExtractor(model, last_layer_name=last_layer_name, feature_reduction=reduction)
    theta = parameters_to_vector(fe.last_layer.parameters()).detach()
    assert torch.allclose(theta, lap.mean)
    prior_prec = torch.diag(lap.prior_precision_diag)
    assert prior_prec.shape == (len(theta), len(theta))
    lml = log_lik_true - 1 / 2 * theta @ prior_prec @ theta
    log_det_post = lap.posterior_precision.log().sum() if laplace == DiagLLLaplace else lap.posterior_precision.logdet()
    lml = lml + 1 / 2 * (prior_prec.logdet() - log_det_post)
    assert torch.allclose(lml, lap.log_marginal_likelihood())
    torch.manual_seed(61)
    assert lap.sample(n_samples=1).shape == (1, len(theta))
    samples = lap.sample(n_samples=200000)
    assert torch.allclose(samples.mean(dim=0), lap.mean, atol=2e-2)          # the reference draws 1e6 and asserts 1e-2
    if laplace == FullLLLaplace:
        Sigma = lap.posterior_covariance
    elif laplace == KronLLLaplace:
        Sigma = lap.posterior_precision.to_matrix(exponent=-1)
    else:
        Sigma = torch.diag(lap.posterior_variance)
    _, phi = fe.forward_with_features(X)
    Js, f_ll = co.jacobians(fe.last_layer, phi.detach())
    true_f_var = torch.einsum("mkp,pq,mcq->mkc", Js, Sigma, Js)
    comp_Js, comp_f = lap.backend.last_layer_jacobians(X)
    assert torch.allclose(Js, comp_Js) and torch.allclose(f_ll, comp_f)
    comp_f_var = lap.functional_variance(comp_Js)
    assert torch.allclose(true_f_var, comp_f_var, rtol=1e-4, atol=1e-6 * float(true_f_var.abs().max()))


@pytest.mark.parametrize("backend_name", ["B200GGN", "B200EF"])
@pytest.mark.parametrize("laplace", FLAVORS)
def test_predictives(laplace, backend_name, model):
    """:471-574 -- regression / classification predictives, every link approximation, predictive samples."""
    backend = {"B200GGN": B200GGN, "B200EF": B200EF}[backend_name]
    dl = loader("regression")
    lap = laplace(model, "regression", sigma_noise=0.3, prior_precision=0.7, backend=backend)
    lap.fit(dl)
    X, _ = dl.dataset.tensors
    with pytest.raises(ValueError):
        lap(X, pred_type="linear")
    f_mu, f_var = lap(X, pred_type="glm")
    assert torch.allclose(f_mu, model(X)) and f_var.shape == (10, 2, 2)
    f_mu, f_var = lap(X, pred_type="nn", link_approx="mc")
    assert f_mu.shape == f_var.shape == (10, 2)
    for pt in ("glm", "nn"):
        assert lap.predictive_samples(X, pred_type=pt, n_samples=100).shape == (100, 10, 2)
    dl = loader("classification")
    lap = laplace(model, "classification", prior_precision=0.7, backend=backend)
    lap.fit(dl)
    X, _ = dl.dataset.tensors
    one = torch.tensor(10.0, dtype=D)
    for kw in (dict(pred_type="glm", link_approx="mc", n_samples=100), dict(pred_type="glm", link_approx="probit"),
               dict(pred_type="glm", link_approx="bridge"), dict(pred_type="glm", link_approx="bridge_norm"),
               dict(pred_type="nn", link_approx="mc", n_samples=100)):
        f_pred = lap(X, **kw)
        assert f_pred.shape == (10, 2) and torch.allclose(f_pred.sum(), one), kw
    for pt in ("glm", "nn"):
        s = lap.predictive_samples(X, pred_type=pt, n_samples=100)
        assert s.shape == (100, 10, 2) and np.allclose(s.sum().item(), 1000)


@pytest.mark.parametrize("laplace", [FullLLLaplace, DiagLLLaplace] if FLAVORS else [])
def test_functional_variance_fast(laplace, model):
    """:578-605 -- the reference's own structured last-layer variance against the Jacobian route, both through our backend
    with ``enable_backprop=True``."""
    dl = loader("regression")
    X, y = dl.dataset.tensors
    X.requires_grad = True
    lap = laplace(model, "regression", enable_backprop=True, backend=B200GGN)
    lap.fit(dl)
    f_mu, f_var = lap.functional_variance_fast(X)
    assert f_mu.shape == f_var.shape == (10, 2)
    Js, f_naive = lap.backend.last_layer_jacobians(X)
    if laplace == DiagLLLaplace:
        naive = torch.einsum("ncp,p,ncp->nc", Js, lap.posterior_variance, Js)
    else:
        naive = torch.einsum("ncp,pq,ncq->nc", Js, lap.posterior_covariance, Js)
    assert torch.allclose(f_mu, f_naive) and torch.allclose(f_var, naive)


@pytest.mark.parametrize("mode", ["glm", "glm_joint", "glm_mc", "nn"])
@pytest.mark.parametrize("laplace", FLAVORS)
def test_backprop_and_output_shapes(laplace, mode, model):
    """:608-705."""
    dl = loader("regression")
    X, y = dl.dataset.tensors
    X.requires_grad = True
    lap = laplace(model, "regression", enable_backprop=True, backend=B200GGN)
    lap.fit(dl)
    kw = {"glm": dict(pred_type="glm"), "glm_joint": dict(pred_type="glm", joint=True),
          "glm_mc": dict(pred_type="glm", link_approx="mc"), "nn": dict(pred_type="nn", link_approx="mc", n_samples=10)}[mode]
    f_mu, f_var = lap(X, **kw)
    assert torch.autograd.grad(f_mu.sum(), X, retain_graph=True)[0].shape == X.shape
    assert torch.autograd.grad(f_var.sum(), X)[0].shape == X.shape
    if mode == "glm":
        b, k = y.shape
        plain = laplace(model, "regression", backend=B200GGN)
        plain.fit(dl)
        Xd = X.detach()
        assert plain(Xd, pred_type="glm", joint=True, diagonal_output=True)[1].shape == (b * k, b * k)
        assert plain(Xd, pred_type="glm", joint=False, diagonal_output=True)[1].shape == (b, k)
        assert plain(Xd, pred_type="glm", joint=False, diagonal_output=False)[1].shape == (b, k, k)
