"""The reference's own test cases for serialization (tests/test_serialization.py:98-336), frozen parameter subsets
(tests/test_subset_params.py:83-140) and subnetwork posteriors (tests/test_subnetlaplace.py:673-922), driven with
``backend=B200GGN`` / ``B200EF`` through the UNMODIFIED reference classes -- same models, loaders, assertions.  Kernels: the
CPU emulation."""
from collections import OrderedDict
from itertools import product

import pytest
import torch
from torch import nn
from torch.nn.utils import parameters_to_vector
from torch.utils.data import DataLoader, TensorDataset

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
    from laplace import (DiagLaplace, DiagLLLaplace, DiagSubnetLaplace, FullLaplace, FullLLLaplace, FullSubnetLaplace,
                         KronLaplace, KronLLLaplace, Laplace, SubnetLaplace)
    from laplace.utils import (LargestMagnitudeSubnetMask, LastLayerSubnetMask, ModuleNameSubnetMask, ParamNameSubnetMask,
                               RandomSubnetMask)

    from laplace_b200 import B200EF, B200GGN

    ALL_W = [FullLaplace, KronLaplace, DiagLaplace]
    LLLA = [FullLLLaplace, KronLLLaplace, DiagLLLaplace]
    SUBNET = [DiagSubnetLaplace, FullSubnetLaplace]
    MASKS = [RandomSubnetMask, LargestMagnitudeSubnetMask, ParamNameSubnetMask, ModuleNameSubnetMask, LastLayerSubnetMask]
    BACKENDS = [B200GGN, B200EF]
else:                                            # pragma: no cover
    ALL_W = LLLA = SUBNET = MASKS = BACKENDS = []


@pytest.fixture(autouse=True)
def _kernels(cpu_kernels):
    torch.manual_seed(240)
    yield


def _model(hidden=20):
    m = nn.Sequential(nn.Linear(3, hidden), nn.Linear(hidden, 2)).to(D)
    m.output_size = 2
    m.n_layers = len(list(m.parameters()))
    m.n_params = len(parameters_to_vector(m.parameters()))
    return m


@pytest.fixture
def model():
    return _model()


@pytest.fixture
def reg_loader():
    return DataLoader(TensorDataset(torch.randn(10, 3, dtype=D), torch.randn(10, 2, dtype=D)), batch_size=3)


@pytest.fixture
def class_loader():
    return DataLoader(TensorDataset(torch.randn(10, 3, dtype=D), torch.randint(2, (10,))), batch_size=3)


def _fitted(laplace, model, loader, lik="regression", **kw):
    la = laplace(model, lik, backend=B200GGN, **kw)
    la.fit(loader)
    la.optimize_prior_precision(n_steps=10)        # (the reference runs the default 100 steps; the count is immaterial here)
    la.sigma_noise = 1231
    return la


# ------------------------------------------------------------------------------------------ serialization
@pytest.mark.parametrize("laplace", ALL_W + LLLA)
def test_serialize(laplace, model, reg_loader, tmp_path):
    """:98-115 and :157-169 -- ``state_dict`` through ``torch.save`` / ``torch.load``: same predictive; only plain containers
    and tensors in the file (it loads with ``weights_only=True``)."""
    la = _fitted(laplace, model, reg_loader)
    fn = tmp_path / "state_dict.bin"
    torch.save(la.state_dict(), fn)
    state = torch.load(fn, weights_only=True)
    for val in state.values():
        if val is not None:
            assert isinstance(val, (list, tuple, int, float, str, bool, torch.Tensor))
    la2 = laplace(model, "regression", backend=B200GGN)
    la2.load_state_dict(state)
    assert la.sigma_noise == la2.sigma_noise
    X, _ = next(iter(reg_loader))
    (f_mean, f_var), (f_mean2, f_var2) = la(X), la2(X)
    assert torch.allclose(f_mean, f_mean2) and torch.allclose(f_var, f_var2)


@pytest.mark.parametrize("laplace", ALL_W)
def test_serialize_override(laplace, model, reg_loader, tmp_path):
    """:135-153 -- continual learning: a second fit on a loaded posterior doubles the curvature."""
    la = _fitted(laplace, model, reg_loader)
    H_orig = la.H_facs.to_matrix() if laplace == KronLaplace else la.H
    torch.save(la.state_dict(), tmp_path / "sd.bin")
    la2 = laplace(model, "regression", backend=B200GGN)
    la2.load_state_dict(torch.load(tmp_path / "sd.bin"))
    la2.fit(reg_loader, override=False)
    H_new = la2.H_facs.to_matrix() if laplace == KronLaplace else la2.H
    assert torch.allclose(2 * H_orig, H_new)      # (KronLaplace.fit rescales the old factors so that the block matrix adds up)


@pytest.mark.parametrize("laplace", SUBNET)
def test_serialize_subnetlaplace(laplace, model, reg_loader, tmp_path):
    """:189-207."""
    idx = torch.LongTensor([1, 10, 104, 44])
    la = _fitted(laplace, model, reg_loader, subnetwork_indices=idx)
    torch.save(la.state_dict(), tmp_path / "sd.bin")
    la2 = laplace(model, "regression", subnetwork_indices=idx, backend=B200GGN)
    la2.load_state_dict(torch.load(tmp_path / "sd.bin"))
    X, _ = next(iter(reg_loader))
    (f_mean, f_var), (f_mean2, f_var2) = la(X), la2(X)
    assert la.sigma_noise == la2.sigma_noise and torch.allclose(f_mean, f_mean2) and torch.allclose(f_var, f_var2)


def test_serialize_mismatches_fail(model, reg_loader, tmp_path):
    """:210-291 -- a state dict only loads into a posterior of the same model, structure, subset and likelihood."""
    fn = tmp_path / "sd.bin"
    for laplace in ALL_W:
        torch.save(_fitted(laplace, model, reg_loader).state_dict(), fn)
        with pytest.raises(ValueError):
            laplace(_model(25), "regression", backend=B200GGN).load_state_dict(torch.load(fn))
    for laplace in ALL_W + LLLA:
        torch.save(_fitted(laplace, model, reg_loader).state_dict(), fn)
        with pytest.raises(ValueError):
            laplace(model, "classification", backend=B200GGN).load_state_dict(torch.load(fn))
    la = Laplace(model, "regression", subset_of_weights="all", hessian_structure="kron", backend=B200GGN)
    la.fit(reg_loader)
    torch.save(la.state_dict(), fn)
    with pytest.raises(ValueError):
        Laplace(model, "regression", subset_of_weights="all", hessian_structure="diag", backend=B200GGN).load_state_dict(torch.load(fn))
    la = Laplace(model, "regression", subset_of_weights="last_layer", hessian_structure="diag", backend=B200GGN)
    la.fit(reg_loader)
    torch.save(la.state_dict(), fn)
    with pytest.raises(ValueError):
        Laplace(model, "regression", subset_of_weights="all", hessian_structure="diag", backend=B200GGN).load_state_dict(torch.load(fn))
    model3 = nn.Sequential(OrderedDict([("fc1", nn.Linear(3, 20)), ("clf", nn.Linear(20, 2))])).to(D)
    for laplace in LLLA:
        torch.save(_fitted(laplace, model, reg_loader, last_layer_name="1").state_dict(), fn)
        with pytest.raises(ValueError):
            laplace(model3, "classification", last_layer_name="clf", backend=B200GGN).load_state_dict(torch.load(fn))


@pytest.mark.parametrize("laplace", ALL_W + SUBNET)
def test_whole_object_pickles(laplace, model, reg_loader, tmp_path):
    """:294-336 (``test_map_location``) -- ``torch.save(la)`` pickles the posterior WITH its backend and the model's forward
    hooks; the loaded object predicts the same and can be fitted further (the backend re-registers its hooks)."""
    kw = dict(subnetwork_indices=torch.LongTensor([1, 10, 104, 44])) if issubclass(laplace, SubnetLaplace) else {}
    la = _fitted(laplace, model, reg_loader, **kw)
    assert la._device.type == "cpu"
    fn = tmp_path / "la.pt"
    torch.save(la, fn)
    la2 = torch.load(fn, map_location="cpu", weights_only=False)
    assert la2._device.type == "cpu" and la.sigma_noise == la2.sigma_noise
    X, _ = next(iter(reg_loader))
    (f_mean, f_var), (f_mean2, f_var2) = la(X), la2(X)
    assert torch.allclose(f_mean, f_mean2) and torch.allclose(f_var, f_var2)
    n = la2.n_data
    la2.fit(reg_loader, override=False)                     # hooks of the unpickled backend capture the unpickled model
    assert la2.n_data == 2 * n
    assert la2.backend._layers is not None and all(L.mod in set(la2.model.modules()) for L in la2.backend._layers)


# ------------------------------------------------------------------------------------------ frozen parameter subsets
@pytest.fixture
def frozen_model():
    m = _model()
    for p in m.parameters():
        p.requires_grad = False
    m[0].weight.requires_grad = True
    return m


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("laplace", ALL_W)
def test_subset_of_parameters(laplace, backend, frozen_model, class_loader, reg_loader):
    """tests/test_subset_params.py:83-140 -- only ``model[0].weight`` trainable: mean / precision sizes, predictive, prior
    precision by marginal likelihood and by grid search."""
    n_params = frozen_model[0].weight.numel()
    for lik, loader in (("classification", class_loader), ("regression", reg_loader)):
        lap = laplace(frozen_model, lik, backend=backend)
        lap.fit(loader)
        assert lap.mean.shape == (n_params,)
    lap = laplace(frozen_model, "classification", backend=backend)
    lap.fit(class_loader)
    if laplace == DiagLaplace:
        assert lap.posterior_precision.shape == (n_params,)
    elif laplace == KronLaplace:
        assert lap.posterior_precision.to_matrix().shape == (n_params, n_params)
    lap(torch.randn(5, 3, dtype=D), pred_type="nn", link_approx="mc")
    lap.optimize_prior_precision(method="marglik", n_steps=10)
    # (the reference sweeps its default 100 grid values x 100 MC samples: minutes per case; the call path is the same)
    lap.optimize_prior_precision(method="gridsearch", val_loader=class_loader, pred_type="nn", link_approx="mc", grid_size=4,
                                 n_samples=10)


# ------------------------------------------------------------------------------------------ subnetwork posteriors
def _mask(mask_cls, model, loader):
    kw = dict(model=model)
    if mask_cls in (RandomSubnetMask, LargestMagnitudeSubnetMask):
        kw.update(n_params_subnet=32)
    elif mask_cls == ParamNameSubnetMask:
        kw.update(parameter_names=["0.weight", "1.bias"])
    elif mask_cls == ModuleNameSubnetMask:
        kw.update(module_names=["0"])
    mask = mask_cls(**kw)
    mask.select(loader)
    return mask


@pytest.mark.parametrize("hs", ["full", "diag"])
@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_full_subnet_mask_equals_all_weights(model, lik, hs, class_loader, reg_loader):
    """:673-710 -- a subnetwork containing every parameter has the all-weights curvature (rtol 1e-3)."""
    loader = class_loader if lik == "classification" else reg_loader
    idx = torch.arange(model.n_params)
    lap = Laplace(model, lik, subset_of_weights="subnetwork", subnetwork_indices=idx, hessian_structure=hs, backend=B200GGN)
    lap.fit(loader)
    assert lap.n_params_subnet == model.n_params
    full_lap = Laplace(model, lik, subset_of_weights="all", hessian_structure=hs, backend=B200GGN)
    full_lap.fit(loader)
    assert torch.allclose(full_lap.H, lap.H, rtol=1e-3)


@pytest.mark.parametrize("mask_cls,hs", list(product(MASKS, ["full", "diag"])))
def test_subnet_predictives_and_marglik(model, reg_loader, class_loader, mask_cls, hs):
    """:713-866."""
    mask = _mask(mask_cls, model, reg_loader)
    lap = Laplace(model, "regression", subset_of_weights="subnetwork", subnetwork_indices=mask.indices, hessian_structure=hs,
                  backend=B200GGN)
    assert isinstance(lap, SubnetLaplace)
    lap.fit(reg_loader)
    X, _ = reg_loader.dataset.tensors
    with pytest.raises(ValueError):
        lap(X, pred_type="linear")
    f_mu, f_var = lap(X, pred_type="glm")
    assert torch.allclose(f_mu, model(X)) and f_var.shape == (10, 2, 2)
    f_mu, f_var = lap(X, pred_type="nn", link_approx="mc")
    assert f_mu.shape == f_var.shape == (10, 2)
    lap.log_marginal_likelihood()
    mask = _mask(mask_cls, model, class_loader)
    lap = Laplace(model, "classification", subset_of_weights="subnetwork", subnetwork_indices=mask.indices,
                  hessian_structure=hs, backend=B200GGN)
    lap.fit(class_loader)
    X, _ = class_loader.dataset.tensors
    one = torch.tensor(10.0, dtype=D)
    for kw in (dict(pred_type="glm", link_approx="mc", n_samples=100), dict(pred_type="glm", link_approx="probit"),
               dict(pred_type="glm", link_approx="bridge"), dict(pred_type="glm", link_approx="bridge_norm"),
               dict(pred_type="nn", link_approx="mc", n_samples=100)):
        f_pred = lap(X, **kw)
        assert f_pred.shape == (10, 2) and torch.allclose(f_pred.sum(), one), kw
    lap.log_marginal_likelihood()


@pytest.mark.parametrize("lik,hs", list(product(["classification", "regression"], ["full", "diag"])))
def test_subnet_sample(model, lik, hs, class_loader, reg_loader):
    """:869-918 -- only the subnetwork's coordinates vary across samples; the generator is honoured."""
    loader = class_loader if lik == "classification" else reg_loader
    mask = RandomSubnetMask(model=model, n_params_subnet=10)
    mask.select()
    lap = Laplace(model, lik, subset_of_weights="subnetwork", subnetwork_indices=mask.indices, hessian_structure=hs,
                  backend=B200GGN)
    lap.fit(loader)
    n = 20
    gen = torch.Generator()
    gen.manual_seed(123)
    state = gen.get_state()
    for g in (gen, None):
        s1, s2 = lap.sample(n_samples=n, generator=g), lap.sample(n_samples=n, generator=g)
        assert not (s1 == s2).all()
    assert s1.shape == (n, model.n_params)
    gen.set_state(state)
    s1 = lap.sample(n_samples=n, generator=gen)
    gen.set_state(state)
    s2 = lap.sample(n_samples=n, generator=gen)
    assert (s1 == s2).all()
    params = parameters_to_vector(model.parameters())
    fixed = torch.ones(model.n_params, dtype=bool)
    fixed[mask.indices] = False
    assert model.n_params - lap.n_params_subnet == int(fixed.sum())
    assert (s1[:, fixed] == params[fixed].repeat(n, 1)).all()
    assert not (s1[:, ~fixed] == params[~fixed].repeat(n, 1)).all()
