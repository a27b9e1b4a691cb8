"""The reference's own Jacobian and in-tree curvature-interface test cases (tests/test_jacobians.py:37-185,
tests/test_curv_backends_interface.py:71-365) with ``B200GGN`` / ``B200EF`` in place of ``CurvatureInterface`` /
``GGNInterface`` / ``EFInterface`` -- same models (200 x 3 inputs, ``Linear(3, 20) -> Linear(20, k)``; the seed-711 MLP and
"complex" conv model on 10 samples), same assertions and tolerances.  Where the reference compares against BackPACK / ASDL
(absent here) the comparison is against the fp64 oracle.  Kernels: the CPU emulation."""
import pytest
import torch
from torch import nn

from laplace_b200 import B200EF, B200GGN
from laplace_b200.posterior import LastLayerModel
from oracle import curvature_oracle as co
from tests.fixtures import load

D = torch.float64


@pytest.fixture
def X():
    torch.manual_seed(15)
    return torch.randn(200, 3, dtype=D)


def _net(k, bias=True, hidden=True):
    torch.manual_seed(3)
    m = nn.Sequential(nn.Linear(3, 20), nn.Linear(20, k)) if hidden else nn.Sequential(nn.Linear(3, k, bias=bias))
    m.output_size = k
    return m.to(D)


def test_linear_jacobians(cpu_kernels, X):
    """:37-48 -- the Jacobian of a bias-free linear map is its input."""
    m = _net(1, bias=False, hidden=False)
    Js, f = B200GGN(m, "classification").jacobians(X)
    true_Js = X.reshape(len(X), 1, -1)
    assert true_Js.shape == Js.shape and torch.allclose(true_Js, Js, atol=1e-5) and torch.allclose(f, m(X), atol=1e-5)


@pytest.mark.parametrize("k", [1, 2])
@pytest.mark.parametrize("Backend", [B200GGN, B200EF])
def test_jacobians_and_last_layer_jacobians(cpu_kernels, X, Backend, k):
    """:51-106 -- all-weights and last-layer Jacobians against naive per-output autograd (max abs error < 1e-6)."""
    m = _net(k)
    Js, f = Backend(m, "classification").jacobians(X)
    Jn, fn = co.jacobians(m, X)
    assert Js.shape == Jn.shape and torch.abs(Js - Jn).max() < 1e-6 and torch.allclose(f, fn) and torch.allclose(m(X), fn)
    fe = LastLayerModel(m)
    Jl, fl = Backend(fe, "classification", last_layer=True).last_layer_jacobians(X)
    _, phi = fe.forward_with_features(X)
    Jln, fln = co.jacobians(fe.last_layer, phi.detach())
    assert Jl.shape == Jln.shape and torch.abs(Jl - Jln).max() < 1e-6 and torch.allclose(fl, fln) and torch.allclose(m(X), fln)


@pytest.mark.parametrize("kind,lik", [("mlp", "classification"), ("conv", "classification"), ("mlp", "regression")])
def test_batchgrad(golden, cpu_kernels, kind, lik):
    """tests/test_curv_backends_interface.py:71-101 -- per-sample loss gradients."""
    model, Xb, y, _ = load(golden, kind, lik)
    Gs, loss = B200EF(model, lik).gradients(Xb, y)
    Jo, fo = co.jacobians(model, Xb)
    Go, lo = co.gradients(Jo, fo, y, lik)
    assert torch.allclose(Gs, Go, atol=1e-8) and torch.allclose(loss, lo)


@pytest.mark.parametrize("structure", ["diag", "full"])
@pytest.mark.parametrize("kind,lik,flavour", [("mlp", "classification", "ggn"), ("conv", "classification", "ggn"),
                                              ("mlp", "classification", "ef"), ("mlp", "regression", "ggn"),
                                              ("mlp", "regression", "ef")])
def test_two_batches_against_the_whole_batch(golden, cpu_kernels, structure, kind, lik, flavour):
    """:104-143, :166-204, :226-264, :286-324, :346-365 -- the curvature of two half batches, summed, against the full
    GGN / EF of all samples computed by a second implementation (there: BackPACK ``full``)."""
    torch.manual_seed(0)
    model, Xb, y, _ = load(golden, kind, lik)
    n_params = sum(p.numel() for p in model.parameters())
    backend = (B200GGN(model, lik, stochastic=False) if flavour == "ggn" else B200EF(model, lik))
    fn = getattr(backend, structure)
    loss, H = fn(Xb[:5], y[:5])
    loss2, H2 = fn(Xb[5:], y[5:])
    loss, H = loss + loss2, H + H2
    assert H.shape == ((n_params,) if structure == "diag" else (n_params, n_params))
    Jo, fo = co.jacobians(model, Xb)
    loss_f, H_ref = (co.ggn_full if flavour == "ggn" else co.ef_full)(Jo, fo, y, lik)
    assert torch.allclose(loss, loss_f)
    if structure == "diag":
        assert torch.allclose(H, H_ref.diag())
    else:
        assert torch.allclose(H, H_ref, atol=1e-6)          # the reference asserts atol 0.1 here (fp32 contraction: 1e-7 relative)


@pytest.mark.parametrize("structure", ["diag", "full"])
@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_stochastic_two_batches_against_the_exact_ggn(golden, cpu_kernels, structure, lik):
    """:146-163, :206-223, :266-283, :326-343 -- MC Fisher with 10 000 samples: diagonal within 1e-2 (classification) / 1e-1
    (regression), full within 1e-1."""
    torch.manual_seed(0)
    model, Xb, y, _ = load(golden, "mlp", lik)
    backend = B200GGN(model, lik, stochastic=True, num_samples=10000)
    fn = getattr(backend, structure)
    loss, H = fn(Xb[:5], y[:5])
    loss2, H2 = fn(Xb[5:], y[5:])
    loss, H = loss + loss2, H + H2
    Jo, fo = co.jacobians(model, Xb)
    loss_f, H_ggn = co.ggn_full(Jo, fo, y, lik)
    assert torch.allclose(loss, loss_f)
    if structure == "diag":
        assert torch.allclose(H, H_ggn.diag(), atol=0.01 if lik == "classification" else 0.1)
    else:
        assert torch.allclose(H, H_ggn, atol=0.1)
