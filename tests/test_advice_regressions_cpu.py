"""Regression tests for host-logic defects found in review (kernels replaced by the CPU emulation):
in-place activations behind a captured layer, modules re-used within one forward pass, tensor-valued scalars in the
Kron algebra (autograd to ``sigma_noise``), weight-pack cache validity, mixed Kron sums keeping the flat buffer."""
import gc

import pytest
import torch
from torch import nn

from laplace_b200 import B200GGN, B200Kron, B200KronDecomposed
from oracle import curvature_oracle as co
from tests.fixtures import rel_fro


def _data(n=12, d=6, c=3, seed=0):
    torch.manual_seed(seed)
    return torch.randn(n, d, dtype=torch.float64), torch.randint(c, (n,))


@pytest.mark.parametrize("act", [lambda: nn.ReLU(inplace=True), lambda: nn.Hardtanh(inplace=True), lambda: nn.ReLU6(inplace=True)])
def test_inplace_activation_module_after_captured_layer(cpu_kernels, act):
    """``autograd.grad(f, out)`` differentiates w.r.t. the current version of ``out``: with ``ReLU(inplace=True)`` behind a
    Linear the gradient used to lack the activation mask (advisor repro).  The reference's hooks fire before the
    in-place op; so must ours."""
    torch.manual_seed(1)
    model = nn.Sequential(nn.Linear(6, 8), act(), nn.Linear(8, 3)).double()
    ref = nn.Sequential(model[0], type(model[1])(), model[2])       # same parameters, out-of-place activation
    X, y = _data()
    _, kf = co.kfac_factors(ref, "classification", X, y, N=len(X))
    be = B200GGN(model, "classification")
    _, kron = be.kron(X, y, N=len(X))
    for F, Fo in zip(kron.kfacs, kf):
        for H, Ho in zip(F, Fo):
            assert rel_fro(H, Ho) < 1e-6
    Js, _ = be.jacobians(X)
    Jo, _ = co.jacobians(ref, X)
    assert rel_fro(Js, Jo) < 1e-6
    assert model[1].inplace is True     # the flag is restored after the captured pass


def test_functional_inplace_activation_is_reported(cpu_kernels):
    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(6, 8), nn.Linear(8, 3)

        def forward(self, x):
            return self.b(torch.relu_(self.a(x)))

    X, y = _data()
    with pytest.raises(RuntimeError, match="modified in place"):
        B200GGN(Net().double(), "classification").kron(X, y, N=len(X))


def test_inplace_residual_add_is_fine(cpu_kernels):
    """``out += identity`` (torchvision's BasicBlock) leaves the gradient w.r.t. the layer output unchanged."""
    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b, self.c = nn.Linear(6, 6), nn.Linear(6, 6), nn.Linear(6, 3)

        def forward(self, x):
            h = torch.tanh(self.a(x))
            out = self.b(h)
            out += h
            return self.c(torch.tanh(out))

    class Ref(Net):
        def forward(self, x):
            h = torch.tanh(self.a(x))
            return self.c(torch.tanh(self.b(h) + h))

    torch.manual_seed(2)
    net = Net().double()
    ref = Ref().double()
    ref.load_state_dict(net.state_dict())
    X, y = _data()
    _, kf = co.kfac_factors(ref, "classification", X, y, N=len(X))
    _, kron = B200GGN(net, "classification").kron(X, y, N=len(X))
    for F, Fo in zip(kron.kfacs, kf):
        for H, Ho in zip(F, Fo):
            assert rel_fro(H, Ho) < 1e-6


def test_module_reused_in_forward_raises(cpu_kernels):
    class Tied(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(6, 6), nn.Linear(6, 3)

        def forward(self, x):
            return self.b(torch.tanh(self.a(torch.tanh(self.a(x)))))

    X, y = _data()
    with pytest.raises(ValueError, match="more than once"):
        B200GGN(Tied().double(), "classification").kron(X, y, N=len(X))


def test_tensor_scalar_keeps_autograd_graph():
    """``KronDecomposed * _H_factor`` with ``_H_factor = 1 / sigma_noise^2`` a tensor (regression,
    baselaplace.py:593-596, :1820): d logdet / d sigma_noise must match the dense formula (utils/matrix.py:372-376)."""
    torch.manual_seed(3)

    def psd(n):
        Z = torch.randn(n, 3 * n, dtype=torch.float64)
        return Z @ Z.T / (3 * n)

    kron = B200Kron([[psd(4), psd(5)], [psd(4)]])
    kd = kron.decompose()
    assert isinstance(kd, B200KronDecomposed)
    sigma = torch.tensor(0.7, dtype=torch.float64, requires_grad=True)
    delta = torch.tensor(0.3, dtype=torch.float64)
    ld = (kd * (1.0 / sigma ** 2) + delta).logdet()
    (g,) = torch.autograd.grad(ld, sigma)
    s2 = sigma.detach().clone().requires_grad_(True)
    dense = torch.block_diag(torch.kron(kron.kfacs[0][0], kron.kfacs[0][1]), kron.kfacs[1][0]) / s2 ** 2
    ld2 = torch.logdet(dense + delta * torch.eye(dense.shape[0], dtype=torch.float64))
    (g2,) = torch.autograd.grad(ld2, s2)
    assert torch.allclose(ld, ld2, rtol=1e-9) and torch.allclose(g, g2, rtol=1e-8)
    # and on the un-decomposed container (utils/matrix.py:116-118)
    t = torch.tensor(2.0, dtype=torch.float64, requires_grad=True)
    tot = sum(h.sum() for F in (kron * t).kfacs for h in F)
    assert torch.autograd.grad(tot, t)[0].abs() > 0


def test_weight_cache_dies_with_the_module(cpu_kernels):
    from laplace_b200 import conv_engine

    conv = nn.Conv2d(3, 4, 3, 1, 1)
    with torch.no_grad():
        packed = conv_engine._CACHE.get(conv, "bwd")
        assert conv_engine._CACHE.get(conv, "bwd") is packed          # same tensor, same version: cache hit
        conv.weight.mul_(2.0)                                           # in-place update bumps the version
        assert conv_engine._CACHE.get(conv, "bwd") is not packed
    n = len(conv_engine._CACHE.store)
    del conv
    gc.collect()
    assert len(conv_engine._CACHE.store) == n - 1                       # no entry outlives its module


def test_sum_with_plain_kron_keeps_flat_buffer():
    from laplace_b200.interface import Kron

    k = B200Kron.zeros([[3, 4], [3]], "cpu")
    k._flat.fill_(1.0)
    plain = Kron([[torch.zeros(3, 3), torch.zeros(4, 4)], [torch.zeros(3, 3)]])
    out = plain + k          # ``la.H += H_batch`` on the first batch: Python prefers the subclass' reflected method
    out2 = k + plain
    for o in (out, out2):
        assert isinstance(o, B200Kron) and o._flat is not None and float(o._flat.sum()) == 9 + 16 + 9
    out2 += k
    assert float(out2._flat.sum()) == 2 * (9 + 16 + 9)
