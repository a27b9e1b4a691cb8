"""The reference's own backend test cases (tests/test_curv_backends_curvlinops.py:84-345), case by case, with ``B200GGN`` /
``B200EF`` in place of ``CurvlinopsGGN`` / ``CurvlinopsEF`` -- same fixtures (seed 711, MLP 3-20-2, the ``Conv2d(3, 4, 2, 2)``
"complex model", 10 samples), same assertions, the reference's tolerances (``torch.allclose`` defaults unless it states one).
Where the reference compares against a second library (ASDL / BackPACK, absent here) the comparison is against the fp64 oracle
and the golden vectors generated from the unmodified reference.  Kernels: the CPU emulation (``-m gpu`` repeats the
invariants on the device, tests/test_gpu_parity.py)."""
import pytest
import torch

from laplace_b200 import B200EF, B200GGN
from oracle import curvature_oracle as co
from tests.fixtures import load, rel_fro

BACKENDS = [B200EF, B200GGN]


def _xy(golden, kind, lik):
    model, X, y, rec = load(golden, kind, lik)                 # float64 models, as in the reference's tests
    return model, X, y, rec


@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_full_ggn_and_ef_vs_second_implementation(golden, cpu_kernels, lik):
    """:98-108, :132-142 (vs ASDL there): loss rtol 1e-4, H rtol 1e-4."""
    model, X, y, rec = _xy(golden, "mlp", lik)
    Jo, fo = co.jacobians(model, X)
    for Backend, ref in ((B200GGN, co.ggn_full), (B200EF, co.ef_full)):
        loss, H = Backend(model, lik).full(X, y)
        loss_ref, H_ref = ref(Jo, fo, y, lik)
        assert torch.allclose(loss, loss_ref, rtol=1e-4)
        assert torch.allclose(H, H_ref, rtol=1e-4, atol=1e-6 * float(H_ref.abs().max()))


def test_full_ggn_stochastic(golden, cpu_kernels):
    """:111-129: same loss, and 100 MC samples are closer to the exact GGN than 1."""
    torch.manual_seed(123)
    model, X, y, _ = _xy(golden, "mlp", "classification")
    loss_mc1, H_mc1 = B200GGN(model, "classification", stochastic=True).full(X, y, mc_samples=1)
    loss_mc100, H_mc100 = B200GGN(model, "classification", stochastic=True).full(X, y, mc_samples=100)
    loss_exact, H_exact = B200GGN(model, "classification", stochastic=False).full(X, y)
    assert torch.allclose(loss_mc1, loss_exact) and torch.allclose(loss_mc100, loss_exact)
    assert torch.norm(H_mc1 - H_exact) > torch.norm(H_mc100 - H_exact)
    # the constructor's ``num_samples`` is the default of the per-call ``mc_samples``
    torch.manual_seed(5)
    _, Ha = B200GGN(model, "classification", stochastic=True, num_samples=7).full(X, y)
    torch.manual_seed(5)
    _, Hb = B200GGN(model, "classification", stochastic=True).full(X, y, mc_samples=7)
    assert torch.equal(Ha, Hb)
    torch.manual_seed(5)
    _, da = B200GGN(model, "classification", stochastic=True, num_samples=7).diag(X, y)
    assert torch.allclose(da, Ha.diagonal(), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_kron_ggn_vs_second_implementation(golden, cpu_kernels, lik):
    """:145-156 (vs BackPACK there): loss equal, ``to_matrix()`` rtol 5e-5."""
    model, X, y, _ = _xy(golden, "mlp", lik)
    loss, kron = B200GGN(model, lik, stochastic=False).kron(X, y, N=1)
    loss_ref, kfacs = co.kfac_factors(model, lik, X, y, N=1)
    assert torch.allclose(loss, loss_ref)
    M_ref = co.kfacs_to_matrix(kfacs)
    assert torch.allclose(kron.to_matrix(), M_ref, rtol=5e-5, atol=1e-6 * float(M_ref.abs().max()))


@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_kron_ggn_stochastic(golden, cpu_kernels, lik):
    """:159-177."""
    torch.manual_seed(0)
    model, X, y, _ = _xy(golden, "mlp", lik)
    loss_mc1, kron_mc1 = B200GGN(model, lik, stochastic=True).kron(X, y, N=1, mc_samples=1)
    loss_mc100, kron_mc100 = B200GGN(model, lik, stochastic=True).kron(X, y, N=1, mc_samples=100)
    loss_ref, kron_exact = B200GGN(model, lik, stochastic=False).kron(X, y, N=1)
    assert torch.allclose(loss_mc1, loss_ref) and torch.allclose(loss_mc100, loss_ref)
    exact = kron_exact.to_matrix()
    assert torch.norm(kron_mc1.to_matrix() - exact) > torch.norm(kron_mc100.to_matrix() - exact)


@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_kron_ggn_set_kfac_approx(golden, cpu_kernels, lik):
    """:180-193: expand and reduce share the loss and differ on a weight-sharing model."""
    model, X, y, _ = _xy(golden, "conv", lik)
    loss_expand, kron_expand = B200GGN(model, lik).kron(X, y, N=1, kfac_approx="expand")
    loss_reduce, kron_reduce = B200GGN(model, lik).kron(X, y, N=1, kfac_approx="reduce")
    assert torch.allclose(loss_expand, loss_reduce)
    assert not torch.allclose(kron_expand.to_matrix(), kron_reduce.to_matrix())
    with pytest.raises(ValueError):
        B200GGN(model, lik).kron(X, y, N=1, kfac_approx="bogus")


def test_kron_ef_vs_second_implementation(golden, cpu_kernels):
    """:195-205 (vs ASDL there)."""
    model, X, y, _ = _xy(golden, "mlp", "classification")
    loss, kron = B200EF(model, "classification").kron(X, y, N=1)
    loss_ref, kfacs = co.kfac_factors(model, "classification", X, y, N=1, fisher="empirical")
    assert torch.allclose(loss, loss_ref)
    M_ref = co.kfacs_to_matrix(kfacs)
    assert torch.allclose(kron.to_matrix(), M_ref, rtol=5e-5, atol=1e-6 * float(M_ref.abs().max()))


@pytest.mark.parametrize("Backend", BACKENDS)
@pytest.mark.parametrize("kind,lik", [("mlp", "classification"), ("mlp", "regression"), ("conv", "classification")])
def test_kron_batching_correction(golden, cpu_kernels, Backend, kind, lik):
    """:208-242, :279-296: the whole batch equals the sum of two parts when ``N`` is the global size."""
    model, X, y, _ = _xy(golden, kind, lik)
    backend = Backend(model, lik)
    n_params = sum(p.numel() for p in model.parameters())
    loss, kron = backend.kron(X, y, N=len(X))
    assert len(kron.diag()) == n_params
    N, M = len(X), 3
    loss1, kron1 = backend.kron(X[:M], y[:M], N=N)
    loss2, kron2 = backend.kron(X[M:], y[M:], N=N)
    kron_two, loss_two = kron1 + kron2, loss1 + loss2
    assert torch.allclose(kron.diag(), kron_two.diag())
    assert torch.allclose(loss, loss_two)


@pytest.mark.parametrize("Backend", BACKENDS)
def test_kron_summing_up_vs_diag(golden, cpu_kernels, Backend):
    """:245-251 (rtol 1e-1) and :299-309 (complex model, rtol 1e-2): the norm of the Kron diagonal tracks the diagonal GGN / EF."""
    for kind, rtol in (("mlp", 1e-1), ("conv", 1e-2)):
        model, X, y, _ = _xy(golden, kind, "classification")
        backend = Backend(model, "classification")
        loss, dggn = backend.diag(X, y, N=len(X))
        loss, kron = backend.kron(X, y, N=len(X))
        assert torch.allclose(kron.diag().norm(), dggn.norm(), rtol=rtol), kind


def test_complex_diag_ggn_stochastic(golden, cpu_kernels):
    """:254-264: size, same loss, same order of magnitude as a second stochastic draw."""
    torch.manual_seed(1)
    model, X, y, _ = _xy(golden, "conv", "classification")
    backend = B200GGN(model, "classification", stochastic=True)
    loss, dggn = backend.diag(X, y)
    assert len(dggn) == sum(p.numel() for p in model.parameters())
    loss_ns, dggn_ns = backend.diag(X, y)
    assert loss_ns == loss
    assert torch.allclose(dggn, dggn_ns, atol=1e-8, rtol=1)


@pytest.mark.parametrize("Backend", BACKENDS)
def test_complex_kron_single_datum_vs_diag(golden, cpu_kernels, Backend):
    """:267-277: for one data point the Kron diagonal has the diagonal curvature's norm (rtol 1e-1)."""
    model, X, y, _ = _xy(golden, "conv", "classification")
    backend = Backend(model, "classification")
    loss, dggn = backend.diag(X[:1], y[:1], N=1)
    assert len(dggn) == sum(p.numel() for p in model.parameters())
    loss, kron = backend.kron(X[:1], y[:1], N=1)
    assert torch.allclose(kron.diag().norm(), dggn.norm(), rtol=1e-1)


@pytest.mark.parametrize("Backend", BACKENDS)
def test_kron_normalization(golden, cpu_kernels, Backend):
    """:312-339: seven copies of one sample with ``N = 7`` give seven times its curvature and loss."""
    model, X, y, _ = _xy(golden, "mlp", "classification")
    xi, yi = X[:1], y[:1]
    backend = Backend(model, "classification")
    loss, kron = backend.kron(xi, yi, N=1)
    kron_true, loss_true = 7 * kron, 7 * loss
    X7, y7 = torch.repeat_interleave(xi, 7, 0), torch.repeat_interleave(yi, 7, 0)
    loss_test, kron_test = backend.kron(X7, y7, N=7)
    assert torch.allclose(kron_true.diag(), kron_test.diag())
    assert torch.allclose(loss_true, loss_test)
