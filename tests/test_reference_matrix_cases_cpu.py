"""The reference's own ``Kron`` / ``KronDecomposed`` test cases (tests/test_matrix.py:33-300), case by case, with ``B200Kron`` /
``B200KronDecomposed`` -- same sizes (``[[20, 3], [20], [2, 20], [2]]``), same seed (7171), float64 as the reference's module
sets, same assertions.  The dense block matrices they compare against are built here from the factors (``torch.kron`` /
``torch.block_diag``); the Jacobians of the ``bmm`` case come from the fp64 oracle instead of ``tests/utils.py``.
The rotation GEMMs of the emulated kernels run in fp32 like the device kernels, so the comparisons that go through them use
rtol 1e-4 / a scale-relative atol where the reference (fp64 throughout) uses the ``allclose`` defaults."""
import pytest
import torch

from laplace_b200 import B200GGN, B200Kron, B200KronDecomposed
from oracle import curvature_oracle as co

SIZES = [[20, 3], [20], [2, 20], [2]]
P = 20 * 3 + 20 + 2 * 20 + 2
D = torch.float64


def psd(n):
    X = torch.randn(n, 3 * n, dtype=D)
    return X @ X.T / (3 * n)


def diag_psd(n):
    return torch.randn(n, dtype=D) ** 2


def dense(kfacs):
    blocks = []
    for F in kfacs:
        F0 = F[0] if F[0].ndim > 1 else F[0].diag()
        blocks.append(F0 if len(F) == 1 else torch.kron(F0, F[1] if F[1].ndim > 1 else F[1].diag()))
    return torch.block_diag(*blocks)


def close(a, b, rtol=1e-4):
    return torch.allclose(a.double(), b.double(), rtol=rtol, atol=rtol * float(b.abs().max()))


def test_zero_initialisation_has_the_layout_of_init_from_model():
    """:33-52 -- ``Kron.init_from_model`` sizes; here: ``B200Kron.zeros`` (what ``kron()`` starts from) and, with the reference
    importable, the inherited ``init_from_model`` itself."""
    model = torch.nn.Sequential(torch.nn.Linear(3, 20), torch.nn.Linear(20, 2))
    expected = [[20 * 20, 3 * 3], [20 * 20], [2 * 2, 20 * 20], [2 * 2]]
    krons = [B200Kron.zeros(SIZES, "cpu", torch.float32)]
    if hasattr(B200Kron, "init_from_model"):
        krons += [B200Kron.init_from_model(model, "cpu", torch.float), B200Kron.init_from_model(model.parameters(), "cpu", torch.float)]
    for kron in krons:
        assert len(kron.kfacs) == len(expected)
        for facs, exp in zip(kron.kfacs, expected):
            assert [f.numel() for f in facs] == exp and all(torch.all(f == 0) for f in facs)


def test_addition(cpu_kernels):
    """:55-62 -- ``kron += to_add`` on a zero-initialised container (what ``la.H += H_batch`` does, baselaplace.py:985)."""
    torch.manual_seed(0)
    to_add = B200Kron.from_kfacs([[torch.randn(i, i) for i in sizes] for sizes in SIZES])
    kron = B200Kron.zeros(SIZES, "cpu", torch.float32)
    kron += to_add
    for facs, exp in zip(kron.kfacs, to_add.kfacs):
        for fi, ei in zip(facs, exp):
            assert torch.allclose(fi, ei)
    from laplace_b200.interface import HAVE_REFERENCE

    if HAVE_REFERENCE:      # a plain reference Kron of zeros on the left: the subclass' reflected add takes over
        from laplace.utils import Kron

        model = torch.nn.Sequential(torch.nn.Linear(3, 20), torch.nn.Linear(20, 2))
        H = Kron.init_from_model(model.parameters(), "cpu", torch.float)
        H += to_add
        assert isinstance(H, B200Kron) and H._flat is not None
        assert all(torch.allclose(fi, ei) for facs, exp in zip(H.kfacs, to_add.kfacs) for fi, ei in zip(facs, exp))


def test_multiplication(cpu_kernels):
    """:65-78 -- a scalar multiplies every block once (split as ``s ** (1/len)`` over a block's factors)."""
    torch.manual_seed(1)
    kfacs = [[torch.randn(i, i, dtype=D) for i in sizes] for sizes in SIZES]
    kron = B200Kron([[f.clone() for f in F] for F in kfacs])
    kron *= 1.5
    for facs, exp in zip(kron.kfacs, kfacs):
        if len(facs) == 1:
            assert torch.allclose(facs[0], 1.5 * exp[0])
        else:
            assert torch.allclose(torch.kron(*facs), 1.5 * torch.kron(*exp))


@pytest.mark.parametrize("make", [psd, diag_psd], ids=["dense", "diagonal"])
def test_decompose(cpu_kernels, make):
    """:81-127 -- eigendecompositions reconstruct the factors; ``decomposed.bmm(W, exponent=1)`` equals the matrix product."""
    torch.manual_seed(7171)
    kfacs = [[make(i) for i in sizes] for sizes in SIZES]
    kron = B200Kron(kfacs)
    dec = kron.decompose()
    assert isinstance(dec, B200KronDecomposed)
    for facs, Qs, ls in zip(kron.kfacs, dec.eigenvectors, dec.eigenvalues):
        recs = [Q @ torch.diag(l) @ Q.T for Q, l in zip(Qs, ls)]
        full = [f if f.ndim > 1 else f.diag() for f in facs]
        if len(facs) == 1:
            assert torch.allclose(full[0], recs[0], rtol=1e-3, atol=1e-10)
        else:
            assert torch.allclose(torch.kron(*full), torch.kron(*recs), rtol=1e-2, atol=1e-10)
    W = torch.randn(P, dtype=D)
    assert close(dec.bmm(W, exponent=1), W @ dense(kfacs))
    if hasattr(kron, "bmm"):                         # the reference's undecomposed product, inherited
        assert close(dec.bmm(W, exponent=1), kron.bmm(W))


@pytest.mark.parametrize("make", [psd, diag_psd], ids=["dense", "diagonal"])
def test_logdet_consistent(cpu_kernels, make):
    """:130-140."""
    torch.manual_seed(7171)
    kron = B200Kron([[make(i) for i in sizes] for sizes in SIZES])
    assert torch.allclose(kron.logdet(), kron.decompose().logdet()) and torch.allclose(kron.logdet(), torch.logdet(dense(kron.kfacs)))


def test_bmm_dense(cpu_kernels):
    """:143-199 -- ``J S``, ``J S J^T``, ``J S^-1 J^T`` (functional variance), ``J S^-1/2`` (sampling) and the 2-D / 1-D input
    shapes, on the factors ``B200GGN.kron`` returns for the reference's small regression model."""
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2)).double()
    X, y = torch.randn(5, 3, dtype=D), torch.randn(5, 2, dtype=D)
    _, kron = B200GGN(model, "regression", stochastic=False).kron(X, y, N=5)
    dec = kron.decompose()
    Js, _ = co.jacobians(model, X)
    S = dense(kron.kfacs)
    assert torch.allclose(S, S.T) and torch.allclose(S.diagonal(), kron.diag())
    JS = dec.bmm(Js, exponent=1)
    JS_true = Js @ S
    assert close(JS, JS_true) and close(torch.bmm(JS, Js.transpose(1, 2)), torch.bmm(JS_true, Js.transpose(1, 2)))
    if hasattr(kron, "bmm"):
        assert close(kron.bmm(Js), JS)
    # the reference's S is singular for this model (more parameters than the rank 5 samples x 2 outputs give): it adds
    # nothing before inverting because torch's inverse happens to succeed; a prior precision makes the case well posed
    delta = torch.tensor(0.7, dtype=D)
    Sd = S + delta * torch.eye(len(S), dtype=D)
    decd = dec + delta
    S_inv = Sd.inverse()
    assert close(decd.inv_square_form(Js), torch.bmm(Js @ S_inv, Js.transpose(1, 2)))
    ev, Q = torch.linalg.eigh(S_inv, UPLO="U")
    JS_half = Js @ Q @ torch.diag(torch.sqrt(ev)) @ Q.T
    got = decd.bmm(Js, exponent=-1 / 2)
    assert close(got, JS_half) and close(torch.bmm(got, Js.transpose(1, 2)), torch.bmm(JS_half, Js.transpose(1, 2)))
    for W in (Js[:, 0, :].squeeze(), Js[0, 0, :].squeeze()):           # 2-D and 1-D inputs
        assert close(dec.bmm(W, exponent=1), W @ S)
        if hasattr(kron, "bmm"):
            assert close(dec.bmm(W, exponent=1), kron.bmm(W))


@pytest.mark.parametrize("make", [psd, diag_psd], ids=["dense", "diagonal"])
def test_matrix_consistent(cpu_kernels, make):
    """:269-300 -- ``to_matrix`` of the container and of its decomposition, inverse through ``exponent=-1``, ``+= delta``."""
    torch.manual_seed(7171)
    kfacs = [[make(i) for i in sizes] for sizes in SIZES]
    kron = B200Kron(kfacs)
    dec = kron.decompose()
    M = kron.to_matrix()
    assert torch.allclose(M, dense(kfacs))
    assert torch.allclose(M, dec.to_matrix(exponent=1), rtol=1e-6, atol=1e-10)
    assert torch.allclose(M.inverse(), dec.to_matrix(exponent=-1), rtol=1e-5, atol=1e-8 * float(M.inverse().abs().max()))
    M_true = M.clone()
    M_true.diagonal().add_(3.4)
    dec += torch.tensor(3.4, dtype=D)
    assert torch.allclose(M_true, dec.to_matrix(exponent=1), rtol=1e-6, atol=1e-10)
