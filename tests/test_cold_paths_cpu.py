"""Cold paths the hot-path tests do not reach (found with a line-coverage run of the CPU suite): serialization round trip +
``adopt``, the ``Kron`` helper algebra against the reference class (utils/matrix.py:79-275), mixed-layout sums,
posterior sampling through ``bmm(exponent=-1/2)`` (utils/matrix.py:463-488), and the host path of ``PrefetchLoader``."""
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

from oracle import ref_shim
from tests.fixtures import load, rel_fro

needs_reference = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


def _spd(n, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(n, n + 3, generator=g)
    return A @ A.t() / (n + 3) + 0.1 * torch.eye(n)


@needs_reference
def test_kron_helpers_match_reference_class(cpu_kernels):
    from laplace.utils.matrix import Kron

    from laplace_b200 import B200Kron

    kfacs = [[_spd(3, 0), _spd(4, 1)], [_spd(3, 2)], [_spd(2, 3), _spd(3, 4)], [_spd(2, 5)]]
    ours, ref = B200Kron.from_kfacs(kfacs), Kron([[H.clone() for H in F] for F in kfacs])
    assert ours._flat is not None and ours._flat.numel() == sum(H.numel() for F in kfacs for H in F)
    assert torch.allclose(ours.diag(), ref.diag()) and torch.allclose(ours.to_matrix(), ref.to_matrix())
    assert torch.allclose(ours.logdet(), ref.logdet(), rtol=1e-6)
    # scalar algebra: factor ** (1 / len(F)) per block (utils/matrix.py:100-118), python and tensor scalars
    for s in (2.5, torch.tensor(0.3)):
        assert torch.allclose((ours * s).to_matrix(), (ref * s).to_matrix(), rtol=1e-6)
        assert torch.allclose((s * ours).to_matrix(), (ref * s).to_matrix(), rtol=1e-6)
    with pytest.raises(ValueError):
        ours * "2"
    with pytest.raises(ValueError):
        ours + 1.0
    # sums: same layout (one fused add on the flat buffers), a plain reference Kron on either side (keeps a flat buffer),
    # and a B200Kron without a flat buffer (factor by factor)
    two, three = (ref + ref).to_matrix(), (ref + ref + ref).to_matrix()      # factor-wise sums (utils/matrix.py:79-98)
    both = ours + ours
    assert both._flat is not None and torch.allclose(both.to_matrix(), two)
    for mixed in (ours + ref, ref + ours):
        assert isinstance(mixed, B200Kron) and mixed._flat is not None and torch.allclose(mixed.to_matrix(), two)
    loose = B200Kron([[H.clone() for H in F] for F in kfacs])
    assert loose._flat is None and torch.allclose((loose + ours).to_matrix(), two)
    acc = B200Kron.from_kfacs(kfacs)
    acc += loose                                   # layouts differ: per-factor in-place adds, still the caller's buffer
    acc += ours
    assert acc._flat is not None and torch.allclose(acc.to_matrix(), three)
    with pytest.raises(ValueError):
        acc += 1.0
    # float64 factors do not fit the fp32 flat buffer: kept as they are
    dbl = B200Kron.from_kfacs([[H.double() for H in F] for F in kfacs])
    assert dbl._flat is None and dbl.kfacs[0][0].dtype == torch.float64
    # decomposition of a diagonal (1-D) factor next to square ones (utils/matrix.py:141-145)
    dec = B200Kron([[torch.tensor([1.0, 2.0, 3.0]), _spd(4, 1)]]).decompose()
    refdec = Kron([[torch.tensor([1.0, 2.0, 3.0]), _spd(4, 1)]]).decompose()
    assert torch.allclose(dec.logdet(), refdec.logdet(), rtol=1e-5)
    assert torch.allclose((dec + torch.tensor(0.5)).to_matrix(), (refdec + torch.tensor(0.5)).to_matrix(), rtol=1e-4, atol=1e-5)


@needs_reference
@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_state_dict_round_trip_adopt_and_sampling(golden, cpu_kernels, lik):
    """``la.state_dict()`` -> ``load_state_dict`` (baselaplace.py:1845-1879) rebuilds ``H_facs`` as a plain ``Kron``;
    ``adopt`` puts the factors back on this package's containers.  Same predictive and marginal likelihood before and after;
    ``sample()`` (``mean + bmm(eps, exponent=-1/2)``) has the posterior covariance."""
    import laplace
    from laplace.utils.matrix import Kron

    from laplace_b200 import B200GGN, B200Kron, B200KronDecomposed, adopt

    model, X, y, _ = load(golden, "mlp", lik, dtype=torch.float32)
    loader = DataLoader(TensorDataset(X, y), batch_size=5)
    la = laplace.Laplace(model, lik, "all", "kron", backend=B200GGN, prior_precision=0.9)
    la.fit(loader)
    f_mu, f_var = la._glm_predictive_distribution(X)
    lml = la.log_marginal_likelihood()
    sd = la.state_dict()
    la2 = laplace.Laplace(model, lik, "all", "kron", backend=B200GGN, prior_precision=0.9)
    la2.load_state_dict(sd)
    assert type(la2.H_facs) is Kron
    assert adopt(la2) is la2 and isinstance(la2.H_facs, B200Kron) and la2.H_facs._flat is not None
    assert isinstance(la2.H, B200KronDecomposed) and adopt(la2).H_facs is la2.H_facs      # idempotent
    f_mu2, f_var2 = la2._glm_predictive_distribution(X)
    assert torch.allclose(f_mu2, f_mu, atol=1e-6) and torch.allclose(f_var2, f_var, rtol=1e-4, atol=1e-7)
    assert torch.allclose(la2.log_marginal_likelihood(), lml, rtol=1e-5)
    # sampling: empirical covariance of theta - mean against the dense posterior covariance
    Sigma = torch.linalg.inv(la2.posterior_precision.to_matrix().double())
    gen = torch.Generator().manual_seed(0)
    S = la2.sample(20000, generator=gen).double() - la2.mean.double()
    emp = S.t() @ S / len(S)
    assert rel_fro(emp, Sigma) < 2.0 * (la2.n_params / len(S)) ** 0.5      # sampling noise of a covariance estimate ~ sqrt(P / n)
    # exponent algebra of the decomposed precision: P^-1/2 P^-1/2 = P^-1, P^1 P^-1 = I (utils/matrix.py:463-488)
    P = la2.posterior_precision
    W = torch.randn(7, la2.n_params, generator=gen)
    half = P.bmm(P.bmm(W, exponent=-0.5), exponent=-0.5)
    assert rel_fro(half, P.bmm(W, exponent=-1)) < 1e-4
    assert rel_fro(P.bmm(P.bmm(W, exponent=1), exponent=-1), W) < 1e-4
    assert rel_fro(P.bmm(W, exponent=-1).double(), W.double() @ Sigma) < 1e-4


def test_prefetch_loader_host_path_and_structure():
    """On a CPU device ``PrefetchLoader`` is a pass-through that still honours the containers (tuples, lists, mappings,
    non-tensor leaves) and forwards ``dataset`` / ``len``."""
    from laplace_b200.data import PrefetchLoader

    X, y = torch.randn(10, 3), torch.arange(10)
    base = DataLoader(TensorDataset(X, y), batch_size=4)
    pl = PrefetchLoader(base, "cpu")
    assert len(pl) == 3 and pl.dataset is base.dataset
    got = list(pl)
    assert isinstance(got[0], (tuple, list)) and torch.equal(torch.cat([b[0] for b in got]), X)
    assert torch.equal(torch.cat([b[1] for b in got]), y)

    class Batches(list):
        dataset = range(4)

    nested = Batches([{"input_ids": X[:2], "labels": y[:2], "meta": "a", "extra": (X[2:4], [y[2:4]])}])
    out = list(PrefetchLoader(nested, torch.device("cpu"), depth=0))
    assert out[0]["meta"] == "a" and torch.equal(out[0]["extra"][1][0], y[2:4]) and isinstance(out[0]["extra"], tuple)
    assert sum(1 for _ in PrefetchLoader._tensors(out[0])) == 4

    class Encoding(dict):                      # Hugging Face ``BatchEncoding``-like: has its own ``.to``
        moved = None

        def to(self, device):
            Encoding.moved = torch.device(device)
            return self

    enc = Encoding(input_ids=X[:2])
    assert PrefetchLoader(Batches([enc]), "cpu")._move(enc) is enc and Encoding.moved == torch.device("cpu")
