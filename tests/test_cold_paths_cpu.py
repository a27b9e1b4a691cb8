"""Cold paths the hot-path tests do not reach (found with a line-coverage run of the CPU suite): serialization round trip +
``adopt``, the ``Kron`` helper algebra against the reference class (utils/matrix.py:79-275), mixed-layout sums,
posterior sampling through ``bmm(exponent=-1/2)`` (utils/matrix.py:463-488), and the host path of ``PrefetchLoader``."""
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

from oracle import ref_shim
from tests.fixtures import load, rel_fro



def _reference_usable():
    if not ref_shim.reference_available():
        return False
    try:
        import laplace  # noqa: F401
    except ImportError:
        return False
    from laplace_b200.interface import HAVE_REFERENCE

    return HAVE_REFERENCE


needs_reference = pytest.mark.skipif(not _reference_usable(), reason="reference package not importable (not mounted, or LPB_NO_REFERENCE=1)")


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


@pytest.mark.parametrize("kind", ["mlp", "conv"])
@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_subnetwork_indices_select_columns_everywhere(golden, cpu_kernels, kind, lik):
    """``subnetwork_indices`` (curvature/curvature.py:76-86, 125-127): Jacobians, gradients, full and diagonal GGN / EF are
    those of the selected parameter columns -- against the fp64 oracle's dense quantities restricted to the same columns."""
    from laplace_b200 import B200EF, B200GGN
    from oracle import curvature_oracle as co

    model, X, y, rec = load(golden, kind, lik, dtype=torch.float32)
    P = sum(p.numel() for p in model.parameters())
    idx = torch.randperm(P, generator=torch.Generator().manual_seed(1))[: P // 3].sort().values
    md = load(golden, kind, lik)[0]
    Jo, fo = co.jacobians(md, X.double())
    yd = y if lik == "classification" else y.double()
    Js_sub = Jo[:, :, idx]
    ggn = B200GGN(model, lik, subnetwork_indices=idx)
    Js, f = ggn.jacobians(X)
    assert Js.shape == (len(X), fo.shape[1], len(idx)) and rel_fro(Js, Js_sub) < 1e-5
    Gs, loss = ggn.gradients(X, y)
    Go, lo = co.gradients(Jo, fo, yd, lik)
    assert rel_fro(Gs, Go[:, idx]) < 1e-5 and torch.allclose(loss.double(), lo, rtol=1e-5)
    _, H = ggn.full(X, y)
    _, Ho = co.ggn_full(Js_sub, fo, yd, lik)
    assert rel_fro(H, Ho) < 1e-5
    _, d = ggn.diag(X, y)
    _, do = co.ggn_diag(Js_sub, fo, yd, lik)
    assert rel_fro(d, do) < 1e-5
    ef = B200EF(model, lik, subnetwork_indices=idx)
    _, He = ef.full(X, y)
    _, Heo = co.ef_full(Js_sub, fo, yd, lik)
    assert rel_fro(He, Heo) < 1e-5
    _, de = ef.diag(X, y)
    _, deo = co.ef_diag(Js_sub, fo, yd, lik)
    assert rel_fro(de, deo) < 1e-5


@needs_reference
def test_enable_backprop_delegates_to_the_reference_path(golden, cpu_kernels):
    """``enable_backprop=True`` (curvature/curvature.py:88-129): the Jacobians must stay attached to the autograd graph of the
    input; the kernels are not differentiable, so both Jacobian entry points hand over to the reference's torch.func code."""
    from laplace_b200 import B200GGN

    model, X, y, rec = load(golden, "mlp", "classification", dtype=torch.float32)
    be = B200GGN(model, "classification")
    Xg = X.clone().requires_grad_(True)
    Js, f = be.jacobians(Xg, enable_backprop=True)
    assert Js.requires_grad and rel_fro(Js.detach(), rec["Js"]) < 1e-5
    (g,) = torch.autograd.grad(Js.square().sum(), Xg)
    assert g.shape == X.shape and torch.isfinite(g).all() and float(g.abs().max()) > 0
    Js_plain, _ = be.jacobians(X)
    assert not Js_plain.requires_grad and torch.allclose(Js_plain, Js.detach(), atol=1e-6)


@pytest.mark.skipif(_reference_usable(), reason="mirror mode only (LPB_NO_REFERENCE=1 / GPU box without baseline/_ref)")
def test_enable_backprop_without_the_reference_says_why(golden, cpu_kernels):
    from laplace_b200 import B200GGN

    model, X, _, _ = load(golden, "mlp", "classification", dtype=torch.float32)
    be = B200GGN(model, "classification")
    for fn in (be.jacobians, be.last_layer_jacobians):
        with pytest.raises(NotImplementedError, match="not differentiable"):
            fn(X, enable_backprop=True)


def test_unsupported_layer_configurations_raise(cpu_kernels):
    from laplace_b200 import B200GGN

    X, y = torch.randn(2, 4, 6, 6), torch.randint(3, (2,))
    for conv in (torch.nn.Conv2d(4, 4, 3, groups=2), torch.nn.Conv2d(4, 4, 3, padding="same"),
                 torch.nn.Conv2d(4, 4, 3, padding=1, padding_mode="reflect")):
        net = torch.nn.Sequential(conv, torch.nn.Flatten(), torch.nn.LazyLinear(3))
        net(X)
        with pytest.raises(ValueError, match="not supported|only zero padding"):
            B200GGN(net, "classification").kron(X, y, N=2)


def test_ops_without_a_batching_rule_fall_back_to_one_pass_per_column(golden, cpu_kernels):
    """A custom ``autograd.Function`` whose reverse pass cannot be vmapped (it reads a value on the host) between the layers: the column-batched reverse pass is not
    available, ``last_backward_mode`` says so, and the per-column passes give the same factors / Jacobians as the oracle."""
    from laplace_b200 import B200GGN
    from oracle import curvature_oracle as co

    class Cube(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            ctx.save_for_backward(x)
            return x ** 3

        @staticmethod
        def backward(ctx, g):
            (x,) = ctx.saved_tensors
            keep = 1.0 if float(g.abs().sum()) >= 0 else 0.0        # a host read: legal in eager mode, impossible under vmap
            return 3 * x ** 2 * g * keep

    class Act(torch.nn.Module):
        def forward(self, x):
            return Cube.apply(x)

    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(4, 6), Act(), torch.nn.Linear(6, 3))
    X, y = torch.randn(7, 4), torch.randint(3, (7,))
    be = B200GGN(net, "classification")
    _, kron = be.kron(X, y, N=7)
    assert be.last_backward_mode.startswith("loop"), be.last_backward_mode
    _, kf = co.kfac_factors(net.double(), "classification", X.double(), y, N=7)
    Jo, _ = co.jacobians(net, X.double())
    net.float()
    assert max(rel_fro(H, Ho) for F, Fo in zip(kron.kfacs, kf) for H, Ho in zip(F, Fo)) < 1e-5
    assert rel_fro(be.jacobians(X)[0], Jo) < 1e-5
    # and the explicit switch gives the same
    _, kron2 = B200GGN(net, "classification", batched_backward=False).kron(X, y, N=7)
    assert max(rel_fro(H, Ho) for F, Fo in zip(kron2.kfacs, kf) for H, Ho in zip(F, Fo)) < 1e-5


@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_standalone_driver_validation_logdets_and_empty_shard_layout(golden, cpu_kernels, lik):
    """``B200Laplace`` (what runs when the reference package is not importable): argument validation like
    ``BaseLaplace.__init__`` (baselaplace.py:100-140), ``log_det_posterior_precision`` of the three structures against the
    dense matrix, the zero curvature of an empty shard laid out like a fitted one, ``__call__`` contract."""
    from laplace_b200.posterior import B200Laplace

    model, X, y, _ = load(golden, "conv", lik, dtype=torch.float32)
    with pytest.raises(ValueError, match="likelihood"):
        B200Laplace(model, "poisson")
    with pytest.raises(ValueError, match="unsupported"):
        B200Laplace(model, lik, "subnetwork", "full")
    with pytest.raises(ValueError, match="unsupported"):
        B200Laplace(model, lik, "all", "lowrank")
    if lik == "classification":
        with pytest.raises(ValueError, match="Sigma noise"):
            B200Laplace(model, lik, sigma_noise=0.5)
    loader = DataLoader(TensorDataset(X, y), batch_size=5)
    las = {s: B200Laplace(model, lik, "all", s, prior_precision=0.6).fit(loader) for s in ("kron", "full", "diag")}
    dense = {"kron": las["kron"].posterior_precision.to_matrix(), "full": las["full"].posterior_precision,
             "diag": torch.diag(las["diag"].posterior_precision)}
    for s, la in las.items():
        assert torch.allclose(la.log_det_posterior_precision.double(), torch.logdet(dense[s].double()), rtol=1e-4), s
        Z = la.zero_curvature()
        H = la.H_facs if s == "kron" else la.H
        if s == "kron":
            assert Z.dims() == H.dims() and float(Z._flat.abs().sum()) == 0
        else:
            assert Z.shape == H.shape and float(Z.abs().sum()) == 0
        out = la(X)
        if lik == "classification":
            assert out.shape == (len(X), 2) and torch.allclose(out.sum(-1), torch.ones(len(X)), atol=1e-5)
            with pytest.raises(ValueError, match="probit"):
                la(X, link_approx="mc")
        else:
            assert out[0].shape == (len(X), 2) and out[1].shape == (len(X), 2, 2)
        with pytest.raises(ValueError, match="GLM"):
            la(X, pred_type="nn")
    # the bias blocks of fully connected layers are exact in KFAC (no weight sharing): equal to the full GGN's diagonal block
    full, kron = las["full"].H, las["kron"].H_facs.to_matrix()
    linear_biases = {id(m.bias) for m in model.modules() if isinstance(m, torch.nn.Linear)}
    off = 0
    for p in model.parameters():
        n = p.numel()
        if id(p) in linear_biases:
            assert rel_fro(kron[off:off + n, off:off + n], full[off:off + n, off:off + n]) < 1e-4
        off += n


def test_batched_mid_size_route_host_logic(monkeypatch):
    """``_symeig_mid_batched`` (one ``cusolverDnXsyevBatched`` call per size class for live blocks <= 513 rows): grouping,
    dead-coordinate compaction, the negative border that pads a block to its class, scatter of the results -- with the
    library call replaced by ``torch.linalg.eigh`` -- and its three ways out (single-member class, per-matrix ``info != 0``,
    library error), which must hand the factors back for the one-by-one route."""
    from laplace_b200 import _cusolver
    from laplace_b200 import matrix as mx

    def fake_solver(batch):
        W, Q = torch.linalg.eigh(batch.double())
        return W.float(), Q.float(), torch.zeros(len(batch), dtype=torch.int32)

    monkeypatch.setattr(_cusolver, "available", lambda: True)
    monkeypatch.setattr(_cusolver, "syev_batched", fake_solver)
    monkeypatch.setattr(mx, "BATCHED_MID_CLASSES", (16, 40))

    def factor(n, live, seed):
        g = torch.Generator().manual_seed(seed)
        idx = torch.randperm(n, generator=g)[:live].sort().values
        X = torch.randn(3 * live, live, generator=g)
        H = torch.zeros(n, n)
        H[idx.unsqueeze(1), idx.unsqueeze(0)] = X.t() @ X
        return H

    # (n, live): two in class 16 (one full, one compacted from 60 rows), three in class 40, one beyond every class, one
    # alone in ... no class of its own (it joins class 40), one all-dead
    specs = [(12, 12), (60, 9), (40, 40), (33, 33), (90, 20), (70, 70), (25, 0)]
    Hs = [factor(n, lv, s) for s, (n, lv) in enumerate(specs)]
    items = sorted([(i, 0, H) for i, H in enumerate(Hs)], key=lambda t: -t[2].shape[0])
    live = [int((H.diagonal() != 0).sum()) for _, _, H in items]
    eigvals = [[None] for _ in Hs]
    eigvecs = [[None] for _ in Hs]
    rest, rest_live = mx._symeig_mid_batched(items, live, eigvals, eigvecs)
    assert sorted(it[0] for it in rest) == [5, 6] and len(rest_live) == 2          # 70 live rows: too large; 0 live rows: nothing to solve
    for i, (n, lv) in enumerate(specs):
        if i in (5, 6):
            assert eigvals[i][0] is None
            continue
        L, W = eigvals[i][0], eigvecs[i][0]
        assert L.shape == (n,) and W.shape == (n, n) and torch.all(L[1:] >= L[:-1]) and torch.all(L >= 0)
        assert int((L == 0).sum()) >= n - lv
        assert torch.allclose(W.t() @ W, torch.eye(n), atol=1e-4)
        assert rel_fro(W @ torch.diag(L) @ W.t(), Hs[i]) < 1e-5
    # a class with a single member is left to the one-by-one route
    eigvals2, eigvecs2 = [[None]], [[None]]
    rest, _ = mx._symeig_mid_batched([(0, 0, Hs[0])], [12], eigvals2, eigvecs2)
    assert len(rest) == 1 and eigvals2[0][0] is None
    # info != 0 for one matrix of a batch: that one is handed back, the other solved
    monkeypatch.setattr(_cusolver, "syev_batched", lambda b: fake_solver(b)[:2] + (torch.tensor([0, 3], dtype=torch.int32),))
    ev, evec = [[None], [None]], [[None], [None]]
    rest, _ = mx._symeig_mid_batched([(0, 0, Hs[2]), (1, 0, Hs[3])], None, ev, evec)
    assert [it[0] for it in rest] == [1] and ev[0][0] is not None and ev[1][0] is None
    # the library raising: a warning, everything handed back
    def broken(batch):
        raise RuntimeError("no such entry point")

    monkeypatch.setattr(_cusolver, "syev_batched", broken)
    ev, evec = [[None], [None]], [[None], [None]]
    with pytest.warns(UserWarning, match="batched eigensolver unavailable"):
        rest, _ = mx._symeig_mid_batched([(0, 0, Hs[2]), (1, 0, Hs[3])], None, ev, evec)
    assert len(rest) == 2 and ev[0][0] is None
    # solver absent altogether: untouched
    monkeypatch.setattr(_cusolver, "available", lambda: False)
    same, same_live = mx._symeig_mid_batched(items, live, eigvals, eigvecs)
    assert same is items and same_live is live


def test_results_do_not_depend_on_the_process_default_dtype(golden, cpu_kernels):
    """``torch.set_default_dtype(torch.float64)`` (what the reference's own test modules do at import time) must not change
    what a float32 model gets back: last-layer full / diagonal curvature (the ones column of ``[phi; 1]``), the empty-shard
    curvature of the stand-alone driver."""
    from laplace_b200 import B200GGN
    from laplace_b200.posterior import B200Laplace, LastLayerModel

    model, X, y, _ = load(golden, "mlp", "classification", dtype=torch.float32)

    def run():
        be = B200GGN(LastLayerModel(model), "classification", last_layer=True)
        return be.full(X, y)[1], be.diag(X, y)[1]

    H32, d32 = run()
    keep = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        H64, d64 = run()
        la = B200Laplace(model, "classification", "all", "full")
        Z = la.zero_curvature()
    finally:
        torch.set_default_dtype(keep)
    assert H64.dtype == H32.dtype == torch.float32 and torch.equal(H64, H32) and torch.equal(d64, d32)
    assert Z.dtype == torch.float32
