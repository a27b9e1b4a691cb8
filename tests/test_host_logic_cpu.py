"""Host logic of the B200 backend on CPU (kernels replaced by the tests/cpu_kernels.py emulation):
layer plan, hook capture, curvature columns, scaling conventions, Kron assembly -- checked against the
CPU oracle and the golden vectors of the unmodified reference."""
import pytest
import torch

from laplace_b200 import B200EF, B200GGN, B200Kron
from oracle import curvature_oracle as co
from oracle import kron_oracle as ko
from tests.fixtures import load, rel_fro

CASES = [(k, l) for k in ("mlp", "conv") for l in ("classification", "regression")]


@pytest.mark.parametrize("kind,lik", CASES)
@pytest.mark.parametrize("batched", [True, False])
def test_jacobians_full_diag(golden, cpu_kernels, kind, lik, batched):
    model, X, y, rec = load(golden, kind, lik)
    be = B200GGN(model, lik, batched_backward=batched)
    Js, f = be.jacobians(X)
    assert Js.dtype == torch.float64 and rel_fro(Js, rec["Js"]) < 1e-6
    assert torch.allclose(f, rec["f"], atol=1e-6)
    loss, H = be.full(X, y)
    assert torch.allclose(loss, rec["ggn_loss"], rtol=1e-6)
    assert rel_fro(H, rec["ggn_full"]) < 1e-5
    loss, d = be.diag(X, y)
    assert rel_fro(d, rec["ggn_diag"]) < 1e-5
    ef = B200EF(model, lik, batched_backward=batched)
    Gs, gl = ef.gradients(X, y)
    assert rel_fro(Gs, rec["Gs"]) < 1e-6 and torch.allclose(gl, rec["grad_loss"], rtol=1e-6)
    loss, Hef = ef.full(X, y)
    assert torch.allclose(loss, rec["ef_loss"], rtol=1e-6)
    if "ef_full" in rec:
        assert rel_fro(Hef, rec["ef_full"]) < 1e-5
    _, de = ef.diag(X, y)
    assert rel_fro(de, rec["ef_diag"]) < 1e-5


@pytest.mark.parametrize("kind,lik", CASES)
@pytest.mark.parametrize("approx", ["expand", "reduce"])
def test_kron_matches_oracle(golden, cpu_kernels, kind, lik, approx):
    model, X, y, _ = load(golden, kind, lik)
    N = 3 * len(X)
    loss_o, kf_o = co.kfac_factors(model, lik, X, y, N=N, kfac_approx=approx)
    loss, kron = B200GGN(model, lik).kron(X, y, N=N, kfac_approx=approx)
    assert isinstance(kron, B200Kron) and torch.allclose(loss, loss_o, rtol=1e-6)
    assert len(kron.kfacs) == len(kf_o)
    for F, Fo in zip(kron.kfacs, kf_o):
        assert len(F) == len(Fo)
        for H, Ho in zip(F, Fo):
            assert H.dtype == torch.float64 and rel_fro(H, Ho) < 1e-5
    loss_o, kf_o = co.kfac_factors(model, lik, X, y, N=N, fisher="empirical", kfac_approx=approx)
    loss, kron = B200EF(model, lik).kron(X, y, N=N, kfac_approx=approx)
    assert torch.allclose(loss, loss_o, rtol=1e-6)
    for F, Fo in zip(kron.kfacs, kf_o):
        for H, Ho in zip(F, Fo):
            assert rel_fro(H, Ho) < 1e-5


def test_kron_frozen_params(golden, cpu_kernels):
    """Blocks are emitted only for trainable parameters, in parameters() order (reference fixture:
    tests/test_subset_params.py:20-33)."""
    model, X, y, _ = load(golden, "mlp", "classification")
    model[0].weight.requires_grad_(False)
    _, kron = B200GGN(model, "classification").kron(X, y, N=len(X))
    assert [[h.shape[0] for h in F] for F in kron.kfacs] == [[20], [2, 20], [2]]
    _, kf_o = co.kfac_factors(model, "classification", X, y, N=len(X))
    for F, Fo in zip(kron.kfacs, kf_o):
        for H, Ho in zip(F, Fo):
            assert rel_fro(H, Ho) < 1e-5


def test_kron_mc_statistics(golden, cpu_kernels):
    model, X, y, _ = load(golden, "mlp", "classification")
    _, exact = co.kfac_factors(model, "classification", X, y, N=len(X))
    torch.manual_seed(0)
    be = B200GGN(model, "classification", stochastic=True)
    errs = []
    for s in (1, 200):
        _, k = be.kron(X, y, N=len(X), mc_samples=s)
        errs.append(float((k.to_matrix() - co.kfacs_to_matrix(exact)).norm()))
    assert errs[1] < 0.5 * errs[0]


def test_unsupported_module_raises(cpu_kernels):
    model = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.LayerNorm(4), torch.nn.Linear(4, 2))
    with pytest.raises(ValueError):
        B200GGN(model, "classification").kron(torch.randn(2, 3), torch.tensor([0, 1]), N=2)


def test_kron_container_algebra(golden, cpu_kernels):
    rec = golden["kron_algebra"]
    kfacs, W = rec["kfacs"], rec["W"]
    kron = B200Kron([[H.clone() for H in F] for F in kfacs])
    two = kron + kron
    assert all(torch.allclose(a, 2 * b) for Fa, Fb in zip(two.kfacs, kfacs) for a, b in zip(Fa, Fb))
    two += kron
    assert all(torch.allclose(a, 3 * b) for Fa, Fb in zip(two.kfacs, kfacs) for a, b in zip(Fa, Fb))
    scaled = kron * 0.25
    assert torch.allclose(scaled.kfacs[0][0], 0.5 * kfacs[0][0]) and torch.allclose(scaled.kfacs[1][0], 0.25 * kfacs[1][0])
    kd = kron.decompose()
    for a, b in zip(kd.eigenvalues, rec["plain_eigvals"]):
        for x, y in zip(a, b):
            assert torch.allclose(x, y, atol=1e-5)
    for name in ("scalar", "layer"):
        P = kd * 1.7 + rec[f"plain_{name}_delta"]
        assert torch.allclose(P.inv_square_form(W), rec[f"plain_{name}_isf"], rtol=1e-4)
        assert torch.allclose(P.logdet(), rec[f"plain_{name}_logdet"], rtol=1e-6)
        assert torch.allclose(P.bmm(W, exponent=-0.5), rec[f"plain_{name}_bmm_m05"], rtol=1e-3, atol=1e-5)
        assert P.damping is False  # reference quirk: * and + drop damping


def test_structured_predictive_matches_dense(golden, cpu_kernels):
    """inv_square_form through the per-layer factors == dense reference algebra (fp32 model)."""
    for kind in ("mlp", "conv"):
        model, X, y, _ = load(golden, kind, "classification", dtype=torch.float32)
        be = B200GGN(model, "classification")
        _, kron = be.kron(X, y, N=len(X))
        P = kron.decompose() * 1.0 + torch.tensor(0.7)
        Js, f = be.jacobians(X)
        assert hasattr(Js, "_lpb_factors")
        fv = P.inv_square_form(Js)
        Qs = [[q.double() for q in Q] for Q in P.eigenvectors]
        ls = [[l.double() for l in L] for L in P.eigenvalues]
        ref = ko.kron_inv_square_form(Qs, ls, torch.tensor(0.7, dtype=torch.float64), Js.double().clone())
        assert torch.allclose(fv.double(), ref, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("name,kw", [("resnet18", {"width": 8}), ("resnet18", {"width": 8, "cifar_stem": True}),
                                     ("wrn28_10", {"depth": 10, "widen": 1}),
                                     ("vit_b16", {"image": 32, "patch": 8, "dim": 32, "depth": 2, "heads": 4, "mlp_dim": 64})])
def test_convolution_engine_paths_on_model_zoo(cpu_kernels, monkeypatch, name, kw):
    """Full engine (implicit/explicit convolutions, Linear layers, frozen-BN affine, ReLU and MaxPool custom reverse
    passes, functorch-batched columns, factor SYRKs from the stashed rows) == oracle, on reduced model shapes."""
    from laplace_b200 import conv_engine, models

    from laplace_b200 import kernels as K

    monkeypatch.setattr(conv_engine, "ELEMENTWISE_MIN_BATCH", 0)
    pool_packs = []
    orig_pack = K.maxpool2d_bwd_pack
    monkeypatch.setattr(K, "maxpool2d_bwd_pack", lambda *a, **k: (pool_packs.append(1), orig_pack(*a, **k))[1])
    model = models.make(name, **kw)
    torch.manual_seed(0)
    X, y = torch.randn(4, 3, 32, 32), torch.randint(10, (4,))
    be = B200GGN(model, "classification")
    _, kron = be.kron(X, y, N=4)
    assert be.last_backward_mode == "batched"
    # torchvision stem (conv -> BN -> ReLU -> max-pool): the pool's reverse map emits the chain's operand rows directly
    assert bool(pool_packs) == (name == "resnet18" and not kw.get("cifar_stem"))
    # conv -> frozen BN -> ReLU chains run as one fused reverse node; the reduced WideResNet adds the raw stem output
    # to a residual (a second consumer of a fused intermediate): detected, repeated unfused, and remembered
    assert be.fuse_elementwise == (name != "wrn28_10")
    assert be._fused == (name != "wrn28_10")
    _, kf = co.kfac_factors(model, "classification", X, y, N=4)
    worst = max(rel_fro(H, Ho) for F, Fo in zip(kron.kfacs, kf) for H, Ho in zip(F, Fo))
    assert worst < 1e-5, worst
    # empirical Fisher: a single curvature column through the same (fused) chains
    ef = B200EF(model, "classification")
    _, kron_ef = ef.kron(X, y, N=4)
    _, kf_ef = co.kfac_factors(model, "classification", X, y, N=4, fisher="empirical")
    worst_ef = max(rel_fro(H, Ho) for F, Fo in zip(kron_ef.kfacs, kf_ef) for H, Ho in zip(F, Fo))
    assert worst_ef < 1e-5, worst_ef
    # diag / jacobian-based paths run through the same engine
    _, d = be.diag(X, y)
    Js, f = co.jacobians(model, X)
    _, dref = co.ggn_diag(Js, f, y, "classification")
    assert rel_fro(d, dref) < 1e-4


def test_mapping_inputs_and_dtype_propagation(golden, cpu_kernels):
    """HF-style MutableMapping batches (baselaplace.py:969-974, curvlinops.py:84-85) and model-dtype outputs
    (reference tests/test_baselaplace.py:895-934)."""
    class DictModel(torch.nn.Module):
        def __init__(self, net):
            super().__init__()
            self.net = net

        def forward(self, batch):
            return self.net(batch["input_ids"])

    model, X, y, _ = load(golden, "mlp", "classification", dtype=torch.float32)
    ref_loss, ref = B200GGN(model, "classification").kron(X, y, N=len(X))
    wrapped = DictModel(model)
    loss, kron = B200GGN(wrapped, "classification").kron({"input_ids": X, "labels": y}, y, N=len(X))
    assert torch.allclose(loss, ref_loss)
    for F, Fr in zip(kron.kfacs, ref.kfacs):
        for H, Hr in zip(F, Fr):
            assert H.dtype == torch.float32 and torch.allclose(H, Hr)
    m64, X64, y64, _ = load(golden, "mlp", "regression", dtype=torch.float64)
    be = B200GGN(m64, "regression")
    assert be.jacobians(X64)[0].dtype == torch.float64 and be.full(X64, y64)[1].dtype == torch.float64
    assert be.diag(X64, y64)[1].dtype == torch.float64 and be.kron(X64, y64, N=10)[1].kfacs[0][0].dtype == torch.float64


def test_implicit_patch_factor_path(cpu_kernels, monkeypatch):
    """64-channel stride-1 convolutions take the im2col-free A-factor path (K.syrk_conv_patches on the forward's NHWC
    rows): same factors as the oracle, in parameter order."""
    from laplace_b200 import conv_engine, kernels as K

    calls = []
    orig = K.syrk_conv_patches
    monkeypatch.setattr(K, "syrk_conv_patches", lambda *a, **k: (calls.append(1), orig(*a, **k))[1])
    monkeypatch.setattr(conv_engine, "ELEMENTWISE_MIN_BATCH", 0)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 64, 3, 1, 1), torch.nn.ReLU(), torch.nn.Conv2d(64, 64, 3, 1, 1, bias=False),
                                torch.nn.ReLU(), torch.nn.Conv2d(64, 16, 1), torch.nn.AdaptiveAvgPool2d(1),
                                torch.nn.Flatten(), torch.nn.Linear(16, 5)).eval()
    X, y = torch.randn(6, 3, 8, 8), torch.randint(5, (6,))
    _, kron = B200GGN(model, "classification").kron(X, y, N=18)
    assert len(calls) == 2          # the 3x3 64->64 and the 1x1 64->16 convolution
    _, kf = co.kfac_factors(model, "classification", X, y, N=18)
    worst = max(rel_fro(H, Ho) for F, Fo in zip(kron.kfacs, kf) for H, Ho in zip(F, Fo))
    assert worst < 1e-5, worst


def test_diag_tensor_core_conv_path(cpu_kernels):
    """Stride-1 convolutions on >= 64-pixel images take the tensor-core diagonal (K.diag_conv_sq): diag GGN and diag EF
    equal the oracle's, in parameter order."""
    from laplace_b200 import kernels as K

    calls = []
    orig = K.diag_conv_sq
    K.diag_conv_sq = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3, 1, 1), torch.nn.ReLU(), torch.nn.Conv2d(16, 24, 3, 1, 1, bias=False),
                                    torch.nn.ReLU(), torch.nn.Conv2d(24, 8, 1), torch.nn.AdaptiveAvgPool2d(1),
                                    torch.nn.Flatten(), torch.nn.Linear(8, 5)).eval()
        X, y = torch.randn(6, 3, 8, 8), torch.randint(5, (6,))
        _, d = B200GGN(model, "classification").diag(X, y)
        assert len(calls) == 2      # the 16->24 3x3 and 24->8 1x1 convolutions (the 3-channel stem stays on the SIMT path)
        Js, f = co.jacobians(model, X)
        _, dref = co.ggn_diag(Js, f, y, "classification")
        assert rel_fro(d, dref) < 1e-5
        _, de = B200EF(model, "classification").diag(X, y)
        _, deref = co.ef_diag(Js, f, y, "classification")
        assert rel_fro(de, deref) < 1e-5
    finally:
        K.diag_conv_sq = orig


def test_engine_path_predicates():
    """Which kernel family a convolution is routed to (pure host logic; the kernels' own argument checks mirror these)."""
    from laplace_b200 import conv_engine as ce, kernels as K

    c3 = torch.nn.Conv2d(64, 64, 3, 1, 1)
    c3s2 = torch.nn.Conv2d(64, 128, 3, 2, 1)
    c1s2 = torch.nn.Conv2d(64, 128, 1, 2, 0)
    c7s2 = torch.nn.Conv2d(3, 64, 7, 2, 3)
    # stride-1 "same": whole images tiling 128 rows, or 128 / W image rows of a larger image
    assert all(ce.implicit_ok(c3, h, w) for h, w in [(1, 1), (2, 2), (4, 4), (8, 8), (16, 8), (16, 16), (32, 32), (64, 64)])
    assert not any(ce.implicit_ok(c3, h, w) for h, w in [(3, 3), (6, 6), (12, 12), (6, 48), (7, 7)])
    assert not ce.implicit_ok(c3s2, 8, 8) and not ce.implicit_ok(torch.nn.Conv2d(8, 8, 3, 1, 0), 8, 8)
    # strided reverse passes: input extent = stride x output extent, output grid tiles 128 rows
    assert ce.strided_ok(c3s2, 8, 8) and ce.strided_ok(c1s2, 8, 8) and ce.strided_ok(c7s2, 32, 32) and ce.strided_ok(c3s2, 2, 2)
    assert not ce.strided_ok(c3s2, 5, 5) and not ce.strided_ok(c3, 8, 8) and not ce.strided_ok(c3s2, 12, 12)
    # im2col-free input factors: <= 9 taps, images tiling 64-row chunks, channel padding to 64 wastes <= 1/3
    assert K.conv_patches_ok(64, 8, 8, 3, 3) and K.conv_patches_ok(160, 32, 32, 3, 3) and K.conv_patches_ok(512, 1, 1, 3, 3)
    assert K.conv_patches_ok(48, 8, 8, 1, 1) and not K.conv_patches_ok(80, 8, 8, 3, 3) and not K.conv_patches_ok(3, 32, 32, 3, 3)
    assert not K.conv_patches_ok(64, 8, 8, 5, 5) and not K.conv_patches_ok(64, 6, 6, 3, 3)
    # tensor-core diagonal: whole 64-pixel chunks per image
    assert K.diag_conv_ok(160, 32, 32, 3, 3) and K.diag_conv_ok(64, 8, 8, 3, 3) and K.diag_conv_ok(16, 16, 8, 1, 1)
    assert not K.diag_conv_ok(128, 4, 4, 3, 3) and not K.diag_conv_ok(8, 8, 8, 3, 3) and not K.diag_conv_ok(64, 12, 12, 3, 3)


def test_fit_batches_on_cpu_match_plain_iteration(golden, cpu_kernels):
    """``B200Laplace._device_batches`` (copy-stream prefetch on CUDA) degrades to plain iteration elsewhere and handles
    empty loaders and HF-style mapping batches."""
    from laplace_b200.posterior import B200Laplace

    model, X, y, _ = load(golden, "mlp", "classification", dtype=torch.float32)
    la = B200Laplace(model, "classification", "all", "kron")
    batches = [(X[:4], y[:4]), (X[4:], y[4:])]
    got = list(la._device_batches(batches))
    assert len(got) == 2 and all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(got, batches))
    assert list(la._device_batches([])) == []
    m = {"input_ids": X[:3], "labels": y[:3]}
    (Xm, ym), = list(la._device_batches([m]))
    assert Xm is m and torch.equal(ym, y[:3])


def test_compact_and_bordered_eigendecompositions(cpu_kernels, monkeypatch):
    """``decompose()`` shortcuts for large factors: (a) coordinates with a zero diagonal (kernel positions that only see
    padding) are split off and only the live block goes through ``eigh``; (b) fp32 matrices of 385..512 rows are
    bordered to 513 rows to reach cuSOLVER's faster solver.  Both must return a valid ascending eigendecomposition."""
    from laplace_b200 import matrix as mx

    torch.manual_seed(0)
    n, k = 90, 30
    idx = torch.randperm(n)[:k].sort().values
    X = torch.randn(200, k, dtype=torch.float64)
    H = torch.zeros(n, n, dtype=torch.float64)
    H[idx.unsqueeze(1), idx.unsqueeze(0)] = X.t() @ X
    L, W = mx._symeig_compact(H)
    assert torch.all(L[1:] >= L[:-1]) and torch.all(L[: n - k] == 0)
    assert torch.allclose(W.t() @ W, torch.eye(n, dtype=torch.float64), atol=1e-10)
    assert torch.allclose(W @ torch.diag(L) @ W.t(), H, atol=1e-8 * float(H.abs().max()))
    assert torch.allclose(L, torch.linalg.eigvalsh(H).clamp(min=0), atol=1e-8 * float(H.abs().max()))
    # through B200Kron.decompose: same posterior functionals with and without the shortcut
    A = torch.randn(50, 12, dtype=torch.float64)
    kron = B200Kron([[A.t() @ A, H.clone()]])
    Wt = torch.randn(3, 2, 12 * n, dtype=torch.float64)
    outs = []
    for flag in (True, False):
        monkeypatch.setattr(mx, "COMPACT_DEAD_COORDINATES", flag)
        P = kron.decompose() * 1.0 + torch.tensor(0.3, dtype=torch.float64)
        outs.append((P.inv_square_form(Wt.clone()), P.logdet()))
    # (the rotation GEMMs run in fp32 in both cases)
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-5) and torch.allclose(outs[0][1], outs[1][1], rtol=1e-6)
    # (b) bordered eigh (exercised on CPU through the test switch)
    monkeypatch.setattr(mx, "_PAD_ON_CPU", True)
    Y = torch.randn(700, 400)
    G = (Y.t() @ Y) * 37.0
    Lp, Wp = mx.symeig_large(G)
    Lr = torch.linalg.eigvalsh(G.double())
    assert Lp.shape == (400,) and Wp.shape == (400, 400) and torch.all(Lp[1:] >= Lp[:-1])
    assert torch.allclose(Lp.double(), Lr, rtol=1e-4, atol=1e-4 * float(Lr.max()))
    assert torch.allclose(Wp.t() @ Wp, torch.eye(400), atol=1e-4)
    assert rel_fro(Wp @ torch.diag(Lp) @ Wp.t(), G) < 1e-5


@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_lazy_jacobian_kron_predictive_matches_dense(golden, cpu_kernels, lik):
    """Conv-layer Kron GLM predictive without the dense ``(B, C, P)`` Jacobian: the ``LazyJacobian`` carries
    ``("conv", G_rows, A_rows, T)`` factors; ``inv_square_form`` rotates the ROWS into the eigenbasis and reduces per tile.
    Same variances as the dense route and as the oracle; any other use of the tensor materialises it correctly."""
    from laplace_b200.posterior import B200Laplace
    from laplace_b200.predictive import LazyJacobian
    from torch.utils.data import DataLoader, TensorDataset

    model, X, y, rec = load(golden, "conv", lik, dtype=torch.float32)
    la = B200Laplace(model, lik, "all", "kron", prior_precision=0.7).fit(DataLoader(TensorDataset(X, y), batch_size=5))
    Jd, f = la.backend.jacobians(X)
    assert not isinstance(Jd, LazyJacobian)
    dense_var = la.functional_variance(Jd)
    la.backend.lazy_jacobians = True
    Jl, f2 = la.backend.jacobians(X)
    assert isinstance(Jl, LazyJacobian) and Jl.shape == Jd.shape and Jl.dtype == Jd.dtype
    assert [b[0] for b in Jl._lpb_factors.blocks] == ["conv", "vec", "outer", "vec", "outer", "vec"]
    lazy_var = la.functional_variance(Jl)
    assert Jl._lpb_dense is None, "the structured path must not materialise the Jacobian"
    assert torch.allclose(lazy_var, dense_var, rtol=1e-4, atol=1e-7)
    P = la.posterior_precision
    P.damping = True                                    # damped spectrum through the same kernel
    assert torch.allclose(P.inv_square_form(Jl), P.inv_square_form(Jd.clone()), rtol=1e-4, atol=1e-7)
    # any other operation sees the dense values
    assert torch.allclose(Jl + 0.0, Jd, atol=1e-6) and torch.allclose(Jl.sum(0), Jd.sum(0), atol=1e-5)
    assert rel_fro(Jl.dense().double(), rec["Js"]) < 1e-6


def _gp_case(kind, golden):
    if kind == "wide_conv":       # small maps, wide channels: the structured sum over (t, t') is the cheaper route
        torch.manual_seed(9)
        model = torch.nn.Sequential(torch.nn.Conv2d(8, 16, 3, 1, 1), torch.nn.Tanh(), torch.nn.Conv2d(16, 12, 3, 1, 1), torch.nn.Tanh(),
                                    torch.nn.Flatten(), torch.nn.Linear(48, 3)).double()
        return model, torch.randn(7, 8, 2, 2, dtype=torch.float64)
    model, X, _, _ = load(golden, kind, "classification")
    return model, X


@pytest.mark.parametrize("kind", ["mlp", "conv", "wide_conv"])
@pytest.mark.parametrize("lazy", [False, True])
def test_gp_kernels_match_reference_einsums(golden, cpu_kernels, kind, lazy):
    """the three kernel tensors of ``FunctionalLaplace`` (baselaplace.py:3026-3122) from the backend's
    Jacobian factors, against the reference's einsums on the oracle's dense Jacobians."""
    from laplace_b200 import gp

    model, X = _gp_case(kind, golden)
    J, _ = co.jacobians(model, X)
    P = J.shape[-1]
    na = len(X) // 2
    be = B200GGN(model.float(), "classification")
    be.lazy_jacobians = lazy
    J1, _ = be.jacobians(X[:na].float())
    J2, _ = be.jacobians(X[na:].float())
    Ja, Jb = J[:na], J[na:]
    assert rel_fro(gp.kernel_batch(J1, J2), torch.einsum("ap,bp->ab", Ja.reshape(-1, P), Jb.reshape(-1, P))) < 1e-5
    assert rel_fro(gp.kernel_batch(J1), torch.einsum("ap,bp->ab", Ja.reshape(-1, P), Ja.reshape(-1, P))) < 1e-5
    assert rel_fro(gp.kernel_star(J1), torch.einsum("bcp,bep->bce", Ja, Ja)) < 1e-5
    assert rel_fro(gp.kernel_star(J1, joint=True), torch.einsum("acp,bep->abce", Ja, Ja)) < 1e-5
    assert rel_fro(gp.kernel_batch_star(J1, J2), torch.einsum("bcp,dep->bdce", Ja, Jb)) < 1e-5
    indep = torch.stack([torch.einsum("bp,ep->be", Ja[:, c], Jb[:, c]) for c in range(J.shape[1])], -1)
    assert rel_fro(gp.kernel_batch(J1, J2, independent_outputs=True), indep) < 1e-5
    assert rel_fro(gp.kernel_star(J1, independent_outputs=True), (Ja ** 2).sum(-1)) < 1e-5
    model.double()


def test_fused_pool_conflict_falls_back(cpu_kernels, monkeypatch):
    """A fused conv -> BN -> ReLU chain whose output feeds a fused max-pool AND a second operation: the pool's pre-packed
    rows would miss the second contribution -- detected (the placeholder gradient was replaced by a sum), the batch is
    repeated unfused, the factors match the oracle."""
    from laplace_b200 import conv_engine

    monkeypatch.setattr(conv_engine, "ELEMENTWISE_MIN_BATCH", 0)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 8, 3, 1, 1, bias=False)
            self.bn = torch.nn.BatchNorm2d(8)
            self.relu = torch.nn.ReLU()
            self.pool = torch.nn.MaxPool2d(3, 2, 1)
            self.conv2 = torch.nn.Conv2d(8, 8, 3, 1, 1)
            self.fc = torch.nn.Linear(8, 3)

        def forward(self, x):
            h = self.relu(self.bn(self.conv(x)))
            z = self.pool(h) + torch.nn.functional.avg_pool2d(h, 2)      # second consumer of the chain output
            return self.fc(torch.tanh(self.conv2(z)).mean((2, 3)))

    torch.manual_seed(4)
    net = Net().eval()
    net.bn.running_mean.normal_(0, 0.1), net.bn.running_var.uniform_(0.5, 1.5)
    for p_ in net.bn.parameters():
        p_.requires_grad_(False)
    X, y = torch.randn(6, 3, 8, 8), torch.randint(3, (6,))
    be = B200GGN(net, "classification")
    _, kron = be.kron(X, y, N=6)
    assert be.fuse_elementwise is False and not be._fused
    _, kf = co.kfac_factors(net.double(), "classification", X.double(), y, N=6)
    net.float()
    worst = max(rel_fro(H, Ho) for F, Fo in zip(kron.kfacs, kf) for H, Ho in zip(F, Fo))
    assert worst < 1e-5, worst


def test_token_shared_linear_lazy_predictive_and_gp(cpu_kernels):
    """A transformer's ``nn.Linear`` applied to every token is a weight shared over ``T`` positions, the same structure as a
    convolution (curvlinops KFAC-expand): its ``LazyJacobian`` blocks are ``("conv", G_rows, A_rows, T)``, the Kron
    predictive and the GP kernels run on them without the dense Jacobian and agree with the dense route and the oracle."""
    from laplace_b200 import gp, models
    from laplace_b200.posterior import B200Laplace
    from laplace_b200.predictive import LazyJacobian
    from torch.utils.data import DataLoader, TensorDataset

    model = models.make("vit_b16", image=16, patch=8, dim=16, depth=1, heads=2, mlp_dim=32)
    torch.manual_seed(0)
    X, y = torch.randn(6, 3, 16, 16), torch.randint(10, (6,))
    la = B200Laplace(model, "classification", "all", "kron", prior_precision=0.5).fit(DataLoader(TensorDataset(X, y), batch_size=3))
    Jd, _ = la.backend.jacobians(X[:3])
    dense_var = la.functional_variance(Jd)
    la.backend.lazy_jacobians = True
    Jl, _ = la.backend.jacobians(X[:3])
    assert isinstance(Jl, LazyJacobian)
    kinds = [b[0] for b in Jl._lpb_factors.blocks]
    assert kinds.count("conv") >= 4 and "outer" in kinds          # per-token linears / the classification head
    assert torch.allclose(la.functional_variance(Jl), dense_var, rtol=1e-4, atol=1e-7)
    assert Jl._lpb_dense is None
    Jo, _ = co.jacobians(model.double(), X[:3].double())
    model.float()
    flat = Jo.reshape(-1, Jo.shape[-1])
    assert rel_fro(gp.kernel_batch(Jl), flat @ flat.t()) < 1e-5
    assert rel_fro(gp.kernel_star(Jl), torch.einsum("bcp,bep->bce", Jo, Jo)) < 1e-5
    assert rel_fro(Jl.dense().double(), Jo) < 1e-5


@pytest.mark.parametrize("structure", ["kron", "full", "diag"])
@pytest.mark.parametrize("lik", ["classification", "regression"])
def test_gridsearch_structures_likelihoods_and_mapping_batches(golden, cpu_kernels, structure, lik):
    """``tuning.gridsearch_prior_precision`` on the stand-alone posterior for every structure and both likelihoods: same
    loss curve as a loop that re-derives Jacobians and variances for every grid value (what ``utils.validate`` does,
    utils/utils.py:39-101), tuple batches and mapping batches (targets under ``dict_key_y``) alike."""
    import math

    from laplace_b200.posterior import B200Laplace
    from laplace_b200.tuning import gridsearch_prior_precision
    from torch.utils.data import DataLoader, TensorDataset

    model, X, y, _ = load(golden, "mlp", lik, dtype=torch.float32)
    la = B200Laplace(model, lik, "all", structure).fit(DataLoader(TensorDataset(X, y), batch_size=5))
    vl = [(X[:6], y[:6]), (X[6:], y[6:])]
    grid = torch.logspace(-2, 2, 7)
    best, losses = gridsearch_prior_precision(la, vl, interval=grid, set_result=False)
    brute = []
    for pp in grid:
        la.prior_precision = pp
        tot = 0.0
        for Xb, yb in vl:
            Js, f = la.backend.jacobians(Xb)
            var = la.functional_variance(Js)
            if lik == "regression":
                tot += float(((f - yb) ** 2).sum())
            else:
                kappa = 1.0 / torch.sqrt(1.0 + math.pi / 8.0 * torch.diagonal(var, dim1=1, dim2=2))
                tot += float(torch.nn.functional.nll_loss(torch.softmax(kappa * f, -1).log(), yb, reduction="sum"))
        brute.append(tot / len(X))
    assert torch.allclose(losses, torch.tensor(brute, dtype=torch.float64), rtol=1e-5, atol=1e-7)
    assert float(best) == float(grid[int(torch.tensor(brute).argmin())])

    class DictModel(torch.nn.Module):
        def __init__(self, net):
            super().__init__()
            self.net = net

        def forward(self, batch):
            return self.net(batch["input_ids"])

    la_d = B200Laplace(DictModel(model), lik, "all", structure)

    class Batches(list):
        dataset = range(len(X))

    la_d.fit(Batches({"input_ids": X[i:i + 5], "labels": y[i:i + 5]} for i in (0, 5)))
    _, losses_d = gridsearch_prior_precision(la_d, [{"input_ids": Xb, "labels": yb} for Xb, yb in vl], interval=grid)
    assert torch.allclose(losses_d, losses, rtol=1e-5, atol=1e-7)
