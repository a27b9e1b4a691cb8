"""Bench-shaped parity (full-width ResNet-18, batch 1024) -- collected LAST so that a failure here cannot hide the
cheap golden / oracle tests under ``pytest -x``.

What is asserted, and against what:

* the path ``bench.py`` times (fused conv->BN->ReLU reverse chains, persistent implicit / strided convolutions,
  im2col-free input factors on a side stream, bf16 hi/lo operands) against the **fp64 oracle at full width**,
  every factor <= 1e-4 rel-fro (reference analogue: tests/test_curv_backends_curvlinops.py:207-238 at toy size);
* the same path against the same backend with all of that switched off and the batch split in two (batch
  additivity across code paths);
* reference invariants at a size the oracle cannot reach cheaply (7x normalisation, symmetry, PSD).

The inputs are restricted to samples whose ReLU pre-activations on the small feature maps keep a margin from zero
(``tests/kinks.py``): a unit inside the rounding error of an fp32 forward pass has no defined mask in ANY fp32
implementation, and one flipped unit on a 2x2 map moves that layer's B factor by 4e-5 at this batch size.  Round 1's
version of the second test compared whole vs split batches on unfiltered samples and measured exactly that (1.0e-4 on
``layers.4.conv1`` B, 3e-7 on every other factor, identical run to run; profiles/r02_kfac_bisect.md)."""
import pytest
import torch

from laplace_b200 import B200GGN, models
from oracle import curvature_oracle as co
from tests.fixtures import rel_fro
from tests.kinks import safe_samples

pytestmark = pytest.mark.gpu
DEV = "cuda"
FACTOR_TOL = 1e-4
B = 1024


def _names(be):
    out = []
    for L in be._plan():
        if L.has_w:
            out += [f"{L.name}.B[{L.d_out}]", f"{L.name}.A[{L.d_in}]"]
        if L.has_b:
            out += [f"{L.name}.bias[{L.d_out}]"]
    return out


@pytest.fixture(scope="module")
def bench_batch():
    """1024 margin-safe samples, the fp64 oracle factors on them, and the bench path's factors."""
    torch.set_num_threads(min(32, torch.get_num_threads() or 32))
    md = models.make("resnet18").double()
    torch.manual_seed(3)
    Xc, yc = torch.randn(3072, 3, 32, 32, dtype=torch.float64), torch.randint(10, (3072,))
    keep = safe_samples(md, Xc, B, tol=3e-5)   # measured forward deviation from fp64: <= 2.5e-5 of the rms (profiles/r02_kfac_bisect.md)
    Xc, yc = Xc[keep], yc[keep]
    ref = None
    for i in range(0, B, 128):
        _, kf = co.kfac_factors(md, "classification", Xc[i:i + 128], yc[i:i + 128], N=50000)
        ref = kf if ref is None else [[p + q for p, q in zip(Fa, Fb)] for Fa, Fb in zip(ref, kf)]
    model = models.make("resnet18").to(DEV)
    X, y = Xc.float().to(DEV), yc.to(DEV)
    be = B200GGN(model, "classification", precision="bf16x3")
    _, k1 = be.kron(X, y, N=50000)
    assert be._fused and be.last_backward_mode == "batched" and be.overlap_factors
    return model, X, y, ref, be, k1


def test_bench_path_vs_fp64_oracle_full_width(bench_batch):
    model, X, y, ref, be, k1 = bench_batch
    errs = [rel_fro(h.cpu(), r) for F, Fo in zip(k1.kfacs, ref) for h, r in zip(F, Fo)]
    report = ", ".join(f"{n}: {e:.1e}" for n, e in zip(_names(be), errs) if e > 0.3 * FACTOR_TOL)
    assert max(errs) < FACTOR_TOL, report
    for F in k1.kfacs:
        for a in F:
            assert torch.isfinite(a).all() and rel_fro(a, a.t()) < 1e-5 and float(a.diagonal().min()) >= 0


def test_bench_path_agrees_with_unfused_explicit_split_batches(bench_batch):
    """Fused / implicit / overlapped whole batch == unfused, patch-row SYRKs, no side stream, two half batches."""
    from laplace_b200 import kernels as K

    model, X, y, ref, be, k1 = bench_batch
    ok = K.conv_patches_ok
    K.conv_patches_ok = lambda *a: False
    try:
        be2 = B200GGN(model, "classification", precision="bf16x3", fuse_elementwise=False)
        be2.overlap_factors = False
        _, ka = be2.kron(X[:B // 2], y[:B // 2], N=50000)
        _, kb = be2.kron(X[B // 2:], y[B // 2:], N=50000)
    finally:
        K.conv_patches_ok = ok
    assert not be2._fused
    k2 = ka + kb
    errs = [rel_fro(a, b) for F1, F2 in zip(k1.kfacs, k2.kfacs) for a, b in zip(F1, F2)]
    report = ", ".join(f"{n}: {e:.1e}" for n, e in zip(_names(be), errs) if e > 1e-5)
    assert max(errs) < 3e-5, report
    # and the unfused / explicit path itself against the fp64 oracle
    errs2 = [rel_fro(h.cpu(), r) for F, Fo in zip(k2.kfacs, ref) for h, r in zip(F, Fo)]
    assert max(errs2) < FACTOR_TOL, max(errs2)


def test_auto_precision_policy_vs_fp64_oracle(bench_batch):
    """``precision="auto"`` (what bench.py runs): input factors summed over >= max(16 384, 128 * d_in) sample rows use ONE
    fp16 product (``backend.A_SINGLE_PRODUCT_*``; at this batch: the stem and the 1x1 downsample convolutions), everything
    else three.  Every factor stays within the 1e-4 gate of the fp64 oracle, every input factor within 3e-5 (the policy's
    white-input bound), and with the threshold forced down to one row per feature the post-ReLU input factors still do."""
    from laplace_b200 import backend as bk

    model, X, y, ref, be, k1 = bench_batch
    keep = bk.A_SINGLE_PRODUCT_ROWS_PER_DIM
    for per_dim in (keep, 1):
        bk.A_SINGLE_PRODUCT_ROWS_PER_DIM = per_dim
        try:
            be_auto = B200GGN(model, "classification", precision="auto")
            _, ka = be_auto.kron(X, y, N=50000)
        finally:
            bk.A_SINGLE_PRODUCT_ROWS_PER_DIM = keep
        names = _names(be_auto)
        errs = [rel_fro(h.cpu(), r) for F, Fo in zip(ka.kfacs, ref) for h, r in zip(F, Fo)]
        assert max(errs) < FACTOR_TOL, ", ".join(f"{n}: {e:.1e}" for n, e in zip(names, errs) if e > 0.3 * FACTOR_TOL)
        for n, e in zip(names, errs):
            if ".A[" in n:
                assert e < 3e-5, (per_dim, n, e)


def test_kfac_invariants_at_scale():
    """Reference invariants (tests/test_curv_backends_curvlinops.py:207-333) on the full-width ResNet-18 shape: batch
    additivity, 7x normalisation, symmetry / PSD of every factor."""
    model = models.make("resnet18").to(DEV)
    torch.manual_seed(2)
    X, y = torch.randn(96, 3, 32, 32, device=DEV), torch.randint(10, (96,), device=DEV)
    be = B200GGN(model, "classification", precision="bf16x3")
    _, whole = be.kron(X, y, N=96)
    _, a = be.kron(X[:40], y[:40], N=96)
    _, b = be.kron(X[40:], y[40:], N=96)
    parts = a + b
    for Fw, Fp in zip(whole.kfacs, parts.kfacs):
        for hw, hp in zip(Fw, Fp):
            assert rel_fro(hp, hw) < 1e-4
            assert rel_fro(hw, hw.t()) < 1e-6 and float(hw.diagonal().min()) >= 0
    _, k7 = be.kron(X[:16].repeat(7, 1, 1, 1), y[:16].repeat(7), N=7 * 16)
    _, k1 = be.kron(X[:16], y[:16], N=16)
    assert rel_fro(k7.diag(), 7 * k1.diag()) < 1e-4
