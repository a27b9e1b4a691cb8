"""Convolution engine (forward / backward-data of nn.Conv2d on the tcgen05 GEMM) against fp64 torch convolutions,
kernel by kernel and through autograd (batched reverse pass over curvature columns)."""
import pytest
import torch
import torch.nn.functional as F

from laplace_b200 import B200GGN, conv_engine, kernels as K, models
from tests import cpu_kernels as ck
from tests.fixtures import rel_fro

pytestmark = pytest.mark.gpu
DEV = "cuda"
GEOMS = [(3, 64, 7, 2, 3, 1, 32), (64, 64, 3, 1, 1, 1, 8), (64, 128, 3, 2, 1, 1, 8), (64, 128, 1, 2, 0, 1, 8), (256, 512, 3, 2, 1, 1, 2),
         (512, 512, 3, 1, 1, 1, 1), (5, 7, 3, 1, 2, 2, 9), (4, 6, 2, 2, 0, 1, 5),
         # implicit-GEMM eligible (stride 1, same padding, H*W | 128), incl. ragged channel counts
         (128, 128, 3, 1, 1, 1, 4), (256, 256, 3, 1, 1, 1, 2), (24, 40, 5, 1, 2, 1, 4), (16, 200, 1, 1, 0, 1, 8), (72, 8, 3, 1, 1, 1, 8),
         # implicit GEMM on images larger than one 128-row tile (tiles of 128 / W image rows)
         (16, 32, 3, 1, 1, 1, 32), (64, 64, 3, 1, 1, 1, 16), (3, 16, 3, 1, 1, 1, 32), (8, 24, 5, 1, 2, 1, 64)]


def test_implicit_path_selected():
    m = torch.nn.Conv2d(64, 64, 3, 1, 1)
    assert conv_engine.implicit_ok(m, 8, 8) and conv_engine.implicit_ok(m, 1, 1) and conv_engine.implicit_ok(m, 16, 16)
    assert conv_engine.implicit_ok(m, 32, 32) and not conv_engine.implicit_ok(m, 12, 12) and not conv_engine.implicit_ok(m, 6, 48)
    assert not conv_engine.implicit_ok(torch.nn.Conv2d(64, 64, 3, 2, 1), 8, 8)


def _dense(p, cols):
    return p.hi[:, :cols].float().cpu() + (p.lo[:, :cols].float().cpu() if p.lo is not None else 0)


@pytest.mark.parametrize("geom", GEOMS)
def test_engine_operand_kernels(geom):
    cin, cout, k, s, p, d, hw = geom
    torch.manual_seed(0)
    mod = torch.nn.Conv2d(cin, cout, k, s, p, dilation=d)
    x = torch.randn(6, cin, hw, hw)
    ref = ck.pack_conv_rows(x, mod, K.F32)
    out = K.pack_conv_rows(x.to(DEV), mod, K.BF16X3)
    assert rel_fro(_dense(out, ref.K), ref.hi[:, :ref.K]) < 2e-5
    o = mod(x)
    g = torch.randn(9, cout, o.shape[2] * o.shape[3])
    ref = ck.pack_nchw_rows(g, K.F32)
    out = K.pack_nchw_rows(g.to(DEV), K.BF16X3)
    assert rel_fro(_dense(out, ref.K), ref.hi[:, :ref.K]) < 2e-5
    w = torch.randn(cout, cin * k * k)
    assert rel_fro(_dense(K.pack_cast(w.to(DEV), K.BF16X3), w.shape[1]), w) < 2e-5
    T = o.shape[2] * o.shape[3]
    Dc = torch.randn(cin * k * k, 9 * T)
    refc = ck.col2im(Dc, (9, cin, hw, hw), mod)
    outc = K.col2im(Dc.to(DEV), (9, cin, hw, hw), mod)
    assert rel_fro(outc.cpu(), refc) < 1e-6
    # channels-last form: rows (q,t), tap-major columns (kh,kw,ci)
    Dn = Dc.reshape(cin, k * k, 9, T).permute(2, 3, 1, 0).reshape(9 * T, k * k * cin).contiguous()
    outn = K.col2im_nhwc(Dn.to(DEV), (9, cin, hw, hw), mod)
    assert outn.shape == refc.shape and outn.permute(0, 2, 3, 1).is_contiguous()
    assert rel_fro(outn.cpu(), refc) < 1e-6


@pytest.mark.parametrize("geom", GEOMS)
def test_conv_forward_backward_vs_fp64(geom):
    cin, cout, k, s, p, d, hw = geom
    torch.manual_seed(1)
    mod = torch.nn.Conv2d(cin, cout, k, s, p, dilation=d).to(DEV)
    x = torch.randn(33, cin, hw, hw, device=DEV)
    ref = F.conv2d(x.double(), mod.weight.double(), mod.bias.double(), s, p, d)
    out = conv_engine.conv_forward(x, mod)
    assert out.shape == ref.shape and rel_fro(out, ref) < 5e-6      # forward: fp16 hi/lo operands (22 bits)
    out_cl = conv_engine.conv_forward(x.contiguous(memory_format=torch.channels_last), mod)
    assert rel_fro(out_cl, ref) < 5e-6
    g = torch.randn(65, *ref.shape[1:], device=DEV)
    gref = torch.nn.grad.conv2d_input((65, cin, hw, hw), mod.weight.double(), g.double(), s, p, d)
    gin = conv_engine.conv_backward_data(g, mod, (65, cin, hw, hw))
    assert gin.shape == gref.shape and rel_fro(gin, gref) < 2e-5
    conv_engine.PASS_ID[0] += 1   # a second reverse pass over the same layer (the backend bumps this per autograd.grad)
    gin_cl = conv_engine.conv_backward_data(g.contiguous(memory_format=torch.channels_last), mod, (65, cin, hw, hw))
    assert rel_fro(gin_cl, gref) < 2e-5


@pytest.mark.parametrize("geom", [(64, 64, 3, 1, 1, 1, 8), (128, 64, 3, 1, 1, 1, 4), (64, 128, 3, 2, 1, 1, 16), (256, 256, 3, 1, 1, 1, 2)])
def test_two_product_backward_data(geom):
    """``LEAN_BWD_MIN_ROWS`` route (opt-in, off by default -- see conv_engine.py): the gradient rows enter the backward-data convolution
    as their bf16 hi half only, the weights stay hi + lo.  The result equals the exact convolution of the ROUNDED gradient
    (5e-6: weights are not rounded) and deviates from the unrounded one by the bf16 rounding of the rows (~2^-9, unbiased:
    the mean signed deviation is two orders of magnitude smaller)."""
    cin, cout, k, s, p, d, hw = geom
    torch.manual_seed(2)
    mod = torch.nn.Conv2d(cin, cout, k, s, p, dilation=d, bias=False).to(DEV)
    oh = (hw + 2 * p - k) // s + 1
    g = torch.randn(130, cout, oh, oh, device=DEV)
    shape = (130, cin, hw, hw)
    gref = torch.nn.grad.conv2d_input(shape, mod.weight.double(), g.double(), s, p, d)
    ground = torch.nn.grad.conv2d_input(shape, mod.weight.double(), g.bfloat16().double(), s, p, d)
    keep = conv_engine.LEAN_BWD_MIN_ROWS
    conv_engine.LEAN_BWD_MIN_ROWS = 1
    try:
        conv_engine.PASS_ID[0] += 1
        gin = conv_engine.conv_backward_data(g, mod, shape)
    finally:
        conv_engine.LEAN_BWD_MIN_ROWS = keep
    assert rel_fro(gin, ground) < 5e-6
    assert 1e-4 < rel_fro(gin, gref) < 4e-3
    assert abs(float((gin.double() - gref).mean() / gref.abs().mean())) < 5e-5
    conv_engine.PASS_ID[0] += 1
    full = conv_engine.conv_backward_data(g, mod, shape)
    assert rel_fro(full, gref) < 2e-5


def test_store_mode_gemm_nonsymmetric():
    torch.manual_seed(2)
    A, B = torch.randn(576, 64), torch.randn(40000, 64)
    pa, pb = K.pack_cast(A.to(DEV), K.BF16X3), K.pack_cast(B.to(DEV), K.BF16X3)
    out = torch.full((576, 40000), 5.0, device=DEV)
    K.gemm_nt(pa, pb, out, 1.0, accumulate=False)
    assert rel_fro(out.cpu(), A.double() @ B.double().t()) < 2e-5


@pytest.mark.parametrize("name,kw", [("resnet18", {"width": 16}), ("wrn28_10", {"depth": 10, "widen": 2})])
def test_layer_gradients_match_fp64_autograd(name, kw):
    """The batched reverse pass through the engine reproduces fp64 autograd gradients at every layer output."""
    model = models.make(name, **kw).to(DEV)
    torch.manual_seed(3)
    X = torch.randn(16, 3, 32, 32, device=DEV)
    be = B200GGN(model, "classification")
    f = be._forward(X)
    cols = be._hessian_sqrt_cols(f.detach())
    grads = be._backward(f, cols)
    assert be.last_backward_mode == "batched"
    md = models.make(name, **kw).double().to(DEV)
    md.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    be64 = B200GGN(md, "classification", conv_engine=False)
    f64 = be64._forward(X.double())
    g64 = be64._backward(f64, cols.double())
    assert rel_fro(f, f64) < 1e-5
    worst = max(rel_fro(a, b) for a, b in zip(grads, g64))
    assert worst < 5e-5, worst


@pytest.mark.parametrize("geom", [(64, 8, 8, 3, 70), (128, 4, 4, 3, 130), (64, 2, 2, 3, 33), (256, 1, 1, 3, 200), (64, 16, 16, 3, 5),
                                  (64, 8, 8, 1, 40), (128, 16, 8, 3, 9), (64, 32, 32, 3, 3),
                                  # channel counts padded to 64 per tap (zero-filled by the TMA unit)
                                  (160, 8, 8, 3, 20), (96, 4, 4, 3, 50), (48, 8, 8, 1, 10),
                                  # degenerate maps: kernel positions that only see padding are dropped (1 resp. 3 live taps)
                                  (128, 1, 1, 3, 70), (128, 1, 4, 3, 17)])
@pytest.mark.parametrize("kind,tol", [(K.BF16X3, 3e-5), (K.F16X3, 1e-5), (K.BF16, 6e-3)])
def test_syrk_conv_patches_vs_unfold(geom, kind, tol):
    """im2col-free A factor (shifted 4-D TMA boxes feeding the MN-major SYRK) == fp64 unfold + P^T P, in the parameter
    order (ci, kh, kw), accumulating into the existing factor; ragged image counts included."""
    Ci, H, W, k, Q = geom
    torch.manual_seed(3)
    mod = torch.nn.Conv2d(Ci, 8, k, 1, k // 2)
    x = torch.randn(Q, Ci, H, W)
    P = F.unfold(x.double(), k, padding=k // 2).transpose(1, 2).reshape(Q * H * W, -1)
    ref = P.t() @ P
    assert K.conv_patches_ok(Ci, H, W, k, k)
    X = conv_engine.nhwc_rows(x.to(DEV).contiguous(memory_format=torch.channels_last), kind)
    base = torch.randn(ref.shape[0], ref.shape[0])
    out = base.to(DEV)
    K.syrk_conv_patches(X, Q, H, W, mod, out, alpha=0.5)
    assert rel_fro(out.cpu().double() - base.double(), 0.5 * ref) < tol
    assert rel_fro(out - base.to(DEV), (out - base.to(DEV)).t()) < 1e-5


def test_implicit_patch_factor_model_parity():
    """A 64-channel convolution stack: KFAC-GGN factors through the im2col-free path == oracle (<= 1e-4)."""
    from laplace_b200 import B200GGN
    from oracle import curvature_oracle as co

    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 64, 3, 1, 1), torch.nn.ReLU(), torch.nn.Conv2d(64, 64, 3, 1, 1, bias=False),
                                torch.nn.ReLU(), torch.nn.Conv2d(64, 128, 1), torch.nn.ReLU(),
                                torch.nn.Conv2d(128, 64, 3, 1, 1), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(),
                                torch.nn.Linear(64, 5)).eval()
    X, y = torch.randn(12, 3, 8, 8), torch.randint(5, (12,))
    calls = []
    orig = K.syrk_conv_patches
    K.syrk_conv_patches = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        _, kron = B200GGN(model.to(DEV), "classification").kron(X.to(DEV), y.to(DEV), N=36)
    finally:
        K.syrk_conv_patches = orig
    assert len(calls) == 3
    _, kf = co.kfac_factors(model.cpu().double(), "classification", X.double(), y, N=36)
    worst = max(rel_fro(H.cpu(), Ho) for F_, Fo in zip(kron.kfacs, kf) for H, Ho in zip(F_, Fo))
    assert worst < 1e-4, worst


@pytest.mark.parametrize("geom", [(64, 32, 8, 8, 3, 6, 2), (160, 40, 8, 8, 3, 5, 1), (16, 130, 16, 16, 3, 3, 3), (24, 8, 32, 32, 1, 2, 2),
                                  (64, 64, 16, 8, 3, 4, 10)])
@pytest.mark.parametrize("kind,tol", [(K.BF16X3, 5e-5), (K.BF16, 2e-2)])
def test_diag_conv_sq_vs_per_sample_gradients(geom, kind, tol):
    """Tensor-core diagonal of a convolution weight: sum over (column, sample) of the squared per-sample weight
    gradient == fp64 einsum on unfolded patches, in the parameter order (co, ci, kh, kw)."""
    Ci, Co, H, W, k, Nimg, ncols = geom
    torch.manual_seed(4)
    mod = torch.nn.Conv2d(Ci, Co, k, 1, k // 2)
    x = torch.randn(Nimg, Ci, H, W)
    g = torch.randn(ncols * Nimg, Co, H, W)
    P = F.unfold(x.double(), k, padding=k // 2).transpose(1, 2)                          # [Nimg, T, Ci*k*k]
    per = torch.einsum("cnto,ntp->cnop", g.double().reshape(ncols, Nimg, Co, H * W).transpose(2, 3), P)
    ref = (per * per).sum((0, 1))
    assert K.diag_conv_ok(Ci, H, W, k, k)
    cl = torch.channels_last
    X = conv_engine.nhwc_rows(x.to(DEV).contiguous(memory_format=cl), kind)
    G = conv_engine.nhwc_rows(g.to(DEV).contiguous(memory_format=cl), kind)
    base = torch.rand(Co, Ci * k * k)
    out = base.to(DEV)
    K.diag_conv_sq(G, X, Nimg, H, W, mod, out, alpha=0.25)
    assert rel_fro(out.cpu().double() - base.double(), 0.25 * ref) < tol


def test_diag_tensor_core_model_parity():
    """diag GGN / diag EF of a convolution stack through the tensor-core diagonal == oracle (rel-fro <= 1e-4)."""
    from laplace_b200 import B200EF
    from oracle import curvature_oracle as co

    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 32, 3, 1, 1), torch.nn.ReLU(), torch.nn.Conv2d(32, 48, 3, 1, 1, bias=False),
                                torch.nn.ReLU(), torch.nn.Conv2d(48, 16, 1), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(),
                                torch.nn.Linear(16, 5)).eval()
    X, y = torch.randn(7, 3, 8, 8), torch.randint(5, (7,))
    calls = []
    orig = K.diag_conv_sq
    K.diag_conv_sq = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        _, d = B200GGN(model.to(DEV), "classification").diag(X.to(DEV), y.to(DEV))
        _, de = B200EF(model, "classification").diag(X.to(DEV), y.to(DEV))
    finally:
        K.diag_conv_sq = orig
    assert len(calls) == 4
    m64 = model.cpu().double()
    Js, f = co.jacobians(m64, X.double())
    _, dref = co.ggn_diag(Js, f, y, "classification")
    _, deref = co.ef_diag(Js, f, y, "classification")
    assert rel_fro(d.cpu(), dref) < 1e-4 and rel_fro(de.cpu(), deref) < 1e-4
