"""Kernel-level parity on the B200: every C-ABI entry point (called through laplace_b200.kernels) against the
layout-exact torch emulation in tests/cpu_kernels.py evaluated on the same seeded inputs."""
import pytest
import torch

from laplace_b200 import kernels as K
from tests import cpu_kernels as ck
from tests.fixtures import rel_fro

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _dense(p):
    return p.hi[:, :p.K].float().cpu() + (p.lo[:, :p.K].float().cpu() if p.lo is not None else 0)


@pytest.mark.parametrize("kind,tol", [(K.F32, 0.0), (K.BF16, 4e-3), (K.BF16X3, 2e-5), (K.F16X3, 5e-7)])
@pytest.mark.parametrize("shape", [(1, 1), (37, 5), (64, 64), (1000, 130), (33, 513)])
def test_pack_rows(kind, tol, shape):
    torch.manual_seed(0)
    src = torch.randn(*shape)
    ref = ck.pack_rows(src, K.F32)
    out = K.pack_rows(src.to(DEV), kind)
    assert out.ldk % 8 == 0
    assert (_dense(out) - ref.hi[:, :ref.K]).abs().max() <= tol * src.abs().max() + 1e-30
    w = torch.rand(3 * shape[0])
    ref = ck.pack_rows(src, K.F32, square=True, row_scale=w, nrep=3, scale=0.5)
    out = K.pack_rows(src.to(DEV), kind, square=True, row_scale=w.to(DEV), nrep=3, scale=0.5)
    assert rel_fro(_dense(out), ref.hi[:, :ref.K]) <= max(tol, 1e-7)


@pytest.mark.parametrize("geom", [(3, 4, 2, 2, 0, 1, 5), (3, 8, 3, 1, 1, 1, 8), (16, 8, 3, 2, 1, 1, 9), (4, 4, 1, 2, 0, 1, 7),
                                  (2, 4, 3, 1, 2, 2, 9), (3, 64, 7, 2, 3, 1, 32)])
def test_pack_conv_and_nchw(geom):
    cin, cout, k, s, p, d, hw = geom
    torch.manual_seed(1)
    mod = torch.nn.Conv2d(cin, cout, k, s, p, dilation=d)
    x = torch.randn(5, cin, hw, hw)
    for reduce in (False, True):
        ref, T = ck.pack_conv(x, mod, K.F32, reduce_mean=reduce)
        out, T2 = K.pack_conv(x.to(DEV), mod, K.F32, reduce_mean=reduce)
        assert T == T2 and rel_fro(_dense(out), ref.hi[:, :ref.K]) < 1e-6
    o = mod(x)
    g = torch.randn(7, cout, o.shape[2] * o.shape[3])
    for reduce in (False, True):
        ref = ck.pack_nchw(g, K.F32, reduce_sum=reduce)
        out = K.pack_nchw(g.to(DEV), K.BF16X3, reduce_sum=reduce)
        assert rel_fro(_dense(out), ref.hi[:, :ref.K]) < 2e-5


GEMM_SHAPES = [(1, 1, 1), (5, 7, 3), (64, 64, 16), (70, 130, 1000), (128, 128, 64), (200, 96, 517), (513, 257, 2048),
               (300, 300, 40000)]


@pytest.mark.parametrize("M,N,Kc", GEMM_SHAPES)
def test_gemm_nt_f32(M, N, Kc):
    torch.manual_seed(2)
    A, B = torch.randn(M, Kc), torch.randn(N, Kc)
    ref = (A.double() @ B.double().t())
    pa, pb = K.pack_rows(A.t().contiguous().to(DEV), K.F32), K.pack_rows(B.t().contiguous().to(DEV), K.F32)
    base = torch.randn(M, N)
    out = base.to(DEV)
    K.gemm_nt(pa, pb, out, alpha=0.5, accumulate=True)
    assert rel_fro(out.cpu(), base.double() + 0.5 * ref) < 2e-6
    K.gemm_nt(pa, pb, out, alpha=1.0, accumulate=False)
    assert rel_fro(out.cpu(), ref) < 2e-6
    if M == N:
        sym = torch.zeros(M, M, device=DEV)
        K.gemm_nt(pa, pa, sym, alpha=1.0, accumulate=True, symmetric=True)
        assert rel_fro(sym.cpu(), A.double() @ A.double().t()) < 2e-6
        assert rel_fro(sym, sym.t()) < 1e-6  # symmetric up to the fp32 order of the split-K reductions


@pytest.mark.parametrize("M,N,Kc", [(128, 128, 64), (128, 128, 4096), (100, 60, 50), (513, 257, 2048), (300, 300, 40000),
                                    (1000, 1000, 512), (64, 2000, 130)])
@pytest.mark.parametrize("kind,tol", [(K.BF16, 6e-3), (K.BF16X3, 3e-5), (K.F16X3, 1e-5)])
def test_gemm_nt_tensor_core(M, N, Kc, kind, tol):
    torch.manual_seed(3)
    A, B = torch.randn(M, Kc), torch.randn(N, Kc)
    ref = A.double() @ B.double().t()
    pa, pb = K.pack_rows(A.t().contiguous().to(DEV), kind), K.pack_rows(B.t().contiguous().to(DEV), kind)
    out = torch.zeros(M, N, device=DEV)
    K.gemm_nt(pa, pb, out, alpha=2.0, accumulate=True)
    K.gemm_nt(pa, pb, out, alpha=-1.0, accumulate=True)
    assert rel_fro(out.cpu(), ref) < tol
    out.fill_(7.0)
    K.gemm_nt(pa, pb, out, alpha=1.0, accumulate=False)
    assert rel_fro(out.cpu(), ref) < tol
    if M == N:
        sym = torch.zeros(M, M, device=DEV)
        K.gemm_nt(pa, pa, sym, alpha=1.0, accumulate=True, symmetric=True)
        assert rel_fro(sym.cpu(), A.double() @ A.double().t()) < tol
        assert rel_fro(sym, sym.t()) < 1e-5
        assert (sym.diagonal() >= 0).all()


def test_tensor_core_syrk_large_properties():
    """BASELINE-size SYRK (d=4608, K=8192) through size-independent properties: symmetry, PSD diagonal,
    trace == sum of squares, linearity in K (two halves add up to the whole)."""
    torch.manual_seed(4)
    d, Kc = 4608, 8192
    X = torch.randn(Kc, d, device=DEV)
    p = K.pack_rows(X, K.BF16X3)
    H = torch.zeros(d, d, device=DEV)
    K.gemm_nt(p, p, H, 1.0, True, symmetric=True)
    assert rel_fro(H, H.t()) < 1e-6
    assert abs(float(H.diagonal().sum() / (X.double() ** 2).sum()) - 1) < 2e-5  # accumulator truncation bias, K<=2048 per TMEM tile
    H2 = torch.zeros(d, d, device=DEV)
    for part in (X[:3000], X[3000:]):
        pp = K.pack_rows(part.contiguous(), K.BF16X3)
        K.gemm_nt(pp, pp, H2, 1.0, True, symmetric=True)
    assert rel_fro(H2, H) < 1e-5
    ref = X[:, :256].double().t() @ X.double()[:, 1000:1300]
    assert rel_fro(H[:256, 1000:1300].double(), ref) < 3e-5


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("dims", [(4, 12, 4, 5, 2), (70, 100, 17, 3, 3), (64, 64, 64, 2, 1), (130, 27, 9, 6, 10)])
def test_shared_weight_contract(mode, dims):
    d_out, d_in, T, Nn, ncols = dims
    torch.manual_seed(5)
    G = ck._alloc(d_out, ncols * Nn * T, K.F32, None)
    G.hi[:, :G.K] = torch.randn(d_out, G.K)
    A = ck._alloc(d_in, Nn * T, K.F32, None)
    A.hi[:, :A.K] = torch.randn(d_in, A.K)
    Gd = K.Packed(G.hi.to(DEV), None, K.F32, G.rows, G.K)
    Ad = K.Packed(A.hi.to(DEV), None, K.F32, A.rows, A.K)
    if mode == 0:
        ref = torch.zeros(d_out, d_in)
        ck.shared_weight_contract(0, G, A, d_out, d_in, T, Nn, ncols, ref, scale=0.3)
        out = torch.zeros(d_out, d_in, device=DEV)
        K.shared_weight_contract(0, Gd, Ad, d_out, d_in, T, Nn, ncols, out, scale=0.3, out_ld=d_in)
        assert rel_fro(out.cpu(), ref) < 1e-5
    else:
        P = d_out * d_in + 3
        ref = torch.zeros(Nn, ncols, P)
        ck.shared_weight_contract(1, G, A, d_out, d_in, T, Nn, ncols, ref[..., 2:], js_stride_n=ncols * P, js_stride_c=P)
        out = torch.zeros(Nn, ncols, P, device=DEV)
        K.shared_weight_contract(1, Gd, Ad, d_out, d_in, T, Nn, ncols, out[..., 2:], js_stride_n=ncols * P, js_stride_c=P)
        assert rel_fro(out.cpu(), ref) < 1e-5


@pytest.mark.parametrize("dims", [(4, 12, 4, 5, 2), (70, 100, 17, 3, 3), (64, 64, 64, 2, 10), (130, 27, 1, 6, 10), (200, 330, 5, 4, 12),
                                  (512, 640, 1, 3, 10)])
@pytest.mark.parametrize("damping", [False, True])
def test_kron_conv_quadform(dims, damping):
    """Fused eigenbasis quadratic form of a weight-sharing layer vs an fp64 einsum on the same rotated rows (ragged tile
    edges, T below / above the 16-deep staging, C = 2..12, plain and damped spectrum)."""
    d_out, d_in, T, Nn, C = dims
    torch.manual_seed(8)
    Gt, At = torch.randn(d_out, C * Nn * T), torch.randn(d_in, Nn * T)
    l1, l2 = torch.rand(d_out) * 2, torch.rand(d_in) * 3
    l1[:2] = 0
    delta = 0.37
    ref = torch.zeros(Nn, C, C, dtype=torch.float64)
    ck.kron_conv_quadform(Gt, At, T, Nn, C, l1, l2, delta, damping, ref)
    out = torch.full((Nn, C, C), 0.5, device=DEV)
    K.kron_conv_quadform(Gt.to(DEV), At.to(DEV), T, Nn, C, l1.to(DEV), l2.to(DEV), delta, damping, out)
    assert rel_fro(out.cpu().double() - 0.5, ref) < 2e-6
    assert rel_fro(out, out.transpose(1, 2)) < 1e-7


def test_jacobian_writers_and_pair_dot():
    torch.manual_seed(6)
    C, Nn, d_out, d_in = 3, 5, 7, 11
    g, a = torch.randn(C, Nn, d_out), torch.randn(Nn, d_in)
    P = d_out * d_in + d_out + 4
    ref = torch.zeros(Nn, C, P)
    ck.jac_linear_write(g, a, ref, C * P, P, 4, 4 + d_out * d_in)
    out = torch.zeros(Nn, C, P, device=DEV)
    K.jac_linear_write(g.to(DEV), a.to(DEV), out, C * P, P, 4, 4 + d_out * d_in)
    assert torch.allclose(out.cpu(), ref, atol=1e-6)
    phi = torch.randn(6, 9)
    for hb in (True, False):
        assert torch.equal(K.ll_jacobian_write(phi.to(DEV), 4, hb).cpu(), ck.ll_jacobian_write(phi, 4, hb))
    X, Z, m = torch.randn(Nn, C, 40), torch.randn(Nn, 2, 40), torch.rand(Nn, 40)
    for mm in (None, m, m[0].contiguous()):
        ref = ck.batched_pair_dot(X, Z, mm, torch.ones(Nn, C, 2), accumulate=True)
        out = K.batched_pair_dot(X.to(DEV), Z.to(DEV), None if mm is None else mm.to(DEV), torch.ones(Nn, C, 2, device=DEV),
                                 accumulate=True)
        assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-5)
    Xp = torch.randn(C, Nn, 40)  # permuted view (c-major storage)
    ref = ck.batched_pair_dot(Xp.permute(1, 0, 2), Xp.permute(1, 0, 2), None, torch.zeros(Nn, C, C))
    out = K.batched_pair_dot(Xp.to(DEV).permute(1, 0, 2), Xp.to(DEV).permute(1, 0, 2), None, torch.zeros(Nn, C, C, device=DEV))
    assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("C,D,hb", [(2, 3, True), (3, 5, False), (10, 33, True)])
def test_last_layer_block_kernels(C, D, hb):
    torch.manual_seed(7)
    Dt, npairs = D + int(hb), C * (C + 1) // 2
    G = torch.randn(Dt, npairs * Dt)
    P = C * D + (C if hb else 0)
    ref = torch.ones(P, P)
    ck.ll_ggn_expand(G, C, D, hb, ref, accumulate=True)
    out = torch.ones(P, P, device=DEV)
    K.ll_ggn_expand(G.to(DEV), C, D, hb, out, accumulate=True)
    assert torch.allclose(out.cpu(), ref)
    S = torch.randn(P, P)
    assert torch.equal(K.ll_sigma_gather(S.to(DEV), C, D, hb).cpu(), ck.ll_sigma_gather(S, C, D, hb))


@pytest.mark.parametrize("n", [1, 2, 3, 10, 20, 63, 64, 127, 128])
def test_eigh_jacobi(n):
    torch.manual_seed(8)
    Z = torch.randn(4, n, 3 * n)
    A = Z @ Z.transpose(1, 2) / (3 * n)
    A[3] = torch.diag(torch.rand(n))            # already diagonal
    if n > 2:
        A[2, :, 0] = 0; A[2, 0, :] = 0          # a zero row/col (rank deficient)
    ev, Q = K.eigh_jacobi(A.to(DEV))
    ev, Q = ev.cpu().double(), Q.cpu().double()
    ref = torch.linalg.eigvalsh(A.double()).clamp(min=0)
    assert torch.allclose(ev, ref, rtol=1e-4, atol=1e-5 * float(ref.max()))
    assert (ev[:, 1:] >= ev[:, :-1]).all()
    rec = Q @ torch.diag_embed(ev) @ Q.transpose(1, 2)
    assert rel_fro(rec, A) < 1e-5
    assert rel_fro(Q.transpose(1, 2) @ Q, torch.eye(n).expand(4, n, n)) < 1e-5
    # only the upper triangle is read (UPLO='U')
    Au = A.clone(); Au[:, 1:, 0] = 123.0
    ev2, _ = K.eigh_jacobi(Au.to(DEV))
    assert torch.allclose(ev2.cpu().double(), ev, atol=1e-6)


@pytest.mark.parametrize("batched,threads", [(True, 4), (False, 4), (False, 1)])
def test_decompose_routes_agree_with_fp64(batched, threads):
    """``B200Kron.decompose`` over a mix of factor sizes -- hand-written Jacobi (<= 128), the batched mid-size route
    (bordered to 256 / 513 rows, dead coordinates compacted), one library call per factor beyond, host threads -- against an
    fp64 eigendecomposition: eigenvalues, reconstruction, orthogonality, ``UPLO='U'`` semantics of ``symeig``
    (utils/utils.py:193-228)."""
    from laplace_b200 import B200Kron, matrix

    torch.manual_seed(12)

    def psd(n, dead=0):
        Z = torch.randn(n, 2 * n)
        A = Z @ Z.T / (2 * n)
        if dead:
            A[:dead] = 0
            A[:, :dead] = 0
        return A

    facs = [psd(64), psd(130), psd(200), psd(256), psd(300), psd(513), psd(600, dead=400), psd(700), psd(1200, dead=900), psd(40), psd(147)]
    kron = B200Kron([[f.to(DEV)] for f in facs])
    keep = (matrix.BATCHED_MID_SIZES, matrix.N_EIGH_THREADS)
    matrix.BATCHED_MID_SIZES, matrix.N_EIGH_THREADS = batched, threads
    try:
        kd = kron.decompose()
    finally:
        matrix.BATCHED_MID_SIZES, matrix.N_EIGH_THREADS = keep
    for f, Q, L in zip(facs, kd.eigenvectors, kd.eigenvalues):
        n = f.shape[0]
        Q, L = Q[0].cpu().double(), L[0].cpu().double()
        assert Q.shape == (n, n) and L.shape == (n,) and (L[1:] >= L[:-1]).all() and float(L.min()) >= 0
        ref = torch.linalg.eigvalsh(f.double()).clamp(min=0)
        assert torch.allclose(L, ref, rtol=1e-4, atol=2e-6 * float(ref.max())), n
        assert rel_fro((Q * L) @ Q.T, f) < 1e-5, n
        assert float((Q.T @ Q - torch.eye(n, dtype=torch.float64)).abs().max()) < 2e-5, n


@pytest.mark.parametrize("Kr,M,N", [(64, 128, 128), (512, 64, 64), (1000, 200, 96), (4096, 128, 128), (5000, 513, 257), (40000, 300, 300), (33, 10, 10)])
@pytest.mark.parametrize("kind,tol", [(K.BF16, 6e-3), (K.BF16X3, 3e-5), (K.F16X3, 1e-5)])
def test_gemm_tn_rows_operands(Kr, M, N, kind, tol):
    """MN-major tcgen05 path: D = A^T B on row-major [samples, features] operands (no transposing pack)."""
    torch.manual_seed(9)
    A, B = torch.randn(Kr, M), torch.randn(Kr, N)
    ref = A.double().t() @ B.double()
    pa, pb = K.pack_cast(A.to(DEV), kind), K.pack_cast(B.to(DEV), kind)
    out = torch.zeros(M, N, device=DEV)
    K.gemm_tn(pa, pb, out, alpha=2.0, accumulate=True)
    K.gemm_tn(pa, pb, out, alpha=-1.0, accumulate=True)
    assert rel_fro(out.cpu(), ref) < tol
    out.fill_(3.0)
    K.gemm_tn(pa, pb, out, alpha=1.0, accumulate=False)
    assert rel_fro(out.cpu(), ref) < tol
    if M == N:
        sym = torch.zeros(M, M, device=DEV)
        K.gemm_tn(pa, pa, sym, alpha=1.0, accumulate=True, symmetric=True)
        assert rel_fro(sym.cpu(), A.double().t() @ A.double()) < tol


@pytest.mark.parametrize("shape", [(6, 8, 4, 4), (5, 7, 3, 5), (33, 64, 8, 8), (3, 5, 1, 1)])
def test_elementwise_reverse_pass_kernels(shape):
    torch.manual_seed(10)
    Q, C, H, W = shape
    g = torch.randn(Q, C, H, W, device=DEV)
    s = torch.rand(C, device=DEV) + 0.5
    ref = g * s.view(1, -1, 1, 1)
    assert torch.allclose(K.scale_channels(g, s), ref, rtol=1e-6)
    gcl = g.contiguous(memory_format=torch.channels_last)
    out = K.scale_channels(gcl, s)
    assert out.stride() == gcl.stride() and torch.allclose(out, ref, rtol=1e-6)
    B = 3 if Q % 3 == 0 else 1
    y = torch.randn(B, C, H, W, device=DEV).relu()
    reps = Q // B
    ref = (g.view(reps, B, C, H, W) * (y > 0)).view(Q, C, H, W)
    assert torch.equal(K.relu_bwd(g, y, reps), ref)


@pytest.mark.parametrize("geom", [(3, 2, 1, 16), (2, 2, 0, 8), (3, 1, 1, 7), (3, 2, 1, 9)])
def test_maxpool_backward_kernel(geom):
    k, s, p, hw = geom
    torch.manual_seed(11)
    x = torch.randn(4, 6, hw, hw, device=DEV, requires_grad=True)
    out, idx = torch.nn.functional.max_pool2d(x, k, s, p, return_indices=True)
    g = torch.randn(3 * 4, 6, *out.shape[2:], device=DEV)          # 3 folded columns
    ref = torch.stack([torch.autograd.grad(out, x, g[i * 4:(i + 1) * 4], retain_graph=True)[0] for i in range(3)]).reshape(12, 6, hw, hw)
    got = K.maxpool2d_bwd(g, idx, x.shape, k, s, p)
    assert torch.allclose(got, ref, atol=1e-6)
    # channels-last operands take the NHWC kernel and stay channels-last
    cl = torch.channels_last
    xc = x.detach().contiguous(memory_format=cl).requires_grad_(True)
    outc, idxc = torch.nn.functional.max_pool2d(xc, k, s, p, return_indices=True)
    assert torch.equal(idxc, idx)
    gotc = K.maxpool2d_bwd(g.contiguous(memory_format=cl), idxc.contiguous(memory_format=cl), x.shape, k, s, p)
    assert gotc.is_contiguous(memory_format=cl) and torch.allclose(gotc, ref, atol=1e-6)


@pytest.mark.parametrize("geom", [(3, 2, 1, 16, 64), (2, 2, 0, 8, 8), (3, 1, 1, 7, 16), (3, 2, 1, 9, 24)])
@pytest.mark.parametrize("with_scale,with_mask", [(True, True), (False, True), (True, False)])
def test_maxpool_backward_fused_with_operand_split(geom, with_scale, with_mask):
    """``maxpool2d_bwd_pack``: un-pool + ReLU mask + BN scale + bf16 hi/lo split in one pass == the three separate kernels."""
    k, s, p, hw, C = geom
    torch.manual_seed(13)
    cl = torch.channels_last
    x = torch.randn(4, C, hw, hw, device=DEV).contiguous(memory_format=cl)
    out, idx = torch.nn.functional.max_pool2d(x, k, s, p, return_indices=True)
    g = torch.randn(3 * 4, C, *out.shape[2:], device=DEV).contiguous(memory_format=cl)
    scale = (torch.rand(C, device=DEV) + 0.5) if with_scale else None
    y = torch.relu(x) if with_mask else None
    P = K.maxpool2d_bwd_pack(g, idx.contiguous(memory_format=cl), x.shape, k, s, p, scale, y)
    un = K.maxpool2d_bwd(g, idx.contiguous(memory_format=cl), x.shape, k, s, p)
    if scale is not None:
        un = un * scale.view(1, -1, 1, 1)
    if y is not None:
        un = (un.reshape(3, 4, C, hw, hw) * (y > 0)).reshape(12, C, hw, hw)
    ref = un.permute(0, 2, 3, 1).reshape(-1, C)
    got = P.hi[:, :C].float() + P.lo[:, :C].float()
    assert P.kind == K.BF16X3 and P.rows == 12 * hw * hw
    assert rel_fro(got, ref) < 2e-5 and torch.equal(P.hi[:, :C], ref.bfloat16())


@pytest.mark.parametrize("M,N,Kc", [(256, 256, 64), (256, 256, 4096), (513, 257, 2048), (300, 300, 40000), (1000, 1000, 512),
                                    (1152, 1152, 3000), (640, 2000, 130), (64, 64, 300000)])
@pytest.mark.parametrize("kind,tol", [(K.BF16, 6e-3), (K.BF16X3, 3e-5), (K.F16X3, 1e-5)])
@pytest.mark.parametrize("rows", [False, True])
@pytest.mark.parametrize("mode", [1, 2])
def test_gemm_cta_pair_tiles(M, N, Kc, kind, tol, rows, mode):
    """Alternative schedules of the tensor-core contraction -- 256 x 256 CTA-pair tiles (tcgen05 cta_group::2) and
    persistent CTAs with double-buffered TMEM -- on K-major and row (MN-major) operands: same results as fp64 and as
    the one-tile-per-CTA kernel, including ragged edges, split-K accumulation, overwrite mode and the mirrored SYRK."""
    torch.manual_seed(11)
    A, B = torch.randn(M, Kc), torch.randn(N, Kc)
    ref = A.double() @ B.double().t()
    if rows:
        pa, pb = K.pack_cast(A.t().contiguous().to(DEV), kind), K.pack_cast(B.t().contiguous().to(DEV), kind)
        gemm = K.gemm_tn
    else:
        pa, pb = K.pack_rows(A.t().contiguous().to(DEV), kind), K.pack_rows(B.t().contiguous().to(DEV), kind)
        gemm = K.gemm_nt
    try:
        K.set_gemm_tile_mode(mode)
        out = torch.zeros(M, N, device=DEV)
        gemm(pa, pb, out, alpha=2.0, accumulate=True)
        gemm(pa, pb, out, alpha=-1.0, accumulate=True)
        assert rel_fro(out.cpu(), ref) < tol
        out.fill_(7.0)
        gemm(pa, pb, out, alpha=1.0, accumulate=False)
        assert rel_fro(out.cpu(), ref) < tol
        K.set_gemm_tile_mode(0)
        single = torch.zeros(M, N, device=DEV)
        gemm(pa, pb, single, alpha=1.0, accumulate=False)
        assert rel_fro(out, single) < (1e-3 if kind == K.BF16 else 2e-5)
        if M == N:
            K.set_gemm_tile_mode(mode)
            sym = torch.zeros(M, M, device=DEV)
            gemm(pa, pa, sym, alpha=1.0, accumulate=True, symmetric=True)
            assert rel_fro(sym.cpu(), A.double() @ A.double().t()) < tol
            assert rel_fro(sym, sym.t()) < 1e-5
            assert (sym.diagonal() >= 0).all()
    finally:
        K.set_gemm_tile_mode(-1)


@pytest.mark.parametrize("shape", [(640, 64), (96, 24), (130, 7), (4096, 512)])
@pytest.mark.parametrize("kind,tol", [(K.BF16X3, 2e-5), (K.F16X3, 5e-7), (K.BF16, 4e-3), (K.F32, 0.0)])
def test_pack_cast_fused(shape, kind, tol):
    """ReLU mask (shared by the folded columns) x BatchNorm scale x operand split in one pass == the three separate maps."""
    rows, cols = shape
    torch.manual_seed(12)
    reps = 5 if rows % 5 == 0 else 2
    g = torch.randn(rows, cols, device=DEV)
    y = torch.randn(rows // reps, cols, device=DEV).clamp_min(0)
    scale = torch.rand(cols, device=DEV) + 0.5
    ref = (g.view(reps, rows // reps, cols) * (y > 0)).view(rows, cols) * scale
    for sc, yy, want in ((scale, y, ref), (None, y, (g.view(reps, -1, cols) * (y > 0)).view(rows, cols)), (scale, None, g * scale),
                         (None, None, g)):
        p = K.pack_cast_fused(g, kind, sc, yy)
        got = p.hi[:, :cols].float() + (p.lo[:, :cols].float() if p.lo is not None else 0)
        err = float((got - want).norm() / want.norm())
        assert err <= tol, (err, sc is None, yy is None)


@pytest.mark.parametrize("rows_y,cols,reps", [(96, 64, 10), (1000, 24, 3), (33, 512, 7), (1 << 14, 256, 3)])
def test_mask_major_kernels_bit_identical(rows_y, cols, reps):
    """``pack_cast_fused`` / ``relu_bwd`` with the mask read once per element (mask-major: what runs for masks >= 4 Mi elements,
    the last shape here by default) against the row-major kernels: the same bits, and both equal to the separate maps."""
    torch.manual_seed(rows_y + reps)
    g = torch.randn(reps * rows_y, cols, device=DEV)
    y = torch.randn(rows_y, cols, device=DEV).clamp_min(0)
    scale = torch.rand(cols, device=DEV) + 0.5
    want = (g.view(reps, rows_y, cols) * scale * (y > 0)).view(-1, cols)
    try:
        if rows_y * cols < (4 << 20):
            K.set_mask_major_min(0)
        outs = {}
        for mode in ("mask-major", "row-major"):
            if mode == "row-major":
                K.set_mask_major_min(-1)
            for kind in (K.BF16X3, K.F16X3):
                for sc in (scale, None):
                    p = K.pack_cast_fused(g, kind, sc, y)
                    outs[mode, kind, sc is None] = (p.hi[:, :cols].clone(), p.lo[:, :cols].clone())
            outs[mode, "relu"] = K.relu_bwd(g, y, reps)
        for key, val in outs.items():
            if key[0] != "mask-major":
                continue
            other = outs[("row-major",) + key[1:]]
            if key[1] == "relu":
                assert torch.equal(val, other) and torch.equal(val, (g.view(reps, rows_y, cols) * (y > 0)).view(-1, cols))
            else:
                assert torch.equal(val[0], other[0]) and torch.equal(val[1], other[1]), key
        hi, lo = outs["mask-major", K.F16X3, False]
        assert float((hi.float() + lo.float() - want).norm() / want.norm()) < 5e-7
    finally:
        K.set_mask_major_min(4 << 20)


def test_layer_level_kfac_entry_points():
    """``lpb_kfac_accum_rows`` / ``lpb_kfac_accum_conv_input`` (SURVEY 8(b) plan-level surface): fp32 tensors in, one call per
    Kronecker factor, against fp64."""
    import ctypes as C

    from laplace_b200 import _lib

    lib = _lib.load()
    torch.manual_seed(3)
    st = torch.cuda.current_stream().cuda_stream
    X = torch.randn(5000, 200, device=DEV)
    ws = torch.empty(lib.lpb_workspace_bytes(5000, 200), device=DEV, dtype=torch.uint8)
    for fp16 in (1, 0):
        out = torch.zeros(200, 200, device=DEV)
        _lib.call("lpb_kfac_accum_rows", X.data_ptr(), 5000, 200, 200, 0.5, fp16, ws.data_ptr(), ws.numel(), out.data_ptr(), 200, st)
        _lib.call("lpb_kfac_accum_rows", X.data_ptr(), 5000, 200, 200, 0.5, fp16, ws.data_ptr(), ws.numel(), out.data_ptr(), 200, st)
        assert rel_fro(out.cpu(), (X.double().T @ X.double()).cpu()) < (1e-5 if fp16 else 3e-5)
    x = torch.randn(40, 16, 9, 9, device=DEV)
    conv = torch.nn.Conv2d(16, 8, 3, 2, 1)
    P = torch.nn.functional.unfold(x.double(), 3, padding=1, stride=2).permute(0, 2, 1).reshape(-1, 144)
    ws2 = torch.empty(lib.lpb_workspace_bytes(P.shape[0], 144), device=DEV, dtype=torch.uint8)
    out = torch.zeros(144, 144, device=DEV)
    _lib.call("lpb_kfac_accum_conv_input", x.data_ptr(), 40, 16, 9, 9, 3, 3, 2, 2, 1, 1, 1, 1, 0.25, ws2.data_ptr(), ws2.numel(),
              out.data_ptr(), 144, st)
    assert rel_fro(out.cpu(), 0.25 * (P.T @ P).cpu()) < 1e-5
    with pytest.raises(RuntimeError, match="workspace too small"):
        _lib.call("lpb_kfac_accum_rows", X.data_ptr(), 5000, 200, 200, 0.5, 1, ws.data_ptr(), 16, out.data_ptr(), 200, st)
