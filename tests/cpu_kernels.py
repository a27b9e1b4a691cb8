"""TEST DOUBLE (tests only): torch-CPU emulation of the native kernel entry points, with the same
operand layouts as ``include/laplace_b200.h``.

Purpose: exercise the *host logic* of ``laplace_b200`` (layer plan, hook capture, column
construction, scaling conventions, Kron assembly, dispatch into the reference's ``Laplace`` classes)
in the GPU-less build container.  It is installed by ``tests/conftest.py``'s ``cpu_kernels`` fixture via
monkeypatching and is never importable from the product package; the product path raises when the
CUDA library / a CUDA device is missing.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from laplace_b200 import kernels as K


def _alloc(rows, Kc, kind, device):
    ldk = max(K.round_up(Kc, 8), 8)
    return K.Packed(torch.zeros(rows, ldk, dtype=torch.float32), None, kind, rows, Kc)


def pack_rows(src, kind, out=None, k0=0, scale=1.0, square=False, row_scale=None, nrep=1, total_K=None):
    Kc, d = src.shape
    if out is None:
        out = _alloc(nrep * d, total_K if total_K is not None else Kc, kind, src.device)
    v = src.float() * scale
    if square:
        v = v * v
    for z in range(nrep):
        vz = v if row_scale is None else v * row_scale.reshape(nrep, Kc)[z].unsqueeze(1)
        out.hi[z * d:(z + 1) * d, k0:k0 + Kc] = vz.t()
    return out


def pack_conv(x, mod, kind, reduce_mean=False, square=False):
    cols = F.unfold(x.float(), mod.kernel_size, dilation=mod.dilation, padding=mod.padding, stride=mod.stride)
    N, d, T = cols.shape
    if square:
        cols = cols * cols
    if reduce_mean:
        rows = cols.mean(2)  # [N, d]
        out = _alloc(d, N, kind, x.device)
        out.hi[:, :N] = rows.t()
    else:
        out = _alloc(d, N * T, kind, x.device)
        out.hi[:, :N * T] = cols.permute(1, 0, 2).reshape(d, N * T)
    return out, T


def pack_nchw(g, kind, reduce_sum=False, square=False):
    Nn, Cc, HW = g.shape
    v = g.float()
    if square:
        v = v * v
    if reduce_sum:
        out = _alloc(Cc, Nn, kind, g.device)
        out.hi[:, :Nn] = v.sum(2).t()
    else:
        out = _alloc(Cc, Nn * HW, kind, g.device)
        out.hi[:, :Nn * HW] = v.permute(1, 0, 2).reshape(Cc, Nn * HW)
    return out


def gemm_nt(A, B, out, alpha=1.0, accumulate=True, symmetric=False):
    res = alpha * (A.hi[:, :A.K].float() @ B.hi[:, :B.K].float().t())
    if accumulate:
        out += res
    else:
        out.copy_(res)
    return out


def shared_weight_contract(mode, G, A, d_out, d_in, T, Nn, ncols, out, scale=1.0, out_ld=0, js_stride_n=0, js_stride_c=0):
    Gm = G.hi[:, :ncols * Nn * T].reshape(d_out, ncols, Nn, T)
    Am = A.hi[:, :Nn * T].reshape(d_in, Nn, T)
    Pq = torch.einsum("icnt,jnt->cnij", Gm, Am)
    if mode == 0:
        out += scale * (Pq ** 2).sum((0, 1))
    else:
        flat = out.reshape(-1) if out.is_contiguous() else None
        base = torch.as_strided(out, (Nn, ncols, d_out * d_in), (js_stride_n, js_stride_c, 1))
        base.copy_(Pq.permute(1, 0, 2, 3).reshape(Nn, ncols, d_out * d_in))


def jac_linear_write(g, a, Js_view, stride_n, stride_c, off_w, off_b):
    Cc, Nn, d_out = g.shape
    d_in = a.shape[1]
    if off_w >= 0:
        dst = torch.as_strided(Js_view, (Nn, Cc, d_out * d_in), (stride_n, stride_c, 1), Js_view.storage_offset() + off_w)
        dst.copy_(torch.einsum("cni,nj->ncij", g, a).reshape(Nn, Cc, -1))
    if off_b >= 0:
        dst = torch.as_strided(Js_view, (Nn, Cc, d_out), (stride_n, stride_c, 1), Js_view.storage_offset() + off_b)
        dst.copy_(g.permute(1, 0, 2))


def ll_jacobian_write(phi, C_out, has_bias):
    Nn, D = phi.shape
    P = C_out * D + (C_out if has_bias else 0)
    Js = torch.zeros(Nn, C_out, P, dtype=phi.dtype)
    for c in range(C_out):
        Js[:, c, c * D:(c + 1) * D] = phi
        if has_bias:
            Js[:, c, C_out * D + c] = 1.0
    return Js


def batched_pair_dot(X, Z, m, out, accumulate=False):
    w = 1.0 if m is None else (m.unsqueeze(1) if m.dim() == 2 else m)
    res = torch.einsum("nci,nki->nck", X * w, Z)
    if accumulate:
        out += res
    else:
        out.copy_(res)
    return out


def ll_ggn_expand(G, C_out, D, has_bias, H, accumulate):
    Dt = D + (1 if has_bias else 0)
    npairs = C_out * (C_out + 1) // 2
    G3 = G.reshape(Dt, npairs, Dt)

    def idx(c, d):
        return c * D + d if d < D else C_out * D + c

    P = C_out * D + (C_out if has_bias else 0)
    out = torch.zeros(P, P, dtype=G.dtype)
    pair = 0
    for c in range(C_out):
        for k in range(c, C_out):
            blk = G3[:, pair, :]
            rows = torch.tensor([idx(c, d) for d in range(Dt)])
            cols = torch.tensor([idx(k, e) for e in range(Dt)])
            out[rows.unsqueeze(1), cols.unsqueeze(0)] = blk
            if k != c:
                out[cols.unsqueeze(1), rows.unsqueeze(0)] = blk.t()
            pair += 1
    if accumulate:
        H += out
    else:
        H.copy_(out)


def ll_sigma_gather(Sigma, C_out, D, has_bias):
    Dt = D + (1 if has_bias else 0)

    def idx(c, d):
        return c * D + d if d < D else C_out * D + c

    Sg = torch.zeros(C_out * C_out, Dt, Dt, dtype=Sigma.dtype)
    for c in range(C_out):
        for k in range(C_out):
            rows = torch.tensor([idx(c, d) for d in range(Dt)])
            cols = torch.tensor([idx(k, e) for e in range(Dt)])
            Sg[c * C_out + k] = Sigma[rows.unsqueeze(1), cols.unsqueeze(0)].t()
    return Sg.reshape(C_out * C_out * Dt, Dt)


def eigh_jacobi(A, max_sweeps=30):
    L, Q = torch.linalg.eigh(A.double(), UPLO="U")
    return torch.nan_to_num(L.clamp(min=0)).float(), torch.nan_to_num(Q).float()


def maxpool2d_bwd_pack(g, idx, in_shape, k, s, p, scale, y):
    Q, C, OH, OW = g.shape
    H, W = in_shape[-2:]
    Nb = idx.shape[0]
    un = maxpool2d_bwd(g.contiguous(), idx, in_shape, k, s, p)
    if scale is not None:
        un = un * scale.view(1, -1, 1, 1)
    if y is not None:
        un = (un.reshape(Q // Nb, Nb, C, H, W) * (y > 0)).reshape(Q, C, H, W)
    return pack_cast(un.permute(0, 2, 3, 1).reshape(Q * H * W, C).contiguous(), K.BF16X3)


def kron_conv_quadform(Gt, At, T, Nn, C, l1, l2, delta, damping, out):
    d_out, d_in = Gt.shape[0], At.shape[0]
    G = Gt[:, :C * Nn * T].reshape(d_out, C, Nn, T).double()
    A = At[:, :Nn * T].reshape(d_in, Nn, T).double()
    Z = torch.einsum("icnt,jnt->ncij", G, A)
    if damping:
        sd = float(delta) ** 0.5
        w = 1.0 / torch.outer(l1.double() + sd, l2.double() + sd)
    else:
        w = 1.0 / (torch.outer(l1.double(), l2.double()) + float(delta))
    out += torch.einsum("ncij,nkij,ij->nck", Z, Z, w).to(out.dtype)
    return out


def install(monkeypatch):
    """Replace the native wrappers by the emulation and lift the CUDA-only guards (tests only)."""
    from laplace_b200 import backend

    for name in ("pack_rows", "pack_conv", "pack_nchw", "gemm_nt", "shared_weight_contract", "jac_linear_write",
                 "ll_jacobian_write", "batched_pair_dot", "ll_ggn_expand", "ll_sigma_gather", "eigh_jacobi",
                 "pack_conv_rows", "pack_nchw_rows", "pack_cast", "pack_cast_fused", "col2im", "col2im_nhwc", "syrk_conv_patches", "diag_conv_sq", "conv_bwd_strided", "conv_nhwc", "gemm_tn", "scale_channels", "relu_bwd", "maxpool2d_bwd",
                 "kron_conv_quadform", "maxpool2d_bwd_pack"):
        monkeypatch.setattr(K, name, globals()[name])
    monkeypatch.setattr(K, "alloc_packed", _alloc)
    monkeypatch.setattr(K, "alloc_rows", _alloc)
    monkeypatch.setattr(backend._B200Mixin, "_device_check", lambda self, t: None)
    from laplace_b200 import conv_engine

    monkeypatch.setattr(conv_engine, "_ALLOW_CPU", True)


# ---- convolution engine operands (emulation) ------------------------------------------------
def pack_conv_rows(x, mod, kind):
    cols = F.unfold(x.float(), mod.kernel_size, dilation=mod.dilation, padding=mod.padding, stride=mod.stride)
    N, d, T = cols.shape
    out = _alloc(N * T, d, kind, x.device)
    out.hi[:, :d] = cols.permute(0, 2, 1).reshape(N * T, d)
    return out


def pack_nchw_rows(g, kind):
    Q, Cc, HW = g.shape
    out = _alloc(Q * HW, Cc, kind, g.device)
    out.hi[:, :Cc] = g.float().permute(0, 2, 1).reshape(Q * HW, Cc)
    return out


def pack_cast(src, kind):
    out = _alloc(src.shape[0], src.shape[1], kind, src.device)
    out.hi[:, :src.shape[1]] = src.float()
    return out


def pack_cast_fused(src, kind, scale=None, y=None):
    v = src.float()
    if scale is not None:
        v = v * scale.view(1, -1)
    if y is not None:
        v = (v.reshape(-1, y.shape[0], y.shape[1]) * (y > 0)).reshape(v.shape)
    return pack_cast(v, kind)


def col2im(Dc, in_shape, mod):
    Q, C, H, W = in_shape
    OH, OW = K.conv_out_hw(in_shape, mod)
    cols = Dc[:, :Q * OH * OW].reshape(Dc.shape[0], Q, OH * OW).permute(1, 0, 2)
    return F.fold(cols, (H, W), mod.kernel_size, dilation=mod.dilation, padding=mod.padding, stride=mod.stride)


def syrk_conv_patches(X, Q, H, W, mod, out, alpha=1.0):
    Ci = X.K
    x = X.hi[:, :Ci].reshape(Q, H, W, Ci).permute(0, 3, 1, 2).float()
    P = F.unfold(x, mod.kernel_size, padding=mod.padding).transpose(1, 2).reshape(Q * H * W, -1)
    out += alpha * (P.t() @ P)
    return out


def diag_conv_sq(G, X, Nimg, H, W, mod, out, alpha=1.0):
    Ci, Co = X.K, G.K
    x = X.hi[:, :Ci].reshape(Nimg, H, W, Ci).permute(0, 3, 1, 2).float()
    P = F.unfold(x, mod.kernel_size, padding=mod.padding).transpose(1, 2)          # [Nimg, T, Ci*KK]
    g = G.hi[:, :Co].reshape(-1, Nimg, H * W, Co).float()                          # [cols, Nimg, T, Co]
    per = torch.einsum("cnto,ntp->cnop", g, P)
    out += alpha * (per * per).sum((0, 1))
    return out


def conv_bwd_strided(G, Q, OH, OW, Wt, mod, in_shape):
    _, Ci, H, W = in_shape
    Co = G.K
    kh, kw = mod.kernel_size
    g = G.hi[:, :Co].reshape(Q, OH, OW, Co).permute(0, 3, 1, 2).float()
    w = Wt.hi[:, :Co].reshape(kh, kw, Ci, Co).permute(3, 2, 0, 1).float()          # [Co, Ci, kh, kw]
    out = torch.nn.grad.conv2d_input((Q, Ci, H, W), w, g, mod.stride, mod.padding, mod.dilation)
    return out.contiguous(memory_format=torch.channels_last)


def col2im_nhwc(Dc, in_shape, mod):
    Q, C, H, W = in_shape
    OH, OW = K.conv_out_hw(in_shape, mod)
    kh, kw = mod.kernel_size
    cols = Dc[:, :kh * kw * C].reshape(Q, OH * OW, kh * kw, C).permute(0, 3, 2, 1).reshape(Q, C * kh * kw, OH * OW)
    out = F.fold(cols, (H, W), mod.kernel_size, dilation=mod.dilation, padding=mod.padding, stride=mod.stride)
    return out.contiguous(memory_format=torch.channels_last)


def conv_nhwc(X, Q, H, W, Wt, N, KH, KW, base_h, base_w, sgn, out, alpha=1.0):
    Kc = X.K
    x = X.hi[:, :Kc].reshape(Q, H, W, Kc)
    w = Wt.hi[:, :Kc].reshape(KH * KW, N, Kc)
    res = torch.zeros(Q, H, W, N, dtype=x.dtype)
    for kh in range(KH):
        for kw in range(KW):
            dh, dw = base_h + sgn * kh, base_w + sgn * kw
            shifted = torch.zeros_like(x)
            hs, he = max(0, -dh), min(H, H - dh)
            ws, we = max(0, -dw), min(W, W - dw)
            if hs < he and ws < we:
                shifted[:, hs:he, ws:we] = x[:, hs + dh:he + dh, ws + dw:we + dw]
            res += torch.einsum("qhwk,nk->qhwn", shifted, w[kh * KW + kw])
    out.copy_(alpha * res.reshape(Q * H * W, N))
    return out


def gemm_tn(A, B, out, alpha=1.0, accumulate=True, symmetric=False):
    res = alpha * (A.hi[:, :A.K].float().t() @ B.hi[:, :B.K].float())
    if accumulate:
        out += res
    else:
        out.copy_(res)
    return out


def scale_channels(g, scale):
    return g * scale.view(1, -1, 1, 1)


def relu_bwd(g, y, reps):
    B = y.shape[0]
    return (g.reshape(reps, B, *y.shape[1:]) * (y > 0)).reshape(g.shape)


def maxpool2d_bwd(g, idx, in_shape, k, s, p):
    Q, C, OH, OW = g.shape
    H, W = in_shape[-2:]
    Nb = idx.shape[0]
    out = torch.zeros(Q, C, H * W, dtype=torch.float32)
    ii = idx.reshape(Nb, C, OH * OW).repeat(Q // Nb, 1, 1)
    out.scatter_add_(2, ii, g.reshape(Q, C, OH * OW).float())
    out = out.reshape(Q, C, H, W)
    # the native wrapper keeps channels-last gradients channels-last
    return out.contiguous(memory_format=torch.channels_last) if not g.is_contiguous() else out
