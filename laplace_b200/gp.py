"""GP / NTK kernels ``K = J J^T`` between batches of Jacobians without the dense ``(B, C, P)`` tensors.

The reference's ``FunctionalLaplace`` builds its kernel matrices from dense Jacobians with einsums
(baselaplace.py:3026-3122: ``_kernel_batch``, ``_kernel_star``, ``_kernel_batch_star``).  With the per-layer structure
of the B200 backend's Jacobians (``JacobianFactors``) the same matrices are sums over parameter blocks of

* ``J = g (x) a``  (no weight sharing):  ``K[(a,c),(b,e)] += <g1_ac, g2_be> * <a1_a, a2_b>``  -- two small GEMMs,
* bias ``J = g``:                         ``K += <g1_ac, g2_be>``,
* ``J_nc = G_nc^T A_n`` (convolutions):   ``K += sum_tt' <g1_act, g2_bet'> <a1_at, a2_bt'>`` when ``T^2`` is small, the
  block's dense rows otherwise,

all on the fp32 GEMM kernel.  ``FunctionalLaplace`` itself hard-asserts its backend (baselaplace.py:2225), so it cannot
take ``B200GGN`` without a one-line host change; these functions return exactly the tensors its three
kernel methods return, for a caller that relaxes the assert.
"""
from __future__ import annotations

import torch

from . import kernels as K
from .matrix import JacobianFactors, materialize_block


def _gemm(A2d: torch.Tensor, Bnk: torch.Tensor) -> torch.Tensor:
    A2d, Bnk = A2d.contiguous(), Bnk.contiguous()
    out = torch.empty(A2d.shape[0], Bnk.shape[0], device=A2d.device, dtype=torch.float32)
    return K.gemm_nt(K.Packed(A2d, None, K.F32, A2d.shape[0], A2d.shape[1]),
                     K.Packed(Bnk, None, K.F32, Bnk.shape[0], Bnk.shape[1]), out, 1.0, accumulate=False)


def _factors(Js) -> JacobianFactors:
    fac = getattr(Js, "_lpb_factors", None)
    if fac is None:
        dense = Js.float().contiguous()
        fac = JacobianFactors([("dense", dense)], dense.shape[0], dense.shape[1], [dense.shape[2]])
    return fac


def kernel4(Js1, Js2=None) -> torch.Tensor:
    """``K[a, c, b, e] = sum_p Js1[a, c, p] Js2[b, e, p]`` (``Js2 = Js1`` if omitted), fp32."""
    f1 = _factors(Js1)
    f2 = f1 if Js2 is None else _factors(Js2)
    if len(f1.blocks) != len(f2.blocks):
        f1 = JacobianFactors([("dense", torch.cat([materialize_block(b, f1.n_batch, f1.n_out) for b in f1.blocks], 2))],
                             f1.n_batch, f1.n_out, [sum(f1.sizes)])
        f2 = JacobianFactors([("dense", torch.cat([materialize_block(b, f2.n_batch, f2.n_out) for b in f2.blocks], 2))],
                             f2.n_batch, f2.n_out, [sum(f2.sizes)])
    N1, N2, C = f1.n_batch, f2.n_batch, f1.n_out
    dev = f1.blocks[0][1].device
    Kout = torch.zeros(N1, C, N2, C, device=dev, dtype=torch.float32)
    for b1, b2 in zip(f1.blocks, f2.blocks):
        kind = b1[0] if b1[0] == b2[0] else "dense"
        if kind == "outer":
            g1, a1, g2, a2 = b1[1], b1[2], b2[1], b2[2]                       # [C, N, d_out], [N, d_in]
            GG = _gemm(g1.reshape(C * N1, -1), g2.reshape(C * N2, -1)).view(C, N1, C, N2)
            AA = _gemm(a1, a2)                                                # [N1, N2]
            Kout += GG.permute(1, 0, 3, 2) * AA.view(N1, 1, N2, 1)
        elif kind == "vec":
            g1, g2 = b1[1], b2[1]
            Kout += _gemm(g1.reshape(C * N1, -1), g2.reshape(C * N2, -1)).view(C, N1, C, N2).permute(1, 0, 3, 2)
        elif kind == "conv" and b1[3] * b2[3] * (C * C * b1[1].shape[1] + b1[2].shape[1]) < C * C * b1[1].shape[1] * b1[2].shape[1]:
            G1, A1, T1 = b1[1], b1[2], b1[3]                                  # [(c,n,t), d_out], [(n,t), d_in]
            G2, A2, T2 = b2[1], b2[2], b2[3]
            GG = _gemm(G1, G2).view(C, N1, T1, C, N2, T2)
            AA = _gemm(A1, A2).view(N1, T1, N2, T2)
            Kout += torch.einsum("catebu,atbu->acbe", GG, AA)
        else:
            J1 = materialize_block(b1, N1, C).reshape(N1 * C, -1)
            J2 = J1 if (f2 is f1) else materialize_block(b2, N2, C).reshape(N2 * C, -1)
            Kout += _gemm(J1, J2).view(N1, C, N2, C)
    return Kout


def kernel_batch(Js1, Js2=None, independent_outputs: bool = False) -> torch.Tensor:
    """``FunctionalLaplace._kernel_batch`` (baselaplace.py:3026-3058): ``(b*C, b2*C)``, or ``(b, b2, C)`` per-output."""
    K4 = kernel4(Js1, Js2)
    N1, C, N2, _ = K4.shape
    if independent_outputs:
        return torch.diagonal(K4, dim1=1, dim2=3).contiguous()              # [N1, N2, C]
    return K4.reshape(N1 * C, N2 * C)


def kernel_star(Js, joint: bool = False, independent_outputs: bool = False) -> torch.Tensor:
    """``FunctionalLaplace._kernel_star`` (baselaplace.py:3060-3090)."""
    K4 = kernel4(Js)
    N, C = K4.shape[0], K4.shape[1]
    if joint:
        if independent_outputs:                                              # einsum "acp,bcp->abcc"
            return torch.diag_embed(torch.diagonal(K4, dim1=1, dim2=3))
        return K4.permute(0, 2, 1, 3).contiguous()                           # "acp,bep->abce"
    idx = torch.arange(N, device=K4.device)
    Kd = K4[idx, :, idx, :]                                                  # [N, C, C]
    return torch.diagonal(Kd, dim1=1, dim2=2).contiguous() if independent_outputs else Kd.contiguous()


def kernel_batch_star(Js1, Js2, independent_outputs: bool = False) -> torch.Tensor:
    """``FunctionalLaplace._kernel_batch_star`` (baselaplace.py:3092-3122): ``(b1, b2, C, C)`` or ``(b1, b2, C)``."""
    K4 = kernel4(Js1, Js2)
    if independent_outputs:
        return torch.diagonal(K4, dim1=1, dim2=3).contiguous()
    return K4.permute(0, 2, 1, 3).contiguous()
