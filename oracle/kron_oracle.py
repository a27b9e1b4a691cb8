"""CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT) for the posterior-side path:
eigendecomposition of Kronecker factors, the decomposed-Kron quadratic forms and the
Full / Diag GLM predictive variances.  Plain torch on CPU, dense and slow on purpose.
"""
from __future__ import annotations

import math

import torch


def symeig(M: torch.Tensor):
    """``laplace.utils.utils.symeig`` (utils/utils.py:193-228): ``eigh(UPLO='U')``, on failure once more on ``M + I``
    with the jitter removed from the eigenvalues (:209-216; LAPACK's divide-and-conquer does give up on some rank-deficient
    factors, e.g. a 576x576 input factor with 46 live coordinates), eigenvalues clamped at 0, NaNs zeroed."""
    try:
        L, W = torch.linalg.eigh(M, UPLO="U")
    except RuntimeError:
        L, W = torch.linalg.eigh(M + torch.eye(M.shape[0], dtype=M.dtype, device=M.device), UPLO="U")
        L = L - 1.0
    return torch.nan_to_num(L.clamp(min=0.0)), torch.nan_to_num(W)


def decompose(kfacs):
    """``Kron.decompose`` (utils/matrix.py:123-150): per-factor ``symeig``.
    Returns ``(eigenvectors, eigenvalues)`` as lists of lists."""
    Qs, ls = [], []
    for F in kfacs:
        pairs = [symeig(Hi) for Hi in F]
        ls.append([p[0] for p in pairs])
        Qs.append([p[1] for p in pairs])
    return Qs, ls


def scale_eigenvalues(eigvals, scalar: float):
    """``KronDecomposed.__mul__`` (utils/matrix.py:357-376): every factor's eigenvalues
    times ``scalar ** (1/len(block))``."""
    return [[(scalar ** (1.0 / len(ls))) * l for l in ls] for ls in eigvals]


def _block_spectrum(ls, delta, damping: bool):
    if len(ls) == 1:
        return ls[0] + delta
    l1, l2 = ls
    if damping:
        sd = math.sqrt(float(delta))
        return torch.outer(l1 + sd, l2 + sd)
    return torch.outer(l1, l2) + delta


def _deltas(deltas, n_blocks, dtype):
    d = torch.as_tensor(deltas, dtype=dtype).reshape(-1)
    return d.expand(n_blocks) if d.numel() == 1 else d


def kron_bmm(eigvecs, eigvals, deltas, W: torch.Tensor, exponent: float = -1.0, damping: bool = False):
    """``KronDecomposed._bmm`` (utils/matrix.py:406-456): ``(Kron + delta)^exponent @ W``
    for ``W (B, K, P)``; block of a weight ``(out, in)`` is ``Q1 (x) Q2`` with the
    flattened parameter reshaped row-major to ``(len(l1), len(l2))``."""
    B, K, P = W.shape
    Wf = W.reshape(B * K, P)
    d = _deltas(deltas, len(eigvals), W.dtype)
    out, cur = [], 0
    for Qs, ls, delta in zip(eigvecs, eigvals, d):
        spec = _block_spectrum(ls, delta, damping) ** exponent
        if len(ls) == 1:
            p = ls[0].numel()
            Wp = Wf[:, cur:cur + p]
            out.append(((Wp @ Qs[0]) * spec) @ Qs[0].T)
        else:
            p1, p2 = ls[0].numel(), ls[1].numel()
            p = p1 * p2
            Wp = Wf[:, cur:cur + p].reshape(-1, p1, p2)
            Z = torch.einsum("ia,nij,jb->nab", Qs[0], Wp, Qs[1]) * spec
            out.append(torch.einsum("ia,nab,jb->nij", Qs[0], Z, Qs[1]).reshape(-1, p))
        cur += p
    return torch.cat(out, 1).reshape(B, K, P)


def kron_inv_square_form(eigvecs, eigvals, deltas, W, damping: bool = False):
    """``KronDecomposed.inv_square_form`` (utils/matrix.py:458-461): ``W P^{-1} W^T`` per
    batch element -> ``(B, K, K)``."""
    SW = kron_bmm(eigvecs, eigvals, deltas, W, -1.0, damping)
    return torch.bmm(W, SW.transpose(1, 2))


def kron_logdet(eigvals, deltas, damping: bool = False):
    """``KronDecomposed.logdet`` (utils/matrix.py:381-404)."""
    d = _deltas(deltas, len(eigvals), eigvals[0][0].dtype)
    return sum(torch.log(_block_spectrum(ls, delta, damping)).sum() for ls, delta in zip(eigvals, d))


def kron_dense(eigvecs, eigvals, deltas, exponent: float = 1.0, damping: bool = False):
    """``KronDecomposed.to_matrix`` (utils/matrix.py:524-556)."""
    d = _deltas(deltas, len(eigvals), eigvals[0][0].dtype)
    blocks = []
    for Qs, ls, delta in zip(eigvecs, eigvals, d):
        spec = (_block_spectrum(ls, delta, damping) ** exponent).reshape(-1)
        Q = Qs[0] if len(ls) == 1 else torch.kron(Qs[0], Qs[1])
        blocks.append((Q * spec) @ Q.T)
    return torch.block_diag(*blocks)


def full_posterior_covariance(H: torch.Tensor, prior_precision_diag: torch.Tensor, H_factor: float = 1.0):
    """``FullLaplace.posterior_precision`` / ``posterior_scale`` / ``posterior_covariance``
    (baselaplace.py:1634-1673) with ``invsqrt_precision`` = torch's
    ``_precision_to_scale_tril`` (utils/utils.py:118-129): Cholesky of the flipped
    precision, flipped back, triangular inverse; ``Sigma = L L^T``."""
    Pm = H_factor * H + torch.diag(prior_precision_diag)
    Lf = torch.linalg.cholesky(torch.flip(Pm, (-2, -1)))
    L_inv = torch.transpose(torch.flip(Lf, (-2, -1)), -2, -1)
    eye = torch.eye(Pm.shape[-1], dtype=Pm.dtype)
    L = torch.linalg.solve_triangular(L_inv, eye, upper=False)
    return L @ L.T


def full_functional_variance(Js, Sigma):
    """``FullLaplace.functional_variance`` (baselaplace.py:1683-1684): ``J_n Sigma J_n^T``."""
    return torch.stack([Jn @ Sigma @ Jn.T for Jn in Js])


def diag_functional_variance(Js, post_var):
    """``DiagLaplace.functional_variance`` (baselaplace.py:2113-2115)."""
    return torch.stack([(Jn * post_var) @ Jn.T for Jn in Js])


def probit_predictive(f_mu, f_var):
    """``_glm_forward_call`` probit link (baselaplace.py:662-664)."""
    kappa = 1.0 / torch.sqrt(1.0 + math.pi / 8.0 * torch.diagonal(f_var, dim1=1, dim2=2))
    return torch.softmax(kappa * f_mu, dim=-1)


def gridsearch_prior_precision(eigvecs, eigvals, Js_batches, f_batches, y_batches, interval, H_factor: float = 1.0,
                               running_metric: bool = False):
    """CPU restatement of ``BaseLaplace._gridsearch`` (baselaplace.py:516-561) + ``utils.validate``
    (utils/utils.py:39-101) for a Kron posterior, classification, probit GLM predictive, ``RunningNLLMetric``
    (utils/metrics.py): per grid value the NLL of the probit predictive over all validation batches.  With
    ``running_metric`` the metric is never reset between grid values, exactly like the reference's loop.
    Returns ``(best, losses)``."""
    ls = scale_eigenvalues(eigvals, H_factor)
    losses, tot, cnt = [], 0.0, 0
    for pp in interval:
        delta = torch.as_tensor(pp, dtype=torch.float64)
        s, n = 0.0, 0
        for Js, f, y in zip(Js_batches, f_batches, y_batches):
            fv = kron_inv_square_form(eigvecs, ls, delta, Js)
            probs = probit_predictive(f, fv)
            s += float(torch.nn.functional.nll_loss(probs.log(), y, reduction="sum"))
            n += len(y)
        if running_metric:
            tot, cnt = tot + s, cnt + n
            losses.append(tot / cnt)
        else:
            losses.append(s / n)
    losses = torch.tensor(losses, dtype=torch.float64)
    return interval[int(torch.argmin(losses))], losses
