"""Data-parallel ``fit()``: one process per GPU, batches sharded across ranks, ONE exchange of the accumulated
curvature at the end.  Every curvature structure of the path is a plain sum over data
(baselaplace.py:984-985), the model is replicated, so there is no data-path collective inside the loop.

``N`` passed to the backend must stay the GLOBAL dataset size (the ``M/N`` rescale of the KFAC ``A`` factors,
curvature/curvlinops.py:46-53, baselaplace.py:964) -- ``ShardedLoader`` keeps ``len(loader.dataset)`` global.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .matrix import B200Kron


class ShardedLoader:
    """Rank ``r`` of ``world`` sees batches ``r, r + world, ...`` of an underlying loader; ``.dataset`` is
    the full dataset so that ``len(loader.dataset)`` is the global ``N``.  No sample is repeated or dropped
    (unlike ``DistributedSampler`` padding), so the sharded sum equals the single-process sum."""

    def __init__(self, loader, rank: int, world: int):
        self.loader, self.rank, self.world = loader, rank, world
        self.dataset = loader.dataset

    def __iter__(self):
        for i, batch in enumerate(self.loader):
            if i % self.world == self.rank:
                yield batch

    def __len__(self):
        n = len(self.loader)
        return (n - self.rank + self.world - 1) // self.world


def allreduce_curvature(H, loss=None, group=None):
    """Sum the accumulated curvature over ranks in place: a ``B200Kron`` is reduced through its single flat
    fp32 buffer (one NCCL launch), dense/diagonal curvature as the tensor itself.  ``loss`` (0-dim tensor) is
    reduced alongside.  Works with the ``gloo`` backend on CPU for tests."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return H, loss
    if isinstance(H, B200Kron) and H._flat is not None:
        dist.all_reduce(H._flat, op=dist.ReduceOp.SUM, group=group)
    elif hasattr(H, "kfacs"):
        for F in H.kfacs:
            for Hi in F:
                dist.all_reduce(Hi, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(H, op=dist.ReduceOp.SUM, group=group)
    if loss is not None and torch.is_tensor(loss):
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    return H, loss


def decompose_sharded(kron, damping: bool = False, group=None):
    """``Kron.decompose`` with the factors partitioned over the ranks (greedy balance on ``n^3``): every rank
    eigendecomposes only its share, one all-reduce of a zero-padded flat ``Q`` buffer (+ eigenvalues) replicates the
    result (SURVEY 8(e) "exchange v2").  Falls back to the local decomposition for a single process."""
    from .matrix import B200KronDecomposed

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return kron.decompose(damping=damping)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mats = [(i, j, H) for i, F in enumerate(kron.kfacs) for j, H in enumerate(F)]
    order = sorted(range(len(mats)), key=lambda k: -mats[k][2].shape[0])
    load, owner = [0.0] * world, [0] * len(mats)
    for k in order:
        r = min(range(world), key=lambda q: load[q])
        owner[k] = r
        load[r] += float(mats[k][2].shape[0]) ** 3
    # local decomposition of the owned factors, through the same code path as the single-process one
    mine = [k for k in range(len(mats)) if owner[k] == rank]
    sub = B200Kron([[mats[k][2]] for k in mine]) if mine else None
    local = sub.decompose(damping=damping) if sub is not None else None
    dev, dt = mats[0][2].device, mats[0][2].dtype
    sizes = [m[2].shape[0] for m in mats]
    Qflat = torch.zeros(sum(n * n for n in sizes), device=dev, dtype=dt)
    Lflat = torch.zeros(sum(sizes), device=dev, dtype=dt)
    qoff, loff, offs = 0, 0, []
    for n in sizes:
        offs.append((qoff, loff))
        qoff += n * n
        loff += n
    for pos, k in enumerate(mine):
        n = sizes[k]
        Qflat[offs[k][0]:offs[k][0] + n * n].copy_(local.eigenvectors[pos][0].reshape(-1))
        Lflat[offs[k][1]:offs[k][1] + n].copy_(local.eigenvalues[pos][0])
    dist.all_reduce(Qflat, group=group)
    dist.all_reduce(Lflat, group=group)
    eigvecs = [[None] * len(F) for F in kron.kfacs]
    eigvals = [[None] * len(F) for F in kron.kfacs]
    for k, (i, j, H) in enumerate(mats):
        n = sizes[k]
        eigvecs[i][j] = Qflat[offs[k][0]:offs[k][0] + n * n].view(n, n)
        eigvals[i][j] = Lflat[offs[k][1]:offs[k][1] + n]
    return B200KronDecomposed(eigvecs, eigvals, damping=damping)


def fit_distributed(la, train_loader, group=None):
    """``B200Laplace.fit`` sharded over the ranks of ``group``: local accumulation, one all-reduce, then the
    (replicated) decomposition.  Returns ``la``."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    la.fit(ShardedLoader(train_loader, rank, world), decompose=False)
    H = la.H_facs if la.structure == "kron" else la.H
    loss = la.loss if torch.is_tensor(la.loss) else None
    H, loss = allreduce_curvature(H, loss, group)
    if loss is not None:
        la.loss = loss
    if la.structure == "kron":
        la.H_facs = H
        la.H = decompose_sharded(H, damping=la.damping, group=group)
    else:
        la.H = H
    return la
