"""Data-parallel ``fit()``: one process per GPU, batches sharded across ranks, ONE exchange of the accumulated
curvature at the end.  Every curvature structure of the path is a plain sum over data
(baselaplace.py:984-985), the model is replicated, so there is no data-path collective inside the loop.

``N`` passed to the backend must stay the GLOBAL dataset size (the ``M/N`` rescale of the KFAC ``A`` factors,
curvature/curvlinops.py:46-53, baselaplace.py:964) -- ``ShardedLoader`` keeps ``len(loader.dataset)`` global.

Exchange (SURVEY 8(e) "v2"): all-reduce of the flat factor buffer, then every rank eigendecomposes only the factors it
owns (greedy balance on the measured cost of the live block) and the eigenvectors / eigenvalues are replicated with ONE all-gather of equally sized
per-rank slabs -- each byte of ``Q`` crosses NVLink once (the round-1 version all-reduced a zero-padded buffer: world
times the traffic, and the reduction arithmetic on top).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .matrix import B200Kron


class ShardedLoader:
    """Rank ``r`` of ``world`` sees batches ``r, r + world, ...`` of an underlying loader; ``.dataset`` is
    the full dataset so that ``len(loader.dataset)`` is the global ``N``.  No sample is repeated or dropped
    (unlike ``DistributedSampler`` padding), so the sharded sum equals the single-process sum.

    Correct only if every rank iterates the SAME batch sequence (``shuffle=False``, or a generator seeded identically on
    all ranks): ``fingerprint`` (batch count + a checksum of the first batch, filled in by iterating) is compared
    across ranks by ``fit_distributed`` and a mismatch raises instead of silently duplicating / dropping samples."""

    def __init__(self, loader, rank: int, world: int):
        self.loader, self.rank, self.world = loader, rank, world
        self.dataset = loader.dataset
        self.fingerprint = None

    @staticmethod
    def _checksum(batch) -> float:
        x = batch[0] if isinstance(batch, (tuple, list)) else next(iter(batch.values()))
        if not torch.is_tensor(x) or x.numel() == 0:
            return 0.0
        return float(x.reshape(-1)[:4096].double().sum())

    def __iter__(self):
        n, first = 0, 0.0
        for i, batch in enumerate(self.loader):
            if i == 0:
                first = self._checksum(batch)
            n += 1
            if i % self.world == self.rank:
                yield batch
        self.fingerprint = (float(n), first)

    def __len__(self):
        n = len(self.loader)
        return (n - self.rank + self.world - 1) // self.world


def _active(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def check_same_order(loader: ShardedLoader, device, group=None) -> None:
    """Raise if the ranks did not iterate the same batch sequence (see ``ShardedLoader``)."""
    if not _active(group) or loader.fingerprint is None:
        return
    world = dist.get_world_size(group)
    mine = torch.tensor(loader.fingerprint, device=device, dtype=torch.float64)
    allf = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allf, mine, group=group)
    for r, other in enumerate(allf):
        if not torch.equal(other, allf[0]):
            raise RuntimeError("ShardedLoader: rank %d iterated a different batch sequence than rank 0 (shuffle without a "
                               "shared seed?); the sharded fit would duplicate or drop samples" % r)


def allreduce_curvature(H, loss=None, group=None):
    """Sum the accumulated curvature over ranks in place: a ``B200Kron`` is reduced through its single flat
    fp32 buffer (one NCCL launch), dense/diagonal curvature as the tensor itself.  ``loss`` (0-dim tensor) is
    reduced alongside.  Works with the ``gloo`` backend on CPU for tests."""
    if not _active(group):
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


def factor_owners(sizes, world: int, cost=None):
    """Greedy balance of the eigendecomposition cost over ``world`` ranks, most expensive first (``cost(n)``: measured
    milliseconds by size, ``matrix.eigh_cost_ms``; ``sizes`` are the LIVE sizes -- a 4608-row input factor of a 3x3
    convolution on a 1x1 map costs what its 460 live coordinates cost).  Deterministic: every rank computes the same table."""
    from .matrix import eigh_cost_ms

    cost = cost or eigh_cost_ms
    w = [float(cost(int(n))) for n in sizes]
    order = sorted(range(len(sizes)), key=lambda k: (-w[k], k))
    load, owner = [0.0] * world, [0] * len(sizes)
    for k in order:
        r = min(range(world), key=lambda q: (load[q], q))
        owner[k] = r
        load[r] += w[k]
    return owner


def decompose_sharded(kron, damping: bool = False, group=None):
    """``Kron.decompose`` with the factors partitioned over the ranks: every rank eigendecomposes only its share through
    the same code path as the single-process ``decompose`` and ONE all-gather of per-rank slabs (``Q`` and ``lambda`` of
    the owned factors back to back, padded to the largest slab) replicates the result.  Falls back to the local
    decomposition for a single process."""
    from .matrix import B200KronDecomposed

    if not _active(group):
        return kron.decompose(damping=damping)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mats = [(i, j, H) for i, F in enumerate(kron.kfacs) for j, H in enumerate(F)]
    from .matrix import live_sizes

    sizes = [int(m[2].shape[0]) for m in mats]
    owner = factor_owners(live_sizes([m[2] for m in mats]), world)    # all ranks hold identical (all-reduced) factors
    owned = [[k for k in range(len(mats)) if owner[k] == r] for r in range(world)]
    slab = max(sum(sizes[k] * (sizes[k] + 1) for k in ks) for ks in owned)     # n*n eigenvectors + n eigenvalues each
    dev, dt = mats[0][2].device, mats[0][2].dtype
    mine = owned[rank]
    local = B200Kron([[mats[k][2]] for k in mine]).decompose(damping=damping) if mine else None
    send = torch.zeros(slab, device=dev, dtype=dt)
    off = 0
    for pos, k in enumerate(mine):
        n = sizes[k]
        send[off:off + n * n].copy_(local.eigenvectors[pos][0].reshape(-1))
        send[off + n * n:off + n * n + n].copy_(local.eigenvalues[pos][0])
        off += n * (n + 1)
    recv = torch.empty(world * slab, device=dev, dtype=dt)
    dist.all_gather_into_tensor(recv, send, group=group)
    eigvecs = [[None] * len(F) for F in kron.kfacs]
    eigvals = [[None] * len(F) for F in kron.kfacs]
    for r, ks in enumerate(owned):
        off = r * slab
        for k in ks:
            i, j, _ = mats[k]
            n = sizes[k]
            eigvecs[i][j] = recv[off:off + n * n].view(n, n)
            eigvals[i][j] = recv[off + n * n:off + n * n + n]
            off += n * (n + 1)
    return B200KronDecomposed(eigvecs, eigvals, damping=damping)


def fit_distributed(la, train_loader, group=None):
    """``B200Laplace.fit`` sharded over the ranks of ``group``: local accumulation, one all-reduce, then the sharded
    decomposition.  A rank whose shard is empty contributes zeros.  Returns ``la``."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    shard = ShardedLoader(train_loader, rank, world)
    la.fit(shard, decompose=False)
    check_same_order(shard, la._device, group)
    H = la.H_facs if la.structure == "kron" else la.H
    if H is None:
        H = la.zero_curvature()
    loss = la.loss if torch.is_tensor(la.loss) else torch.zeros((), device=la._device)
    H, loss = allreduce_curvature(H, loss, group)
    la.loss = loss
    if la.structure == "kron":
        la.H_facs = H
        la.H = decompose_sharded(H, damping=la.damping, group=group)
    else:
        la.H = H
    return la
