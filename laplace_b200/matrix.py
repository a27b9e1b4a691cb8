"""Kronecker-factored containers whose arithmetic runs on the B200 kernels.

``B200Kron`` / ``B200KronDecomposed`` derive from the reference's ``Kron`` / ``KronDecomposed``
(utils/matrix.py:16-560) when the reference is importable, so ``la.H += H_batch``
(baselaplace.py:985), ``H_facs.decompose(...)`` (baselaplace.py:1809), ``self.H * self._H_factor +
self.prior_precision`` (baselaplace.py:1820) and ``posterior_precision.inv_square_form(Js)``
(baselaplace.py:1834-1835) dispatch here without touching host code:

* ``la.H`` starts as a plain reference ``Kron`` of zeros; the first ``la.H += H_batch`` resolves to
  the subclass' reflected ``__radd__`` (Python prefers the reflected method of a subclass operand),
  so ``la.H`` is a ``B200Kron`` from the first batch on and later batches use ``__iadd__`` (one
  fused in-place add over the flat factor buffer).
* ``decompose`` -> batched Jacobi kernel for factors up to 128x128, cuSOLVER (``torch.linalg.eigh``,
  library -- see DESIGN.md) beyond.
* ``inv_square_form`` -> structured eigenbasis quadratic form (SURVEY App. A "Structure the kernels
  can exploit") when the Jacobian carries its per-layer factors, dense rotation GEMMs otherwise.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch

from . import kernels as K
from .interface import Kron, KronDecomposed


def _is_scalar(s) -> bool:
    if isinstance(s, (int, float)):
        return True
    return torch.is_tensor(s) and s.numel() == 1 and s.ndim <= 1


def _root(s, k: int):
    """``s ** (1/k)`` for a Python number or a (possibly autograd-tracked) 0-/1-element tensor."""
    if torch.is_tensor(s):
        return torch.pow(s.reshape(()), 1.0 / k)
    return math.pow(s, 1.0 / k)


def _as_f32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.float32 else t.float()


# =========================================================================================
class B200Kron(Kron):
    """Per-parameter Kronecker factors (``kfacs[i]`` = ``[B, A]`` for a weight, ``[B]`` for a bias,
    ``parameters()`` order) stored as views into one flat device buffer."""

    def __init__(self, kfacs, flat: torch.Tensor | None = None):
        super().__init__(kfacs)
        self._flat = flat

    # -- construction ---------------------------------------------------------------------
    @classmethod
    def zeros(cls, dims: Sequence[Sequence[int]], device, dtype=torch.float32) -> "B200Kron":
        total = sum(d * d for F in dims for d in F)
        flat = torch.zeros(total, device=device, dtype=dtype)
        kfacs, off = [], 0
        for F in dims:
            blk = []
            for d in F:
                blk.append(flat[off:off + d * d].view(d, d))
                off += d * d
            kfacs.append(blk)
        return cls(kfacs, flat)

    @classmethod
    def from_kfacs(cls, kfacs) -> "B200Kron":
        """Copy a list-of-lists of square factors (``state_dict()["H"]`` of a ``KronLaplace``, baselaplace.py:1867-1871)
        into one flat device buffer."""
        ref = kfacs[0][0]
        if any(H.ndim != 2 for F in kfacs for H in F) or ref.dtype != torch.float32:
            return cls([[H for H in F] for F in kfacs])
        out = cls.zeros([[int(H.shape[0]) for H in F] for F in kfacs], ref.device, ref.dtype)
        for Fo, Fi in zip(out.kfacs, kfacs):
            for Ho, Hi in zip(Fo, Fi):
                Ho.copy_(Hi)
        return out

    def dims(self):
        return [[int(H.shape[0]) for H in F] for F in self.kfacs]

    def _same_layout(self, other) -> bool:
        return (isinstance(other, B200Kron) and self._flat is not None and other._flat is not None
                and self._flat.shape == other._flat.shape and self._flat.dtype == other._flat.dtype
                and self.dims() == other.dims())

    # -- reference Kron.__add__ (utils/matrix.py:79-98): zip-aligned factor-wise sum ---------
    def __add__(self, other):
        if not isinstance(other, Kron):
            raise ValueError("Can only add Kron to Kron.")
        if self._same_layout(other):
            out = B200Kron.zeros(self.dims(), self._flat.device, self._flat.dtype)
            torch.add(self._flat, other._flat, out=out._flat)
            return out
        if self._flat is not None and len(self.kfacs) == len(other.kfacs) and all(
                len(Fi) == len(Fj) and all(Hi.shape == Hj.shape for Hi, Hj in zip(Fi, Fj))
                for Fi, Fj in zip(self.kfacs, other.kfacs)):
            # the other operand is a plain reference ``Kron`` (``la.H`` before the first batch, baselaplace.py:985): the
            # sum lives in a fresh flat buffer so that later batches take the fused ``__iadd__`` / single all-reduce
            out = B200Kron.zeros(self.dims(), self._flat.device, self._flat.dtype)
            for Fo, Fi, Fj in zip(out.kfacs, self.kfacs, other.kfacs):
                for Ho, Hi, Hj in zip(Fo, Fi, Fj):
                    torch.add(Hi, Hj.to(Hi.dtype) if Hj.dtype != Hi.dtype else Hj, out=Ho)
            return out
        kfacs = [[Hi.add(Hj) for Hi, Hj in zip(Fi, Fj)] for Fi, Fj in zip(self.kfacs, other.kfacs)]
        return B200Kron(kfacs)

    __radd__ = __add__

    def __iadd__(self, other):
        if not isinstance(other, Kron):
            raise ValueError("Can only add Kron to Kron.")
        if self._same_layout(other):
            self._flat.add_(other._flat)
        else:
            for Fi, Fj in zip(self.kfacs, other.kfacs):
                for Hi, Hj in zip(Fi, Fj):
                    Hi.add_(Hj)
        return self

    # -- reference Kron.__mul__ (utils/matrix.py:100-118): scalar ** (1/len(F)) on every factor -
    def __mul__(self, scalar):
        if not _is_scalar(scalar):
            raise ValueError("Input not valid python or torch scalar.")
        # tensor scalars stay tensors (utils/matrix.py:116-118 uses ``pow(scalar, 1/len(F))``): the autograd graph to a
        # differentiable ``sigma_noise`` / temperature survives and no host sync is forced
        s = scalar if torch.is_tensor(scalar) else float(scalar)
        return B200Kron([[_root(s, len(F)) * Hi for Hi in F] for F in self.kfacs])

    __rmul__ = __mul__

    def __len__(self):
        return len(self.kfacs)

    # -- reference Kron.decompose (utils/matrix.py:123-150) + symeig (utils/utils.py:193-228) ---
    def decompose(self, damping: bool = False) -> "B200KronDecomposed":
        mats = [(i, j, H) for i, F in enumerate(self.kfacs) for j, H in enumerate(F)]
        if mats and mats[0][2].is_cuda:
            wait_prewarm(mats[0][2].device)
        eigvecs = [[None] * len(F) for F in self.kfacs]
        eigvals = [[None] * len(F) for F in self.kfacs]
        by_size: dict[int, list] = {}
        for i, j, H in mats:
            if H.ndim == 1:  # diagonal factor (utils/matrix.py:141-145)
                eigvals[i][j] = H
                eigvecs[i][j] = torch.eye(len(H), dtype=H.dtype, device=H.device)
            else:
                by_size.setdefault(int(H.shape[0]), []).append((i, j, H))
        large = []
        for n, group in by_size.items():
            dtype = group[0][2].dtype
            if n <= K.EIGH_MAX_N and group[0][2].is_cuda and EIGH_FP64_MAX_N < n:
                stack = torch.stack([_as_f32(H) for _, _, H in group])
                ev, Q = K.eigh_jacobi(stack)
                for b, (i, j, _) in enumerate(group):
                    eigvals[i][j], eigvecs[i][j] = ev[b].to(dtype), Q[b].to(dtype)
            else:
                large.extend(group)
        if large:
            _symeig_concurrent(large, eigvals, eigvecs)
        return B200KronDecomposed(eigvecs, eigvals, damping=damping)

    # -- cold-path helpers (utils/matrix.py:222-275); plain device tensor algebra --------------
    def diag(self) -> torch.Tensor:
        out = []
        for F in self.kfacs:
            d0 = F[0].diagonal() if F[0].ndim > 1 else F[0]
            if len(F) == 1:
                out.append(d0)
            else:
                d1 = F[1].diagonal() if F[1].ndim > 1 else F[1]
                out.append(torch.outer(d0, d1).reshape(-1))
        return torch.cat(out)

    def to_matrix(self) -> torch.Tensor:
        blocks = []
        for F in self.kfacs:
            F0 = F[0] if F[0].ndim > 1 else F[0].diag()
            if len(F) == 1:
                blocks.append(F0)
            else:
                F1 = F[1] if F[1].ndim > 1 else F[1].diag()
                blocks.append(torch.kron(F0, F1))
        return torch.block_diag(*blocks)

    def logdet(self) -> torch.Tensor:
        def ld(H):   # diagonal factors are stored as vectors (utils/matrix.py:230-238)
            return torch.logdet(H) if H.ndim > 1 else H.log().sum()

        total = 0
        for F in self.kfacs:
            if len(F) == 1:
                total = total + ld(F[0])
            else:
                total = total + F[1].shape[0] * ld(F[0]) + F[0].shape[0] * ld(F[1])
        return total


def adopt(la):
    """Re-wrap the Kronecker factors of a fitted / loaded ``KronLaplace`` as ``B200Kron`` and decompose them with this
    package's kernels.  Needed after ``la.load_state_dict(...)`` only: the reference rebuilds ``H_facs`` as a plain
    ``Kron`` there (baselaplace.py:1873-1879), so the posterior would run on its torch code path.  Returns ``la``."""
    facs = getattr(la, "H_facs", None)
    if facs is not None and isinstance(facs, Kron) and not isinstance(facs, B200Kron):
        la.H_facs = B200Kron.from_kfacs(facs.kfacs)
        la.H = la.H_facs.decompose(damping=getattr(la, "damping", False))
    return la


N_EIGH_STREAMS = 1


def _symeig_concurrent(items, eigvals, eigvecs):
    """Large factors: one cuSOLVER ``syevd`` each, largest first.  Measured on B200 (ResNet-18's 29 large factors):
    585 ms issued back to back on the current stream vs 640-770 ms spread over 4 streams -- and the side streams' private
    allocator pools made the time erratic (0.7-4 s depending on what the caching allocator held) -- so the default is
    the serial order; ``N_EIGH_STREAMS > 1`` re-enables the overlap.

    Factors with structurally dead coordinates are compacted first (``_symeig_compact``): the input factor of a 3x3
    convolution on a 1x1 feature map has 8 of its 9 kernel positions looking at padding only, i.e. a 4608 x 4608 matrix
    whose non-zero part is 512 x 512 -- the dense ``eigh`` spends 77 ms on it, the compact one 7 ms."""
    items = sorted(items, key=lambda t: -t[2].shape[0])
    live = None
    if COMPACT_DEAD_COORDINATES:
        # one host read for all factors: how many coordinates carry any mass (PSD: zero diagonal <=> zero row/column)
        live = torch.stack([(H.diagonal() != 0).sum() for _, _, H in items]).tolist()
    if BATCHED_MID_SIZES and items[0][2].is_cuda and items[0][2].dtype == torch.float32:
        items, live = _symeig_mid_batched(items, live, eigvals, eigvecs)
        if not items:
            return
    serial = not items[0][2].is_cuda or len(items) == 1 or N_EIGH_STREAMS <= 1

    def one(k, H):
        if live is not None and live[k] <= COMPACT_MAX_LIVE_FRACTION * H.shape[0]:
            try:
                return _symeig_compact(H)
            except RuntimeError as e:   # same result through the dense route; say so instead of hiding it
                import warnings

                warnings.warn(f"laplace_b200: compact eigendecomposition failed ({e}); using the dense one")
        return symeig_large(H)

    if N_EIGH_THREADS > 1 and items[0][2].is_cuda and len(items) > 1:
        _symeig_threaded(items, one, eigvals, eigvecs)
        return
    if serial:
        for k, (i, j, H) in enumerate(items):
            eigvals[i][j], eigvecs[i][j] = one(k, H)
        return
    cur = torch.cuda.current_stream()
    streams = [torch.cuda.Stream() for _ in range(min(N_EIGH_STREAMS, len(items)))]
    for s_ in streams:
        s_.wait_stream(cur)
    for k, (i, j, H) in enumerate(items):
        with torch.cuda.stream(streams[k % len(streams)]):
            eigvals[i][j], eigvecs[i][j] = one(k, H)
            H.record_stream(streams[k % len(streams)])
    for s_ in streams:
        cur.wait_stream(s_)


# Factors whose live block has at most this many rows go through ONE batched library call per size class (256 / 513 rows,
# bordered like ``_eigh_padded``) instead of one call each: 17 factors of <= 513 rows 96 -> 30 ms, 6 of 256 rows 32 -> 3.8 ms.
BATCHED_MID_SIZES = True
BATCHED_MID_CLASSES = (256, 513)


def _symeig_mid_batched(items, live, eigvals, eigvecs):
    """Handles every factor whose live block fits ``BATCHED_MID_CLASSES`` with ``cusolverDnXsyevBatched``; returns the
    remaining ``(items, live)`` for the one-by-one route.  Compaction (dead coordinates dropped) and bordering (negative
    diagonal block that decouples exactly and sorts first) are the same as in ``_symeig_compact`` / ``_eigh_padded``."""
    from . import _cusolver

    if not _cusolver.available():
        return items, live
    groups: dict[int, list] = {}
    rest, rest_live = [], []
    for k, it in enumerate(items):
        n = it[2].shape[0]
        lv = live[k] if (live is not None and live[k] <= COMPACT_MAX_LIVE_FRACTION * n) else n
        cls = next((c for c in BATCHED_MID_CLASSES if lv <= c), None)
        if cls is None or lv == 0:
            rest.append(it)
            rest_live.append(live[k] if live is not None else None)
        else:
            groups.setdefault(cls, []).append((it, lv))
    for cls, grp in groups.items():
        if len(grp) < 2:       # a single factor gains nothing from the batched entry point
            for it, lv in grp:
                rest.append(it)
                rest_live.append(lv if live is not None else None)
            continue
        dev = grp[0][0][2].device
        batch = torch.zeros(len(grp), cls, cls, device=dev, dtype=torch.float32)
        meta = []
        for b, ((i, j, H), lv) in enumerate(grp):
            n = H.shape[0]
            if lv < n:
                alive = H.diagonal() != 0
                idx, dead = torch.nonzero(alive).squeeze(1), torch.nonzero(~alive).squeeze(1)
                Hs = H.index_select(0, idx).index_select(1, idx)
            else:
                idx = dead = None
                Hs = H
            k_ = Hs.shape[0]
            batch[b, :k_, :k_] = torch.triu(Hs) + torch.triu(Hs, 1).t()          # UPLO = 'U' (utils/utils.py:207)
            if k_ < cls:
                border = -(Hs.diagonal().abs().max() + 1.0)
                batch[b, k_:, k_:] = torch.diag_embed(border.expand(cls - k_))
            meta.append((i, j, n, k_, idx, dead))
        try:
            W, Q, info = _cusolver.syev_batched(batch)
        except RuntimeError as e:
            import warnings

            warnings.warn(f"laplace_b200: batched eigensolver unavailable ({e}); using one call per factor")
            for it, lv in grp:
                rest.append(it)
                rest_live.append(lv if live is not None else None)
            continue
        bad = info.ne(0).tolist()          # one host read per size class (the reference's symeig retries on failure too)
        for b, (i, j, n, k_, idx, dead) in enumerate(meta):
            if bad[b]:
                rest.append(grp[b][0])
                rest_live.append(grp[b][1] if live is not None else None)
                continue
            p = cls - k_
            Ls = torch.nan_to_num(W[b, p:].clamp(min=0.0))
            Ws = torch.nan_to_num(Q[b, :k_, p:])
            if idx is None:
                eigvals[i][j], eigvecs[i][j] = Ls, Ws.contiguous()
            else:
                L = torch.cat([torch.zeros(n - k_, dtype=Ls.dtype, device=dev), Ls])
                Wf = torch.zeros(n, n, dtype=Ws.dtype, device=dev)
                Wf[dead, torch.arange(n - k_, device=dev)] = 1
                Wf[idx.unsqueeze(1), (n - k_) + torch.arange(k_, device=dev).unsqueeze(0)] = Ws
                eigvals[i][j], eigvecs[i][j] = L, Wf
    if live is None:
        rest_live = None
    order = sorted(range(len(rest)), key=lambda q: -rest[q][2].shape[0])
    return [rest[q] for q in order], (None if rest_live is None else [rest_live[q] for q in order])


# The library eigensolver synchronises with the host between its phases, so factors issued from ONE thread run strictly
# one after another even on separate streams, each keeping a fraction of the SMs busy.  Worker threads (the GIL is
# released inside ``torch.linalg.eigh``) with one stream each let the phases of different factors interleave on the device.
N_EIGH_THREADS = int(__import__("os").environ.get("LPB_EIGH_THREADS", "4"))   # measured: 315 (1) / 236 (2) / 209 (4) / 209 ms (8)
_EIGH_POOL = {}


def _symeig_threaded(items, one, eigvals, eigvecs):
    import queue
    import threading

    dev = items[0][2].device
    cur = torch.cuda.current_stream(dev)
    nthreads = min(N_EIGH_THREADS, len(items))
    streams = _EIGH_POOL.setdefault((dev, nthreads), [torch.cuda.Stream(dev) for _ in range(nthreads)])
    q = queue.SimpleQueue()
    for k, it in enumerate(items):          # largest first: `items` is sorted by size
        q.put((k, it))
    sizes = [int(it[2].shape[0]) for it in items]
    dt = items[0][2].dtype
    Qflat = torch.empty(sum(n * n for n in sizes), device=dev, dtype=dt)
    Lflat = torch.empty(sum(sizes), device=dev, dtype=dt)
    out, qo, lo = [], 0, 0
    for n in sizes:
        out.append((Lflat[lo:lo + n], Qflat[qo:qo + n * n].view(n, n)))
        qo, lo = qo + n * n, lo + n
    errors = []

    def work(stream):
        try:
            torch.cuda.set_device(dev)
            stream.wait_stream(cur)
            with torch.cuda.stream(stream):
                while True:
                    try:
                        k, (i, j, H) = q.get_nowait()
                    except queue.Empty:
                        break
                    L, W = one(k, H)
                    # results land in buffers the CALLER's stream allocated (``out``): the worker's own pool then only
                    # ever holds one factor's temporaries, so repeated decompositions do not grow it with synchronising
                    # cudaMallocs (429 vs 220 ms for ResNet-18 when the previous result was still alive)
                    Lo, Wo = out[k]
                    Lo.copy_(L), Wo.copy_(W)
                    H.record_stream(stream)
                    eigvals[i][j], eigvecs[i][j] = Lo, Wo
                    del L, W
        except BaseException as e:  # noqa: BLE001 -- re-raised in the caller's thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(s_,), daemon=True) for s_ in streams]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for s_ in streams:
        cur.wait_stream(s_)
    if errors:
        raise errors[0]


_PREWARM = {}


def prewarm_eigensolver(device) -> None:
    """Load and initialise the library eigensolvers on a background thread (own stream) so that their one-time start-up
    -- 0.3 .. 19 s on a fresh process: the cuSOLVER image is paged in and its handles / workspaces are created on first
    use -- overlaps the data pass of ``fit()`` instead of landing inside the first ``decompose()``.  Called by the backend
    on its first ``kron()``; idempotent; failures are left to the real call."""
    dev = torch.device(device)
    if dev.type != "cuda" or dev in _PREWARM:
        return
    import threading

    def work():
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(torch.cuda.Stream(dev)):
                for n in (8, 300, 600):
                    A = torch.eye(n, device=dev, dtype=torch.float32) + 0.01
                    torch.linalg.eigh(A, UPLO="U")
                if BATCHED_MID_SIZES:
                    from . import _cusolver

                    if _cusolver.available():
                        _cusolver.syev_batched((torch.eye(256, device=dev, dtype=torch.float32) + 0.01).expand(2, 256, 256).contiguous())
                torch.cuda.current_stream(dev).synchronize()
        except Exception:  # noqa: BLE001 -- a warm-up must never fail a fit
            pass

    t = threading.Thread(target=work, daemon=True, name="lpb-eigh-prewarm")
    _PREWARM[dev] = t
    t.start()


def wait_prewarm(device) -> None:
    t = _PREWARM.get(torch.device(device))
    if t is not None and t.is_alive():
        t.join()


def live_sizes(mats):
    """Number of coordinates of each PSD factor that carry any mass (zero diagonal <=> zero row / column) -- ONE host
    read for all factors.  ``decompose`` works on the live block only (``_symeig_compact``), so this is the size that
    determines a factor's cost."""
    if not COMPACT_DEAD_COORDINATES or not mats:
        return [int(H.shape[0]) for H in mats]
    live = torch.stack([(H.diagonal() != 0).sum() if H.ndim == 2 else torch.as_tensor(H.numel(), device=H.device)
                        for H in mats]).tolist()
    return [int(l) if l <= COMPACT_MAX_LIVE_FRACTION * H.shape[0] else int(H.shape[0]) for l, H in zip(live, mats)]


# measured on B200 (profiles/r02_eigh.md): milliseconds of one decomposition by (live) size, fp32
_EIGH_MS = ((1, 0.2), (64, 0.3), (128, 1.2), (129, 5.6), (513, 5.6), (576, 6.2), (1024, 11.3), (1152, 13.3), (2304, 30.5), (4608, 85.5))


def eigh_cost_ms(n: int) -> float:
    for (n0, t0), (n1, t1) in zip(_EIGH_MS, _EIGH_MS[1:]):
        if n <= n1:
            return t0 + (t1 - t0) * (max(n, n0) - n0) / max(1, n1 - n0)
    return _EIGH_MS[-1][1] * (n / _EIGH_MS[-1][0]) ** 3


COMPACT_DEAD_COORDINATES = True
COMPACT_MAX_LIVE_FRACTION = 0.75


def _symeig_compact(H: torch.Tensor):
    """Eigendecomposition of a PSD matrix whose rows / columns outside ``idx = {i : H_ii != 0}`` vanish identically:
    ``eigh`` of the live block, embedded.  Dead coordinates get eigenvalue 0 and unit eigenvectors (any orthonormal
    basis of the null space is an equally valid output of the reference's dense ``eigh``); eigenvalues stay ascending
    (the dead zeros first, then the clamped, non-negative live spectrum)."""
    n = H.shape[0]
    alive = H.diagonal() != 0
    idx = torch.nonzero(alive).squeeze(1)
    dead = torch.nonzero(~alive).squeeze(1)
    k = idx.numel()
    if k == 0:
        return torch.zeros(n, dtype=H.dtype, device=H.device), torch.eye(n, dtype=H.dtype, device=H.device)
    Hs = H.index_select(0, idx).index_select(1, idx)
    Ls, Ws = symeig_large(Hs)
    L = torch.cat([torch.zeros(n - k, dtype=Ls.dtype, device=H.device), Ls])
    W = torch.zeros(n, n, dtype=Ws.dtype, device=H.device)
    W[dead, torch.arange(n - k, device=H.device)] = 1
    W[idx.unsqueeze(1), (n - k) + torch.arange(k, device=H.device).unsqueeze(0)] = Ws
    return L, W


# torch.linalg.eigh routes fp32 CUDA matrices of 32..512 rows to cuSOLVER's Jacobi solver (syevj) and everything larger to
# the divide-and-conquer one (syevd); on B200 syevd is ~2.5x faster just above the threshold (576: 6.7 ms) than syevj
# just below it (512: 17 ms).  Matrices in PAD_EIGH_RANGE are therefore bordered with a negative diagonal block up to 513
# rows: the border decouples exactly, sorts first, and is dropped.
PAD_EIGH_RANGE = (129, 512)
PAD_EIGH_TO = 513
PAD_EIGH = True       # measured on B200 (profiles/r02_eigh.md): 256: 5.3 (syevj) vs 5.6 ms padded, 384: 9.8 vs 5.6,
                      # 512: 14.7 vs 5.6; the whole ResNet-18 decomposition 410 -> 315 ms
_PAD_ON_CPU = False   # tests flip this to exercise the bordering logic without a GPU


def _eigh_padded(H: torch.Tensor):
    n = H.shape[0]
    p = PAD_EIGH_TO - n
    Hp = torch.zeros(PAD_EIGH_TO, PAD_EIGH_TO, dtype=H.dtype, device=H.device)
    Hp[:n, :n] = H
    # strictly below the spectrum of a PSD H by far more than any rounding error of its eigenvalues
    border = -(H.diagonal().abs().max() + 1.0)
    Hp[n:, n:] = torch.diag_embed(border.expand(p))
    L, W = torch.linalg.eigh(Hp, UPLO="U")
    return L[p:], W[:n, p:]


# Factors of at most this many rows are decomposed in fp64 (library eigh on an upcast copy, results rounded back to
# the factor dtype).  0 = off.  Experiment knob: how much of the predictive error is the fp32 eigendecomposition.
EIGH_FP64_MAX_N = int(__import__("os").environ.get("LPB_EIGH_FP64_MAX_N", "0"))


def symeig_large(H: torch.Tensor):
    if H.shape[0] <= EIGH_FP64_MAX_N and H.dtype != torch.float64:
        L, W = torch.linalg.eigh(H.double(), UPLO="U")
        return torch.nan_to_num(L.clamp(min=0.0)).to(H.dtype), torch.nan_to_num(W).to(H.dtype)
    return _symeig_large(H)


def _symeig_large(H: torch.Tensor):
    """Factors beyond the Jacobi kernel's limit: cuSOLVER ``syevd`` via ``torch.linalg.eigh`` (LIBRARY call,
    declared as such in DESIGN.md) with the reference's post-processing (utils/utils.py:207-228): jitter retry,
    clamp at 0, NaN -> 0; raises ``LinAlgError`` instead of the reference's ``exit()``."""
    try:
        if ((PAD_EIGH and H.is_cuda) or _PAD_ON_CPU) and H.dtype == torch.float32 \
                and PAD_EIGH_RANGE[0] <= H.shape[0] <= PAD_EIGH_RANGE[1]:
            L, W = _eigh_padded(H)
        else:
            L, W = torch.linalg.eigh(H, UPLO="U")
    except RuntimeError:
        eye = torch.eye(H.shape[0], device=H.device, dtype=H.dtype)
        try:
            L, W = torch.linalg.eigh(H + eye, UPLO="U")
            L = L - 1.0
        except RuntimeError as e:  # pragma: no cover
            raise torch.linalg.LinAlgError(f"symmetric eigendecomposition failed: {e}") from e
    return torch.nan_to_num(L.clamp(min=0.0)), torch.nan_to_num(W)


# =========================================================================================
class JacobianFactors:
    """Per-parameter-block structure of a batch of Jacobians, attached to the dense tensor returned by
    ``B200GGN.jacobians`` as ``Js._lpb_factors``.  ``blocks[i]`` describes parameter block ``i``:

    * ``("outer", g, a)``: ``J[n,c] = g[c,n,:] (x) a[n,:]``  (weight of a layer without weight sharing)
    * ``("vec", g)``:      ``J[n,c] = g[c,n,:]``             (bias of such a layer)
    * ``("dense", J)``:    ``J [Nn, C, p]`` fp32 contiguous  (anything else)
    """

    def __init__(self, blocks, n_batch: int, n_out: int, sizes):
        self.blocks, self.n_batch, self.n_out, self.sizes = blocks, n_batch, n_out, sizes
        self._proj = (None, {})
        # rotated rows of weight-sharing layers double the memory of a batch: cached only when a sweep over prior precisions
        # will reuse them (``backend.cached_jacobians()``, ``laplace_b200.tuning``)
        self.keep_projections = False

    def projections(self, basis_key) -> dict:
        """Per-block eigenbasis projections of these Jacobians for ONE decomposition (identified by the rotation cache all
        its scaled / shifted copies share).  A different decomposition starts an empty cache."""
        if self._proj[0] is not basis_key:
            self._proj = (basis_key, {})
        return self._proj[1]


class B200KronDecomposed(KronDecomposed):
    """Eigendecomposed Kronecker factors (+ per-block ``deltas``); reference utils/matrix.py:282-560."""

    def __init__(self, eigenvectors, eigenvalues, deltas=None, damping: bool = False):
        ref = eigenvectors[0][0]
        if deltas is None:
            deltas = torch.zeros(len(eigenvalues), device=ref.device, dtype=ref.dtype)
        else:
            self._validate(deltas, len(eigenvalues))
        self.eigenvectors = eigenvectors
        self.eigenvalues = eigenvalues
        self.deltas = deltas
        self.damping = damping
        self._cache = {}

    @staticmethod
    def _validate(deltas, n):
        if not isinstance(deltas, torch.Tensor):
            raise ValueError("Can only add torch.Tensor to KronDecomposed.")
        if not (deltas.ndim == 0 or (deltas.ndim == 1 and len(deltas) in (1, n))):
            raise ValueError("Invalid shape of delta added to KronDecomposed.")

    def detach(self):
        self.deltas = self.deltas.detach()
        return self

    def __len__(self):
        return len(self.eigenvalues)

    # NB: like the reference (utils/matrix.py:355, :376) ``+`` and ``*`` rebuild the object with the
    # default ``damping=False`` -- reproduced on purpose so results stay identical.
    def __add__(self, deltas):
        self._validate(deltas, len(self))
        out = B200KronDecomposed(self.eigenvectors, self.eigenvalues, self.deltas + deltas)
        out._cache = self._cache  # rotation operands do not depend on deltas / scaling
        return out

    __radd__ = __add__

    def __mul__(self, scalar):
        if not _is_scalar(scalar):
            raise ValueError("Invalid argument, can only multiply Kron with scalar.")
        # utils/matrix.py:372-376: ``pow(scalar, 1/len(ls)) * eigval`` -- a tensor scalar (``_H_factor`` with a
        # differentiable ``sigma_noise``) keeps its graph, so d logdet / d sigma_noise flows like in the reference
        s = scalar if torch.is_tensor(scalar) else float(scalar)
        ev = [[_root(s, len(ls)) * l for l in ls] for ls in self.eigenvalues]
        out = B200KronDecomposed(self.eigenvectors, ev, self.deltas)
        out._cache = self._cache
        return out

    __rmul__ = __mul__

    # -- spectra ---------------------------------------------------------------------------
    def _delta_list(self):
        d = self.deltas
        if d.ndim == 0 or d.numel() == 1:
            return [d.reshape(())] * len(self)
        return list(d)

    def _spectrum(self, ls, delta):
        if len(ls) == 1:
            return ls[0] + delta
        if self.damping:
            sd = torch.sqrt(delta)
            return torch.outer(ls[0] + sd, ls[1] + sd)
        return torch.outer(ls[0], ls[1]) + delta

    def logdet(self) -> torch.Tensor:
        """``log det (Kron + deltas)`` (utils/matrix.py:381-404); plain tensor ops so that it stays
        differentiable w.r.t. ``deltas`` (marginal-likelihood optimisation, baselaplace.py:363-561)."""
        total = 0
        for ls, delta in zip(self.eigenvalues, self._delta_list()):
            total = total + torch.log(self._spectrum(ls, delta)).sum()
        return total

    # -- fp32 rotation operands (cached; shared by every scaled / shifted copy) ----------------
    def _Q32(self, i, j, transposed: bool):
        key = (i, j, transposed)
        if key not in self._cache:
            Q = _as_f32(self.eigenvectors[i][j])
            self._cache[key] = (Q.t() if transposed else Q).contiguous()
        return self._cache[key]

    @staticmethod
    def _gemm(A2d: torch.Tensor, Bnk: torch.Tensor) -> torch.Tensor:
        """``A2d [M,K] @ Bnk[N,K]^T`` through the fp32 kernel (operands already K-major)."""
        A2d = A2d.contiguous()
        out = torch.empty(A2d.shape[0], Bnk.shape[0], device=A2d.device, dtype=torch.float32)
        pa = K.Packed(A2d, None, K.F32, A2d.shape[0], A2d.shape[1])
        pb = K.Packed(Bnk, None, K.F32, Bnk.shape[0], Bnk.shape[1])
        return K.gemm_nt(pa, pb, out, 1.0, accumulate=False)

    def _rotate_in(self, i, Wp: torch.Tensor) -> torch.Tensor:
        """``Wp [R, p1, p2]`` -> ``Zt [R, p2, p1]`` with ``Zt[r] = (Q1^T Wp[r] Q2)^T``."""
        R, p1, p2 = Wp.shape
        Y = self._gemm(Wp.reshape(R * p1, p2), self._Q32(i, 1, True)).view(R, p1, p2)      # Wp Q2
        Yt = Y.transpose(1, 2).contiguous().view(R * p2, p1)
        return self._gemm(Yt, self._Q32(i, 0, True)).view(R, p2, p1)                      # (Q1^T Y)^T

    def _rotate_out(self, i, Zt: torch.Tensor) -> torch.Tensor:
        """inverse of ``_rotate_in``: ``Zt [R, p2, p1]`` -> ``Q1 Z Q2^T`` as ``[R, p1, p2]``."""
        R, p2, p1 = Zt.shape
        Y = self._gemm(Zt.reshape(R * p2, p1), self._Q32(i, 0, False)).view(R, p2, p1)     # (Q1 Z)^T
        Yt = Y.transpose(1, 2).contiguous().view(R * p1, p2)
        return self._gemm(Yt, self._Q32(i, 1, False)).view(R, p1, p2)

    # -- reference KronDecomposed._bmm (utils/matrix.py:406-456) --------------------------------
    def _bmm(self, W: torch.Tensor, exponent: float = -1) -> torch.Tensor:
        assert W.ndim == 3
        B, Kk, P = W.shape
        dtype = W.dtype
        Wf = _as_f32(W).reshape(B * Kk, P)
        out, cur = [], 0
        for i, (ls, delta) in enumerate(zip(self.eigenvalues, self._delta_list())):
            spec = torch.pow(_as_f32(self._spectrum(ls, delta)), exponent)
            if len(ls) == 1:
                p = ls[0].numel()
                Z = self._gemm(Wf[:, cur:cur + p], self._Q32(i, 0, True)) * spec            # (Q^T w)^T * spec
                out.append(self._gemm(Z, self._Q32(i, 0, False)))
            else:
                p1, p2 = ls[0].numel(), ls[1].numel()
                p = p1 * p2
                Zt = self._rotate_in(i, Wf[:, cur:cur + p].reshape(-1, p1, p2)) * spec.t()
                out.append(self._rotate_out(i, Zt).reshape(-1, p))
            cur += p
        return torch.cat(out, dim=1).reshape(B, Kk, P).to(dtype)

    def bmm(self, W: torch.Tensor, exponent: float = -1) -> torch.Tensor:
        if W.ndim == 1:
            return self._bmm(W.unsqueeze(0).unsqueeze(0), exponent).squeeze()
        if W.ndim == 2:
            return self._bmm(W.unsqueeze(1), exponent).squeeze()
        if W.ndim == 3:
            return self._bmm(W, exponent)
        raise ValueError("Invalid shape for W")

    # -- reference KronDecomposed.inv_square_form (utils/matrix.py:458-461) ----------------------
    def inv_square_form(self, W: torch.Tensor) -> torch.Tensor:
        fac = getattr(W, "_lpb_factors", None)
        if fac is not None and len(fac.blocks) == len(self) and not W.requires_grad:
            return self._structured_isf(fac).to(W.dtype)
        Wf = _as_f32(W).contiguous()
        SW = _as_f32(self._bmm(Wf, exponent=-1)).contiguous()
        out = torch.empty(W.shape[0], W.shape[1], W.shape[1], device=W.device, dtype=torch.float32)
        K.batched_pair_dot(Wf, SW, None, out)
        return out.to(W.dtype)

    def _structured_isf(self, fac: JacobianFactors) -> torch.Tensor:
        """``f_var[n,c,k] = sum_blocks sum_ij Zc[i,j] Zk[i,j] / spec[i,j]`` with the per-layer rank structure:
        for ``J = g (x) a``: ``sum_i gt_c[i] gt_k[i] m[i]``, ``gt = Q1^T g``, ``m = (Q2^T a)^2 @ (1/spec)^T``.

        The eigenbasis projections (``gt``, ``(Q2^T a)^2``, rotated dense blocks) depend on the Jacobian and on ``Q`` only,
        not on ``deltas`` or the eigenvalue scaling: they are cached on ``fac`` (keyed by this decomposition's shared
        rotation cache), so a sweep over prior precisions on the same validation batch -- ``optimize_prior_precision``'s
        grid search, baselaplace.py:516-561 -- pays for them once."""
        Nn, C = fac.n_batch, fac.n_out
        dev = self.eigenvectors[0][0].device
        out = torch.zeros(Nn, C, C, device=dev, dtype=torch.float32)
        proj = fac.projections(self._cache)
        for i, (blk, ls, delta) in enumerate(zip(fac.blocks, self.eigenvalues, self._delta_list())):
            inv_spec = torch.reciprocal(_as_f32(self._spectrum(ls, delta))).contiguous()
            kind = blk[0]
            if kind == "outer" and len(ls) == 2:
                if i not in proj:
                    g, a = blk[1], blk[2]                                # [C, Nn, d_out], [Nn, d_in]
                    d_out = g.shape[2]
                    gt = self._gemm(g.reshape(C * Nn, d_out), self._Q32(i, 0, True)).view(C, Nn, d_out)
                    at = self._gemm(a, self._Q32(i, 1, True))
                    proj[i] = (gt.permute(1, 0, 2), at * at)             # [Nn, C, d_out] view, [Nn, d_in]
                gtn, at2 = proj[i]
                m = self._gemm(at2, inv_spec)                            # [Nn, d_out]
                K.batched_pair_dot(gtn, gtn, m, out, accumulate=True)
            elif kind == "conv" and len(ls) == 2 and C <= 12:
                if i in proj:
                    Gt, At, T = proj[i]
                else:
                    Grows, Arows, T = blk[1], blk[2], blk[3]            # [(c,n,t), d_out], [(n,t), d_in]
                    Gt = self._gemm(self._Q32(i, 0, True), Grows)        # Q1^T G^T -> [d_out, C*Nn*T] (K-major)
                    At = self._gemm(self._Q32(i, 1, True), Arows)        # Q2^T A^T -> [d_in, Nn*T]
                    if fac.keep_projections:   # as large as the rows themselves: kept only for sweeps over deltas
                        proj[i] = (Gt, At, T)
                K.kron_conv_quadform(Gt, At, T, Nn, C, _as_f32(ls[0]).contiguous(), _as_f32(ls[1]).contiguous(),
                                     float(delta), self.damping, out)
            elif kind == "vec" and len(ls) == 1:
                if i not in proj:
                    g = blk[1]
                    d_out = g.shape[2]
                    gt = self._gemm(g.reshape(C * Nn, d_out), self._Q32(i, 0, True)).view(C, Nn, d_out)
                    proj[i] = (gt.permute(1, 0, 2),)
                gtn = proj[i][0]
                K.batched_pair_dot(gtn, gtn, inv_spec, out, accumulate=True)
            else:
                if i not in proj:
                    J = blk[1] if kind == "dense" else materialize_block(blk, Nn, C)
                    if J.dim() == 2:
                        J = J.unsqueeze(2)
                    p = J.shape[2]
                    if len(ls) == 1:
                        proj[i] = (self._gemm(J.reshape(Nn * C, p), self._Q32(i, 0, True)).view(Nn, C, p),)
                    else:
                        p1, p2 = ls[0].numel(), ls[1].numel()
                        proj[i] = (self._rotate_in(i, J.reshape(Nn * C, p1, p2)).reshape(Nn, C, p),)
                Z = proj[i][0]
                w = inv_spec if len(ls) == 1 else inv_spec.t().contiguous().reshape(-1)
                K.batched_pair_dot(Z, Z, w, out, accumulate=True)
        return out

    # -- cold-path helpers (utils/matrix.py:490-556) ---------------------------------------------
    def diag(self, exponent: float = 1) -> torch.Tensor:
        out = []
        for Qs, ls, delta in zip(self.eigenvectors, self.eigenvalues, self._delta_list()):
            spec = torch.pow(self._spectrum(ls, delta), exponent)
            if len(ls) == 1:
                out.append(((Qs[0] * spec) * Qs[0]).sum(1))
            else:
                out.append(((Qs[0] ** 2) @ spec @ (Qs[1] ** 2).T).reshape(-1))
        return torch.cat(out)

    def to_matrix(self, exponent: float = 1) -> torch.Tensor:
        blocks = []
        for Qs, ls, delta in zip(self.eigenvectors, self.eigenvalues, self._delta_list()):
            spec = torch.pow(self._spectrum(ls, delta), exponent).reshape(-1)
            Q = Qs[0] if len(ls) == 1 else torch.kron(Qs[0].contiguous(), Qs[1].contiguous())   # kron() views its inputs
            blocks.append((Q * spec) @ Q.T)
        return torch.block_diag(*blocks)


def materialize_block(blk, Nn: int, C: int) -> torch.Tensor:
    """Dense ``[Nn, C, p]`` rows of an ``outer`` / ``vec`` block (only reached when a block's factor structure
    and the decomposed block's structure disagree)."""
    if blk[0] == "dense":
        return blk[1]
    if blk[0] == "vec":
        return blk[1].permute(1, 0, 2).contiguous()
    if blk[0] == "conv":
        # J_{n,c} = G_{n,c}^T A_n through the per-sample contraction kernel (needs K-major operands)
        Grows, Arows, T = blk[1], blk[2], blk[3]
        d_out, d_in = Grows.shape[1], Arows.shape[1]
        Gk = K.Packed(Grows.t().contiguous(), None, K.F32, d_out, Grows.shape[0])
        Ak = K.Packed(Arows.t().contiguous(), None, K.F32, d_in, Arows.shape[0])
        J = torch.empty(Nn, C, d_out * d_in, device=Grows.device, dtype=torch.float32)
        K.shared_weight_contract(1, Gk, Ak, d_out, d_in, T, Nn, C, J, js_stride_n=C * d_out * d_in, js_stride_c=d_out * d_in)
        return J
    g, a = blk[1], blk[2]
    d_out, d_in = g.shape[2], a.shape[1]
    J = torch.empty(Nn, C, d_out * d_in, device=g.device, dtype=torch.float32)
    K.jac_linear_write(g, a, J, C * d_out * d_in, d_out * d_in, 0, -1)
    return J
