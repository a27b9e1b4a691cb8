"""Input pipeline helper for the (unmodified) reference front end.

``ParametricLaplace.fit`` moves every batch with a blocking ``X.to(device)`` right before the curvature call
(baselaplace.py:969-975): the host-to-device copy (2 ms for 4096 CIFAR-sized images over PCIe) sits on the critical
path of every step.  ``PrefetchLoader`` wraps any loader and yields batches that are ALREADY on the device, copied on a
side stream one batch ahead of the consumer, so the reference's ``.to(device)`` becomes a no-op and the copy of batch
``i + 1`` overlaps the curvature kernels of batch ``i``.  ``len(loader.dataset)`` (the global ``N`` the KFAC scaling
needs, baselaplace.py:964) is passed through.
"""
from __future__ import annotations

from collections.abc import MutableMapping

import torch


class PrefetchLoader:
    def __init__(self, loader, device, depth: int = 1):
        self.loader, self.device = loader, torch.device(device)
        self.dataset = loader.dataset
        self.depth = max(1, int(depth))
        self._stream = None

    def __len__(self):
        return len(self.loader)

    def _move(self, obj):
        if torch.is_tensor(obj):
            return obj.to(self.device, non_blocking=True)
        if isinstance(obj, MutableMapping):
            if hasattr(obj, "to"):              # e.g. a Hugging Face BatchEncoding
                return obj.to(self.device)
            return {k: self._move(v) for k, v in obj.items()}
        if isinstance(obj, (tuple, list)):
            return type(obj)(self._move(v) for v in obj)
        return obj

    @staticmethod
    def _tensors(obj):
        if torch.is_tensor(obj):
            yield obj
        elif isinstance(obj, MutableMapping):
            for v in obj.values():
                yield from PrefetchLoader._tensors(v)
        elif isinstance(obj, (tuple, list)):
            for v in obj:
                yield from PrefetchLoader._tensors(v)

    def __iter__(self):
        if self.device.type != "cuda":
            for batch in self.loader:
                yield self._move(batch)
            return
        main = torch.cuda.current_stream(self.device)
        if self._stream is None:
            self._stream = torch.cuda.Stream(self.device)   # kept: its allocator pool stays warm across epochs
        copy = self._stream
        queue = []

        def fetch(batch):
            with torch.cuda.stream(copy):
                moved = self._move(batch)
            ev = torch.cuda.Event()
            ev.record(copy)
            queue.append((moved, ev))

        it = iter(self.loader)
        try:
            for _ in range(self.depth):
                fetch(next(it))
        except StopIteration:
            pass
        while queue:
            moved, ev = queue.pop(0)
            main.wait_event(ev)
            for t in self._tensors(moved):
                if t.is_cuda:
                    t.record_stream(main)     # allocated on the copy stream, consumed on the compute stream
            try:
                fetch(next(it))
            except StopIteration:
                pass
            yield moved
