"""Making the reference front end importable next to the B200 backend.

``laplace-torch`` imports every curvature library it can adapt to at module import time (``laplace/baselaplace.py:11,
18-22``: torchmetrics, asdl, backpack, curvlinops; ``laplace/utils/matrix.py:520``: opt_einsum).  A deployment that uses
``backend=B200GGN`` needs none of them -- the B200 backend replaces exactly those libraries -- but without them
``import laplace`` fails before a backend can be chosen.  ``enable_reference()`` registers inert placeholder modules for
the ones that are missing (nothing in them is ever executed on the B200 path: instantiating a placeholder raises) and
puts an installed copy of the reference on ``sys.path``:

* ``$LPB_REFERENCE_PATH`` if set,
* ``<repo>/baseline/_ref`` (``pip install --no-deps --target baseline/_ref`` of the reference, see
  ``tools/install_reference.sh``),
* ``/root/reference`` (the source tree of the build container).

``laplace_b200.interface`` calls it lazily, so ``from laplace import Laplace; Laplace(model, ..., backend=B200GGN)``
works on a machine that has nothing but PyTorch, the reference package and this backend.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_OPTIONAL = (
    "curvlinops", "curvlinops._base",
    "backpack", "backpack.context", "backpack.extensions",
    "asdl", "asdl.fisher", "asdl.grad_maker", "asdl.gradient", "asdl.hessian", "asdl.matrices",
)


class _Absent:
    """Any attribute of an absent optional dependency: usable as a base class or a name, never instantiable."""

    def __init__(self, *a, **k):
        raise ModuleNotFoundError(f"{type(self).__module__}.{type(self).__name__}: optional dependency of laplace-torch "
                                  "that is not installed (the B200 backend does not need it)")

    def __init_subclass__(cls, **kw):
        pass


class _AbsentModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (_Absent,), {"__module__": self.__name__})
        setattr(self, name, obj)
        return obj


def _have(name: str) -> bool:
    if name in sys.modules:
        return True
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def _torchmetrics_placeholder() -> types.ModuleType:
    """``laplace/utils/metrics.py`` subclasses ``torchmetrics.Metric`` (running NLL for marglik training); a state-holding
    ``nn.Module`` with ``add_state`` / ``reset`` is all it relies on."""
    import torch

    tm = types.ModuleType("torchmetrics")

    class Metric(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self._lpb_defaults = {}

        def add_state(self, name, default, dist_reduce_fx=None):
            self._lpb_defaults[name] = default
            if torch.is_tensor(default):
                self.register_buffer(name, default.clone(), persistent=False)   # follows ``.to(device)`` like the real states
            else:
                setattr(self, name, list(default))

        def reset(self):
            for name, default in self._lpb_defaults.items():
                if torch.is_tensor(default):
                    getattr(self, name).copy_(default)
                else:
                    setattr(self, name, list(default))

        def forward(self, *a, **k):
            self.update(*a, **k)
            return self.compute()

    class MeanSquaredError(Metric):
        def __init__(self, num_outputs: int = 1, **k):
            super().__init__()
            self.add_state("sum_squared_error", torch.zeros(num_outputs))
            self.add_state("total", torch.tensor(0.0))

        def update(self, preds, target):
            self.sum_squared_error += ((preds - target) ** 2).sum(0)
            self.total += target.shape[0]

        def compute(self):
            return self.sum_squared_error / self.total

    tm.Metric, tm.MeanSquaredError = Metric, MeanSquaredError
    return tm


def candidate_paths() -> list[str]:
    out = []
    env = os.environ.get("LPB_REFERENCE_PATH")
    if env:
        out.append(env)
    out += [os.path.join(_REPO, "baseline", "_ref"), "/root/reference"]
    return [p for p in out if os.path.isdir(os.path.join(p, "laplace"))]


def enable_reference(path: str | None = None) -> bool:
    """Returns True when ``import laplace`` works afterwards.  Idempotent; never raises for a missing reference."""
    if os.environ.get("LPB_NO_REFERENCE") == "1":
        return False
    if "laplace" in sys.modules:
        return True
    if not _have("laplace"):
        paths = [path] if path else candidate_paths()
        if not paths:
            return False
        if paths[0] not in sys.path:
            sys.path.insert(0, paths[0])
    if not _have("torchmetrics"):
        sys.modules["torchmetrics"] = _torchmetrics_placeholder()
    if not _have("opt_einsum"):
        import torch

        oe = types.ModuleType("opt_einsum")
        oe.contract = torch.einsum   # its one use: laplace/utils/matrix.py:520
        sys.modules["opt_einsum"] = oe
    missing = {top for top in {n.split(".")[0] for n in _OPTIONAL} if not _have(top)}
    for name in _OPTIONAL:
        if name.split(".")[0] in missing and name not in sys.modules:
            sys.modules[name] = _AbsentModule(name)
    try:
        importlib.import_module("laplace")
    except Exception:  # noqa: BLE001 -- a broken reference install must not break the stand-alone backend
        return False
    return True
