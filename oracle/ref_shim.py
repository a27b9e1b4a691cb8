"""TEST INFRASTRUCTURE ONLY -- import shim that makes the *unmodified* reference
package under ``/root/reference`` importable in the build container.

The reference (``laplace/baselaplace.py:11,18-22``) imports ``torchmetrics``,
``asdl``, ``backpack`` and ``curvlinops`` unconditionally; none of them is
installed here.  We register placeholder modules in ``sys.modules`` so that the
pure-torch half of the reference (``CurvatureInterface``, ``GGNInterface``,
``EFInterface``, ``Kron``, ``KronDecomposed``, ``Full/Kron/Diag(LL)Laplace``)
imports and runs on CPU.  Nothing from the placeholders is ever *executed* by
the code paths we use as the oracle.

``/root/reference`` does not exist on the GPU box: callers must check
``reference_available()`` and skip otherwise.  Only ``tests/`` and the golden
vector generator (``tests/golden/make_golden.py``) may import this file.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LAPLACE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "laplace"))


class _Placeholder:
    """Stands in for any class/constant of an absent third-party package."""

    def __init__(self, *a, **k):
        raise ModuleNotFoundError("placeholder for an absent third-party dependency")

    def __init_subclass__(cls, **kw):
        pass


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (_Placeholder,), {})
        setattr(self, name, obj)
        return obj


def _install_stub(name: str) -> types.ModuleType:
    if name in sys.modules:
        return sys.modules[name]
    mod = _StubModule(name)
    mod.__path__ = []  # behave like a package so that submodules resolve
    sys.modules[name] = mod
    return mod


def _install_torchmetrics() -> None:
    if "torchmetrics" in sys.modules:
        return
    try:
        importlib.import_module("torchmetrics")
        return
    except ModuleNotFoundError:
        pass
    import torch

    tm = types.ModuleType("torchmetrics")

    class Metric(torch.nn.Module):  # minimal surface used by laplace/utils/metrics.py
        def __init__(self, *a, **k):
            super().__init__()
            self._defaults = {}

        def add_state(self, name, default, dist_reduce_fx=None):
            self._defaults[name] = default.clone()
            setattr(self, name, default.clone())

        def reset(self):
            for k, v in self._defaults.items():
                setattr(self, k, v.clone())

    class MeanSquaredError(Metric):
        def __init__(self, num_outputs=1, **k):
            super().__init__()
            self.add_state("sum_squared_error", torch.zeros(num_outputs))
            self.add_state("total", torch.tensor(0.0))

        def update(self, preds, target):
            self.sum_squared_error = self.sum_squared_error + ((preds - target) ** 2).sum(0)
            self.total = self.total + target.shape[0]

        def compute(self):
            return self.sum_squared_error / self.total

    tm.Metric = Metric
    tm.MeanSquaredError = MeanSquaredError
    sys.modules["torchmetrics"] = tm


def install() -> bool:
    """Make ``import laplace`` resolve to the reference tree.  Returns False when
    the reference is not mounted (GPU box)."""
    if not reference_available():
        return False
    _install_torchmetrics()
    for pkg in ("opt_einsum",):
        try:
            importlib.import_module(pkg)
        except ModuleNotFoundError:
            import torch

            m = types.ModuleType(pkg)
            m.contract = torch.einsum  # only use: laplace/utils/matrix.py:520
            sys.modules[pkg] = m
    for name in (
        "curvlinops", "curvlinops._base",
        "backpack", "backpack.context", "backpack.extensions",
        "asdl", "asdl.fisher", "asdl.grad_maker", "asdl.gradient", "asdl.hessian", "asdl.matrices",
    ):
        try:
            importlib.import_module(name)
        except ModuleNotFoundError:
            _install_stub(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    importlib.import_module("laplace")
    return True
