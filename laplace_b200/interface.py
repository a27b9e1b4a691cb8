"""Resolution of the plug-in base classes.

When the reference library is importable (a user environment), the B200 backend derives from
the *real* ``laplace.curvature.GGNInterface`` / ``EFInterface`` and returns subclasses of the
real ``laplace.utils.matrix.Kron`` / ``KronDecomposed`` so that ``Laplace(model, ...,
backend=B200GGN)`` works unchanged (reference construction site: baselaplace.py:179-194;
``SubnetLaplace`` ``issubclass`` check: subnetlaplace.py:104).

On a machine without the reference (the GPU box of this build) the same names resolve to small
host-side mirrors that reproduce only the constructor contract of
``CurvatureInterface.__init__`` (curvature/curvature.py:46-86) and the container contract of
``Kron`` / ``KronDecomposed`` (utils/matrix.py:30-31, 304-324) -- all arithmetic lives in the
B200 subclasses either way.
"""
from __future__ import annotations

import torch
from torch import nn

from . import compat

try:  # pragma: no cover - depends on the environment
    if not compat.enable_reference():
        raise ImportError("reference front end not available")
    from laplace.curvature import CurvatureInterface, EFInterface, GGNInterface  # type: ignore
    from laplace.utils.matrix import Kron, KronDecomposed  # type: ignore

    HAVE_REFERENCE = True
except ImportError:
    HAVE_REFERENCE = False

    class CurvatureInterface:  # mirror of curvature/curvature.py:12-86 (constructor contract only)
        def __init__(self, model, likelihood, last_layer=False, subnetwork_indices=None, dict_key_x="input_ids",
                     dict_key_y="labels"):
            if likelihood not in ("regression", "classification"):
                raise AssertionError(f"unsupported likelihood {likelihood!r}")
            self.likelihood = likelihood
            self.model = model
            self.last_layer = last_layer
            self.subnetwork_indices = subnetwork_indices
            self.dict_key_x = dict_key_x
            self.dict_key_y = dict_key_y
            regression = likelihood == "regression"
            self.lossfunc = nn.MSELoss(reduction="sum") if regression else nn.CrossEntropyLoss(reduction="sum")
            self.factor = 0.5 if regression else 1.0
            owner = self._model
            self.params = [p for p in owner.parameters() if p.requires_grad]
            self.params_dict = {k: v for k, v in owner.named_parameters() if v.requires_grad}
            self.buffers_dict = dict(self.model.named_buffers())

        @property
        def _model(self):
            return self.model.last_layer if self.last_layer else self.model

        def jacobians(self, x, enable_backprop=False):
            raise NotImplementedError

        def last_layer_jacobians(self, x, enable_backprop=False):
            raise NotImplementedError

        def gradients(self, x, y):
            raise NotImplementedError

        def full(self, x, y, **kwargs):
            raise NotImplementedError

        def kron(self, x, y, N, **kwargs):
            raise NotImplementedError

        def diag(self, x, y, **kwargs):
            raise NotImplementedError

    class GGNInterface(CurvatureInterface):  # mirror of curvature/curvature.py:294-339
        def __init__(self, model, likelihood, last_layer=False, subnetwork_indices=None, dict_key_x="input_ids",
                     dict_key_y="labels", stochastic=False, num_samples=1):
            self.stochastic = stochastic
            self.num_samples = num_samples
            super().__init__(model, likelihood, last_layer, subnetwork_indices, dict_key_x, dict_key_y)

    class EFInterface(CurvatureInterface):  # mirror of curvature/curvature.py:436-465
        pass

    class Kron:  # container contract of utils/matrix.py:16-31
        def __init__(self, kfacs):
            self.kfacs = kfacs

        def __len__(self):
            return len(self.kfacs)

    class KronDecomposed:  # container contract of utils/matrix.py:282-324
        def __init__(self, eigenvectors, eigenvalues, deltas=None, damping=False):
            self.eigenvectors = eigenvectors
            self.eigenvalues = eigenvalues
            ref = eigenvectors[0][0]
            self.deltas = torch.zeros(len(eigenvalues), device=ref.device, dtype=ref.dtype) if deltas is None else deltas
            self.damping = damping

        def __len__(self):
            return len(self.eigenvalues)


__all__ = ["CurvatureInterface", "GGNInterface", "EFInterface", "Kron", "KronDecomposed", "HAVE_REFERENCE"]
