"""Stand-alone host driver for machines without the reference library (e.g. the GPU box of this build).

It mirrors, for the structures on the hot path only, the host flow of the reference --
``ParametricLaplace.fit`` (baselaplace.py:904-987), ``KronLaplace.fit`` post-processing (:1779-1809),
``posterior_precision`` (:1663-1673, :1811-1820, :2097-2100), ``_glm_predictive_distribution``
(:1306-1342; last layer: lllaplace.py:212-237) and the probit link (:662-664) -- so that parity tests
and ``bench.py`` can run ``fit`` + GLM predictive end to end through the same backend calls the
reference would make.  With the reference installed, use ``laplace.Laplace(..., backend=B200GGN)``
directly; this class is not needed.
"""
from __future__ import annotations

import math
import os
from collections.abc import MutableMapping

import torch
from torch import nn
from torch.nn.utils import parameters_to_vector

from . import kernels as K
from .backend import B200GGN
from .interface import Kron
from .matrix import B200Kron


class LastLayerModel(nn.Module):
    """Minimal mirror of ``laplace.utils.feature_extractor.FeatureExtractor`` (feature_extractor.py:97-157):
    exposes ``last_layer`` and ``forward_with_features`` (features = input of the last ``nn.Linear``)."""

    def __init__(self, model: nn.Module, last_layer_name: str | None = None):
        super().__init__()
        self.model = model
        if last_layer_name is None:
            names = [n for n, m in model.named_modules() if isinstance(m, nn.Linear)]
            if not names:
                raise ValueError("Use model with a linear last layer.")
            last_layer_name = names[-1]
        self._last_layer_name = last_layer_name
        self.last_layer = dict(model.named_modules())[last_layer_name]
        if not isinstance(self.last_layer, nn.Linear):
            raise ValueError("Use model with a linear last layer.")
        self._features = None
        self.last_layer.register_forward_hook(self._hook)

    def _hook(self, mod, inp, out):
        self._features = inp[0].detach()

    def forward(self, x):
        return self.model(x)

    def forward_with_features(self, x):
        out = self.model(x)
        return out, self._features


class B200Laplace:
    """``Laplace(model, likelihood, subset_of_weights, hessian_structure)`` for
    ``("all" | "last_layer") x ("kron" | "full" | "diag")`` on top of a B200 backend."""

    def __init__(self, model, likelihood, subset_of_weights="all", hessian_structure="kron", prior_precision=1.0,
                 sigma_noise=1.0, temperature=1.0, backend=B200GGN, backend_kwargs=None, asdl_fisher_kwargs=None,
                 damping=False):
        if likelihood not in ("classification", "regression"):
            raise ValueError(f"Invalid likelihood type {likelihood}")
        if subset_of_weights not in ("all", "last_layer") or hessian_structure not in ("kron", "full", "diag"):
            raise ValueError("unsupported (subset_of_weights, hessian_structure)")
        if sigma_noise != 1 and likelihood != "regression":
            raise ValueError("Sigma noise != 1 only available for regression.")
        self.likelihood, self.structure, self.last_layer = likelihood, hessian_structure, subset_of_weights == "last_layer"
        self.model = LastLayerModel(model) if self.last_layer else model
        kwargs = dict(backend_kwargs or {})
        if self.last_layer:
            kwargs["last_layer"] = True
        self.backend = backend(self.model, likelihood, **kwargs)
        self.params = self.backend.params
        self.n_params = sum(p.numel() for p in self.params)
        self.n_layers = len(self.params)
        self.prior_precision = float(prior_precision)
        self.sigma_noise, self.temperature = float(sigma_noise), float(temperature)
        self.fisher_kwargs = dict(asdl_fisher_kwargs or {})
        self.damping = damping
        self.H = self.H_facs = None
        self.loss, self.n_data = 0.0, 0
        self._Sigma = None
        self._ll_cache = None
        self._copy_stream = None

    @property
    def _device(self):
        return next(self.model.parameters()).device

    @property
    def _H_factor(self):
        return 1.0 / (self.sigma_noise ** 2) / self.temperature

    # ------------------------------------------------------------------ fit (baselaplace.py:904-987)
    def fit(self, train_loader, decompose: bool = True):
        self.model.eval()
        self.mean = parameters_to_vector(self.params).detach()
        N = len(train_loader.dataset)
        self.loss, H = 0.0, None
        for X, y in self._device_batches(train_loader):
            self.model.zero_grad()
            if self.structure == "kron":
                loss_b, H_b = self.backend.kron(X, y, N=N, **self.fisher_kwargs)
            elif self.structure == "full":
                loss_b, H_b = self.backend.full(X, y, N=N)
            else:
                loss_b, H_b = self.backend.diag(X, y, N=N)
            self.loss = self.loss + loss_b
            if H is None:
                H = H_b
            else:
                H += H_b
        self.n_data = N
        self._Sigma = self._ll_cache = None
        if self.structure == "kron":
            self.H_facs = H
            self.H = H.decompose(damping=self.damping) if decompose else None
        else:
            self.H = H
        return self

    def _device_batches(self, loader):
        """Yield ``(X, y)`` on the model's device.  Host batches are copied on a side stream one batch ahead, so the
        host-to-device transfer of batch ``i + 1`` (2 ms for 4096 CIFAR-sized images over PCIe) overlaps the curvature
        kernels of batch ``i``; the reference does one blocking ``.to(device)`` per batch (baselaplace.py:974)."""
        dev = self._device

        def to_dev(data):
            if isinstance(data, MutableMapping):
                return data, data[self.backend.dict_key_y].to(dev, non_blocking=True)
            X, y = data
            return X.to(dev, non_blocking=True), y.to(dev, non_blocking=True)

        if dev.type != "cuda" or os.environ.get("LPB_NO_PREFETCH") == "1":
            for data in loader:
                yield to_dev(data)
            return
        main = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            # one copy stream per instance: its allocator pool stays warm across fit() calls (a fresh stream starts with
            # an empty pool and pays synchronising cudaMallocs for the first batches)
            self._copy_stream = torch.cuda.Stream(dev)
        copy = self._copy_stream

        def fetch(data):
            with torch.cuda.stream(copy):
                X, y = to_dev(data)
            ev = torch.cuda.Event()
            ev.record(copy)
            return X, y, ev

        it = iter(loader)
        try:
            nxt = fetch(next(it))
        except StopIteration:
            return
        while nxt is not None:
            X, y, ev = nxt
            main.wait_event(ev)
            for t in (X, y):
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(main)   # allocated on the copy stream, consumed on the compute stream
            try:
                nxt = fetch(next(it))
            except StopIteration:
                nxt = None
            yield X, y

    def decompose(self):
        self.H = self.H_facs.decompose(damping=self.damping)

    def zero_curvature(self):
        """The curvature of an empty data shard, laid out like a fitted one (``Kron.init_from_model``,
        utils/matrix.py:33-77: ``[out x out, in x in]`` per weight, ``[n x n]`` per bias, ``parameters()`` order)."""
        dev = self._device
        if self.structure == "kron":
            dims = []
            for p in self.params:
                dims.append([p.shape[0]] if p.ndim == 1 else [p.shape[0], int(p.numel() // p.shape[0])])
            return B200Kron.zeros(dims, dev, torch.float32)
        if self.structure == "full":
            return torch.zeros(self.n_params, self.n_params, device=dev, dtype=self.params[0].dtype)
        return torch.zeros(self.n_params, device=dev, dtype=self.params[0].dtype)

    # ------------------------------------------------------------------ posterior
    @property
    def posterior_precision(self):
        if self.structure == "kron":
            dev, dt = self.H.eigenvectors[0][0].device, self.H.eigenvectors[0][0].dtype
            return self.H * self._H_factor + torch.tensor([self.prior_precision], device=dev, dtype=dt)
        if self.structure == "full":
            return self._H_factor * self.H + self.prior_precision * torch.eye(self.n_params, device=self.H.device,
                                                                              dtype=self.H.dtype)
        return self._H_factor * self.H + self.prior_precision

    @property
    def posterior_covariance(self):
        """``P^{-1}`` for the full structure (baselaplace.py:1634-1661).  Cholesky factorisation + inverse are
        cuSOLVER LIBRARY calls (``torch.linalg``), declared in DESIGN.md."""
        if self._Sigma is None:
            L = torch.linalg.cholesky(self.posterior_precision)
            self._Sigma = torch.cholesky_inverse(L)
        return self._Sigma

    @property
    def log_det_posterior_precision(self):
        if self.structure == "kron":
            return self.posterior_precision.logdet()
        if self.structure == "full":
            return torch.logdet(self.posterior_precision)
        return self.posterior_precision.log().sum()

    # ------------------------------------------------------------------ GLM predictive
    def functional_variance(self, Js):
        if self.structure == "kron":
            return self.posterior_precision.inv_square_form(Js)
        if self.structure == "diag":
            ll = getattr(Js, "_lpb_ll", None)
            if ll is not None:
                from .predictive import ll_diag_variance

                return ll_diag_variance(ll[0], ll[1], ll[2], 1.0 / self.posterior_precision).to(Js.dtype)
            Jf = Js.float().contiguous()
            var = (1.0 / self.posterior_precision).float().contiguous()
            out = torch.empty(Js.shape[0], Js.shape[1], Js.shape[1], device=Js.device, dtype=torch.float32)
            return K.batched_pair_dot(Jf, Jf, var, out).to(Js.dtype)
        ll = getattr(Js, "_lpb_ll", None)
        if ll is not None:
            return self._ll_full_variance(*ll).to(Js.dtype)
        # dense: Y = J Sigma (GEMM-NT, Sigma symmetric), then row-pair dots
        Bc, C, P = Js.shape
        Jf = Js.float().reshape(Bc * C, P).contiguous()
        Sig = self.posterior_covariance.float().contiguous()
        Y = torch.empty(Bc * C, P, device=Js.device, dtype=torch.float32)
        K.gemm_nt(K.Packed(Jf, None, K.F32, Bc * C, P), K.Packed(Sig, None, K.F32, P, P), Y, 1.0, accumulate=False)
        out = torch.empty(Bc, C, C, device=Js.device, dtype=torch.float32)
        return K.batched_pair_dot(Y.view(Bc, C, P), Jf.view(Bc, C, P), None, out).to(Js.dtype)

    def _ll_full_variance(self, phi, C, has_bias):
        """Structured last-layer form of baselaplace.py:1683-1684 (``laplace_b200/predictive.py``)."""
        from .predictive import ll_full_variance

        out, self._ll_cache = ll_full_variance(phi, C, has_bias, self.posterior_covariance, self._ll_cache)
        return out

    def glm_predictive_distribution(self, X):
        if self.last_layer:
            Js, f_mu = self.backend.last_layer_jacobians(X)
        else:
            Js, f_mu = self.backend.jacobians(X)
        return f_mu.detach(), self.functional_variance(Js).detach()

    def __call__(self, x, pred_type="glm", link_approx="probit"):
        if pred_type != "glm":
            raise ValueError("only the GLM predictive is on the B200 hot path")
        f_mu, f_var = self.glm_predictive_distribution(x)
        if self.likelihood == "regression":
            return f_mu, f_var
        if link_approx != "probit":
            raise ValueError("only link_approx='probit' is provided by this stand-alone driver")
        kappa = 1.0 / torch.sqrt(1.0 + math.pi / 8.0 * torch.diagonal(f_var, dim1=1, dim2=2))
        return torch.softmax(kappa * f_mu, dim=-1)
