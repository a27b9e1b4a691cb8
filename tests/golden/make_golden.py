"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

Imports the reference through ``oracle/ref_shim.py`` and records, for the fixtures of
the reference's own tests (tests/test_curv_backends_curvlinops.py:23-81, seed 711,
float64), the outputs of

* ``CurvatureInterface.jacobians`` / ``last_layer_jacobians``   (curvature/curvature.py:88-167)
* ``GGNInterface.full`` / ``.diag``, ``EFInterface.full`` / ``.diag`` (curvature/curvature.py:375-505)
* ``Kron.decompose`` + ``KronDecomposed`` ``*``, ``+``, ``inv_square_form``, ``logdet`` (utils/matrix.py)
* ``FullLaplace`` / ``DiagLaplace`` fit + GLM predictive (probit) (baselaplace.py)

into ``tests/golden/reference_vectors.pt`` (~2.5 MB).
"""
from __future__ import annotations

import os
import sys

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402

assert ref_shim.install(), "reference not mounted"

from laplace import Laplace  # noqa: E402
from laplace.curvature import EFInterface, GGNInterface  # noqa: E402
from laplace.utils import FeatureExtractor, Kron  # noqa: E402
from torch.utils.data import DataLoader, TensorDataset  # noqa: E402


def mlp():
    torch.manual_seed(711)
    return nn.Sequential(nn.Linear(3, 20), nn.Tanh(), nn.Linear(20, 2))


def convnet():
    torch.manual_seed(711)
    return nn.Sequential(nn.Conv2d(3, 4, 2, 2), nn.Flatten(), nn.Tanh(), nn.Linear(16, 20), nn.Tanh(), nn.Linear(20, 2))


def data(kind, likelihood):
    torch.manual_seed(711)
    X = torch.randn(10, 3) if kind == "mlp" else torch.randn(10, 3, 5, 5)
    if likelihood == "classification":
        y = torch.randint(2, (10,))
    else:
        y = torch.randn(10, 2)
    return X, y


def main():
    torch.set_default_dtype(torch.float64)
    out = {}
    for kind, make in (("mlp", mlp), ("conv", convnet)):
        for lik in ("classification", "regression"):
            model = make().double()
            X, y = data(kind, lik)
            X = X.double()
            if lik == "regression":
                y = y.double()
            rec = {"state_dict": {k: v.clone() for k, v in model.state_dict().items()}, "X": X, "y": y}
            ggn = GGNInterface(model, lik)
            Js, f = ggn.jacobians(X)
            rec["Js"], rec["f"] = Js, f
            rec["ggn_loss"], rec["ggn_full"] = ggn.full(X, y)
            _, rec["ggn_diag"] = ggn.diag(X, y)
            ef = EFInterface(model, lik)
            rec["ef_loss"], ef_full = ef.full(X, y)
            if kind == "mlp":  # conv: Gs (below) determines it; keep the fixture small
                rec["ef_full"] = ef_full
            _, rec["ef_diag"] = ef.diag(X, y)
            Gs, gl = ef.gradients(X, y)
            rec["Gs"], rec["grad_loss"] = Gs.detach(), gl.detach()
            # last-layer Jacobians through the reference FeatureExtractor
            fe = FeatureExtractor(make().double())
            fe.load_state_dict({"model." + k: v for k, v in rec["state_dict"].items()}, strict=False)
            fe.eval()
            ll = GGNInterface(fe, lik, last_layer=False)
            with torch.no_grad():
                fe.find_last_layer(X[:1])
            ll = GGNInterface(fe, lik, last_layer=True)
            rec["ll_Js"], rec["ll_f"] = ll.last_layer_jacobians(X)
            rec["ll_ggn_loss"], rec["ll_ggn_full"] = ll.full(X, y)

            # Full / Diag Laplace end-to-end with the in-tree backend
            loader = DataLoader(TensorDataset(X, y), batch_size=4)
            for hs in ("full", "diag"):
                la = Laplace(model, lik, "all", hs, backend=GGNInterface, prior_precision=0.7)
                la.fit(loader)
                f_mu, f_var = la._glm_predictive_distribution(X)
                if hs == "diag":  # the full H equals ggn_full by additivity; keep the fixture small
                    rec[f"la_{hs}_H"] = la.H.clone()
                rec[f"la_{hs}_loss"] = torch.as_tensor(la.loss).clone()
                rec[f"la_{hs}_f_mu"], rec[f"la_{hs}_f_var"] = f_mu, f_var
                if lik == "classification":
                    rec[f"la_{hs}_probit"] = la(X, pred_type="glm", link_approx="probit")
                rec[f"la_{hs}_logmarglik"] = la.log_marginal_likelihood().detach()
            if kind == "conv":  # 434^2 doubles -> store as float32 (compared at 1e-6)
                rec["ggn_full"] = rec["ggn_full"].float()
            out[f"{kind}_{lik}"] = rec

    # Kron algebra on random PSD factors laid out like Kron.init_from_model(mlp)
    torch.manual_seed(7)

    def psd(d):
        Z = torch.randn(d, 3 * d)
        return Z @ Z.T / (3 * d)

    kfacs = [[psd(20), psd(3)], [psd(20)], [psd(2), psd(20)], [psd(2)]]
    kron = Kron([[F.clone() for F in Fs] for Fs in kfacs])
    W = torch.randn(5, 2, 20 * 3 + 20 + 2 * 20 + 2)
    rec = {"kfacs": kfacs, "W": W}
    for damping in (False, True):
        kd = kron.decompose(damping=damping)
        tag = "damp" if damping else "plain"
        rec[f"{tag}_eigvals"] = [[l.clone() for l in ls] for ls in kd.eigenvalues]
        for name, delta in (("scalar", torch.tensor(0.3)), ("layer", torch.tensor([0.3, 1.1, 0.05, 2.0]))):
            P = kd * 1.7 + delta
            if damping:
                # NB reference quirk: ``*`` / ``+`` rebuild the object with damping=False
                # (utils/matrix.py:355,376), so construct the damped operator directly.
                from laplace.utils import KronDecomposed

                P = KronDecomposed(P.eigenvectors, P.eigenvalues, P.deltas, damping=True)
            rec[f"{tag}_{name}_delta"] = delta
            rec[f"{tag}_{name}_isf"] = P.inv_square_form(W)
            rec[f"{tag}_{name}_logdet"] = P.logdet()
            rec[f"{tag}_{name}_bmm_m05"] = P.bmm(W, exponent=-0.5)
    out["kron_algebra"] = rec

    path = os.path.join(HERE, "reference_vectors.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
