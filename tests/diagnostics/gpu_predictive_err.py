"""Where the GLM-predictive error comes from: the posterior-side kernels given OUR factors (oracle algebra in fp64 on the
factors the GPU produced) vs end to end, with the eigendecomposition in fp32 (hand-written Jacobi / library) and fp64."""
import os, sys

import torch
from torch.utils.data import DataLoader, TensorDataset

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from laplace_b200 import B200GGN, matrix, models  # noqa: E402
from laplace_b200.posterior import B200Laplace  # noqa: E402
from oracle import curvature_oracle as co  # noqa: E402
from oracle import kron_oracle as ko  # noqa: E402
from tests.fixtures import load  # noqa: E402

DEV = "cuda"


def relmax(a, ref):
    return float((a.cpu().double() - ref).abs().max() / ref.abs().max())


def kron_case(tag, model_d, X, y, lik, bs, prior, Xt):
    """model_d: fp64 CPU model."""
    import copy

    kfs = None
    for i in range(0, len(X), bs):
        _, kf = co.kfac_factors(model_d, lik, X[i:i + bs], y[i:i + bs], N=len(X))
        kfs = kf if kfs is None else [[a + b for a, b in zip(Fa, Fb)] for Fa, Fb in zip(kfs, kf)]
    Js, f = co.jacobians(model_d, Xt)
    delta = torch.tensor(prior, dtype=torch.float64)
    Qo, lo = ko.decompose(kfs)
    ref_e2e = ko.kron_inv_square_form(Qo, lo, delta, Js)
    m32 = copy.deepcopy(model_d).float().to(DEV)
    yd = y.to(DEV) if y.dtype == torch.long else y.float().to(DEV)
    for fp64n in (0, 100000):
        matrix.EIGH_FP64_MAX_N = fp64n
        la = B200Laplace(m32, lik, "all", "kron", prior_precision=prior).fit(
            DataLoader(TensorDataset(X.float().to(DEV), yd), batch_size=bs))
        f_mu, f_var = la.glm_predictive_distribution(Xt.float().to(DEV))
        # oracle algebra on OUR factors
        ours = [[h.cpu().double() for h in F] for F in la.H_facs.kfacs]
        Q2, l2 = ko.decompose(ours)
        ref_same = ko.kron_inv_square_form(Q2, l2, delta, Js)
        fac_err = max(float((h - ho).norm() / ho.norm()) for F, Fo in zip(ours, kfs) for h, ho in zip(F, Fo))
        print(f"{tag} eigh_fp64<= {fp64n}: factors {fac_err:.1e}; var vs oracle-on-our-factors {relmax(f_var, ref_same):.2e}; "
              f"e2e {relmax(f_var, ref_e2e):.2e}; oracle(our factors) vs oracle(own) {relmax(ref_same, ref_e2e):.2e}", flush=True)
    matrix.EIGH_FP64_MAX_N = 0


def main():
    golden = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_vectors.pt"), weights_only=False)
    for kind in ("mlp", "conv"):
        for lik in ("classification", "regression"):
            model, X, y, _ = load(golden, kind, lik)
            kron_case(f"golden {kind}/{lik}", model, X, y, lik, 5, 0.7, X)
    torch.manual_seed(2)
    X, y = torch.randn(1000, 784, dtype=torch.float64), torch.randint(10, (1000,))
    torch.manual_seed(0)
    md = models.make("mlp").double()
    kron_case("config1 mlp", md, X, y, "classification", 128, 1.0, X[:32])
    # full / diag / last-layer posteriors vs golden
    for hs in ("full", "diag"):
        for lik in ("classification", "regression"):
            model, X, y, rec = load(golden, "mlp", lik, dtype=torch.float32)
            yd = y.to(DEV) if y.dtype == torch.long else y.to(DEV)
            la = B200Laplace(model.to(DEV), lik, "all", hs, prior_precision=0.7).fit(
                DataLoader(TensorDataset(X.to(DEV), yd), batch_size=4))
            f_mu, f_var = la.glm_predictive_distribution(X.to(DEV))
            print(f"golden mlp {hs}/{lik}: var vs golden {relmax(f_var, rec[f'la_{hs}_f_var']):.2e}", flush=True)
    for lik in ("classification", "regression"):
        model, X, y, rec = load(golden, "mlp", lik, dtype=torch.float32)
        yd = y.to(DEV) if y.dtype == torch.long else y.to(DEV)
        la = B200Laplace(model.to(DEV), lik, "last_layer", "full", prior_precision=0.7).fit(
            DataLoader(TensorDataset(X.to(DEV), yd), batch_size=4))
        f_mu, f_var = la.glm_predictive_distribution(X.to(DEV))
        Sigma = ko.full_posterior_covariance(rec["ll_ggn_full"], torch.full((la.n_params,), 0.7, dtype=torch.float64))
        ref = ko.full_functional_variance(rec["ll_Js"], Sigma)
        Sig2 = ko.full_posterior_covariance(la.H.cpu().double(), torch.full((la.n_params,), 0.7, dtype=torch.float64))
        ref2 = ko.full_functional_variance(rec["ll_Js"], Sig2)
        print(f"golden mlp LL-full/{lik}: var vs golden {relmax(f_var, ref):.2e}; vs oracle on our H {relmax(f_var, ref2):.2e}", flush=True)


if __name__ == "__main__":
    main()
