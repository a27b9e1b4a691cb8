"""smoke()'s predictive check under different eigensolver routes: which part of the posterior-side error is the fp32
eigendecomposition (hand-written Jacobi n <= 128, library syevj n <= 512, library syevd beyond / padded)."""
import copy, os, sys

import torch
from torch.utils.data import DataLoader, TensorDataset

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from laplace_b200 import matrix  # noqa: E402
from laplace_b200.posterior import B200Laplace  # noqa: E402
from oracle import curvature_oracle as co  # noqa: E402
from oracle import kron_oracle as ko  # noqa: E402


def main():
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, 1, 1), torch.nn.ReLU(), torch.nn.Flatten(),
                                torch.nn.Linear(8 * 8 * 8, 64), torch.nn.Tanh(), torch.nn.Linear(64, 10)).eval()
    X, y = torch.randn(512, 3, 8, 8), torch.randint(10, (512,))
    md = copy.deepcopy(model).double()
    model = model.float().cuda()
    Js, _ = co.jacobians(md, X[:16].double())
    one = torch.tensor(1.0, dtype=torch.float64)
    for tag, pad, fp64n, jac in (("pad on", True, 0, True), ("pad off (syevj)", False, 0, True), ("fp64 eigh everywhere", True, 10 ** 6, True),
                                 ("fp64 for n>128 only", True, 10 ** 6, False)):
        matrix.PAD_EIGH, matrix.EIGH_FP64_MAX_N = pad, fp64n
        keep = matrix.K.EIGH_MAX_N
        if fp64n and not jac:
            pass
        la = B200Laplace(model, "classification", "all", "kron", prior_precision=1.0).fit(
            DataLoader(TensorDataset(X.cuda(), y.cuda()), batch_size=256), decompose=False)
        if fp64n and not jac:
            # library fp64 only beyond the Jacobi kernel's range
            matrix.EIGH_FP64_MAX_N = 0
            small = la.H_facs.decompose()
            matrix.EIGH_FP64_MAX_N = fp64n
            big = la.H_facs.decompose()
            for i, F in enumerate(la.H_facs.kfacs):
                for j, H in enumerate(F):
                    if H.shape[0] <= 128:
                        big.eigenvectors[i][j], big.eigenvalues[i][j] = small.eigenvectors[i][j], small.eigenvalues[i][j]
            la.H = big
        else:
            la.decompose()
        f_mu, f_var = la.glm_predictive_distribution(X[:16].cuda())
        Qs, ls = ko.decompose([[h.cpu().double() for h in F] for F in la.H_facs.kfacs])
        ref = ko.kron_inv_square_form(Qs, ls, one, Js)
        err = float((f_var.cpu().double() - ref).abs().max() / ref.abs().max())
        # eigen-residuals of our decomposition per factor
        res = []
        for F, Q, L in zip(la.H_facs.kfacs, la.H.eigenvectors, la.H.eigenvalues):
            for h, q, l in zip(F, Q, L):
                hd, qd, ld = h.double(), q.double(), l.double()
                r = float((hd @ qd - qd * ld).norm() / hd.norm())
                o = float((qd.T @ qd - torch.eye(len(ld), device=qd.device, dtype=torch.float64)).norm())
                res.append(f"n={len(ld)}: resid {r:.1e} orth {o:.1e}")
        print(f"{tag}: var err on the same factors {err:.2e} | " + "; ".join(res), flush=True)
    matrix.PAD_EIGH, matrix.EIGH_FP64_MAX_N = True, 0


if __name__ == "__main__":
    main()
