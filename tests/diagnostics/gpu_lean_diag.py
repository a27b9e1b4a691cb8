"""Per-factor error of the reduced-width ResNet-18 front-end test config under the lean-precision switches."""
import os, sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from laplace_b200 import B200GGN, backend as bk, conv_engine, models  # noqa: E402
from oracle import curvature_oracle as co  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def main():
    for width, B in ((16, 64), (64, 256)):
        model = models.make("resnet18", width=width)
        torch.manual_seed(5)
        X, y = torch.randn(2 * B, 3, 32, 32), torch.randint(10, (2 * B,))
        md = model.double()
        kfs = None
        torch.set_num_threads(32)
        for i in (0, B):
            _, kf = co.kfac_factors(md, "classification", X[i:i + B].double(), y[i:i + B], N=2 * B)
            kfs = kf if kfs is None else [[a + b for a, b in zip(Fa, Fb)] for Fa, Fb in zip(kfs, kf)]
        model = model.float().cuda()
        keep = conv_engine.ELEMENTWISE_MIN_BATCH
        conv_engine.ELEMENTWISE_MIN_BATCH = 0
        thr = bk.A_SINGLE_PRODUCT_MIN_ROWS
        for tag, prec, t in (("bf16x3", "bf16x3", thr), ("auto", "auto", thr), ("auto, lean off", "auto", 1 << 60), ("auto, lean from 1 row", "auto", 1)):
            bk.A_SINGLE_PRODUCT_MIN_ROWS = t
            be = B200GGN(model, "classification", precision=prec)
            H = None
            for i in (0, B):
                _, k = be.kron(X[i:i + B].cuda(), y[i:i + B].cuda(), N=2 * B)
                H = k if H is None else H + k
            names = []
            for L in be._plan():
                if L.has_w:
                    names += [f"{L.name}.B[{L.d_out}]", f"{L.name}.A[{L.d_in}]"]
                if L.has_b:
                    names += [f"{L.name}.bias"]
            errs = [rel(h.cpu(), r) for F, Fo in zip(H.kfacs, kfs) for h, r in zip(F, Fo)]
            top = sorted(zip(errs, names), reverse=True)[:6]
            print(f"width {width} B {B} [{tag}] worst {max(errs):.2e}: " + ", ".join(f"{n} {e:.1e}" for e, n in top), flush=True)
        bk.A_SINGLE_PRODUCT_MIN_ROWS = thr
        conv_engine.ELEMENTWISE_MIN_BATCH = keep


if __name__ == "__main__":
    main()
