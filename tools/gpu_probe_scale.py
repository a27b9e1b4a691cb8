"""Bench-scale consistency probe (ResNet-18, B = 4096): KFAC factors of the fused / im2col-free path vs the unfused
explicit path of the same backend, finiteness, and where decompose() spends its time."""
import os, sys, time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_b200 import B200GGN, kernels as K, models  # noqa: E402


def main():
    dev = "cuda"
    model = models.make("resnet18").to(dev)
    torch.manual_seed(0)
    X, y = torch.randn(4096, 3, 32, 32, device=dev), torch.randint(10, (4096,), device=dev)
    be = B200GGN(model, "classification", precision="bf16x3")
    _, k1 = be.kron(X, y, N=50000)
    torch.cuda.synchronize()
    print("fused:", be._fused, be.fuse_elementwise, be.last_backward_mode)
    ok = K.conv_patches_ok
    K.conv_patches_ok = lambda *a: False
    be2 = B200GGN(model, "classification", precision="bf16x3", fuse_elementwise=False)
    _, k2 = be2.kron(X, y, N=50000)
    K.conv_patches_ok = ok
    torch.cuda.synchronize()
    worst = 0.0
    for i, (Fa, Fb) in enumerate(zip(k1.kfacs, k2.kfacs)):
        for a, b in zip(Fa, Fb):
            assert torch.isfinite(a).all() and torch.isfinite(b).all(), i
            r = float((a - b).norm() / b.norm())
            s = float((a - a.t()).norm() / a.norm())
            worst = max(worst, r)
            if r > 2e-5 or s > 1e-5:
                print(f"  block {i} shape {tuple(a.shape)}: fused-vs-unfused rel {r:.2e}, asym {s:.2e}")
    print(f"worst fused/implicit vs unfused/explicit factor rel-fro: {worst:.2e}")
    # decompose timing by factor size
    for n_w in (256, 1024, 4096):
        torch.linalg.eigh(torch.eye(n_w, device=dev) + 0.01)
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        k1.decompose()
        torch.cuda.synchronize()
        print(f"decompose #{rep}: {(time.perf_counter() - t0) * 1e3:.0f} ms")
    sizes = {}
    for F in k1.kfacs:
        for H in F:
            if H.shape[0] > 128:
                t0 = time.perf_counter()
                torch.linalg.eigh(H)
                torch.cuda.synchronize()
                sizes.setdefault(H.shape[0], []).append((time.perf_counter() - t0) * 1e3)
    for n, ts in sorted(sizes.items()):
        print(f"  eigh n={n}: {len(ts)} factors, {sum(ts) / len(ts):.1f} ms each (serial), total {sum(ts):.0f} ms")


if __name__ == "__main__":
    main()
