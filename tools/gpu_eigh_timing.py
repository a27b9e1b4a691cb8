"""Where decompose() spends its time on the box: per-size library eigh timings, batched variants, the compact / padded
shortcuts of laplace_b200/matrix.py, and the hand-written Jacobi kernel."""
import os, sys, time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_b200 import B200GGN, kernels as K, matrix, models  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


def psd(n, dev, batch=None):
    g = torch.Generator(device=dev).manual_seed(n)
    shape = (n, 2 * n) if batch is None else (batch, n, 2 * n)
    X = torch.randn(*shape, device=dev, generator=g)
    return X @ X.transpose(-1, -2) / (2 * n)


def main():
    dev = "cuda"
    if os.environ.get("EIGH_SIZES") == "1":
        for n in (64, 128, 147, 256, 384, 512, 513, 576, 1024, 1152, 2304, 4608):
            H = psd(n, dev)
            print(f"eigh n={n}: {timed(lambda: torch.linalg.eigh(H)):.2f} ms", flush=True)
        for n, b in ((64, 5), (128, 5), (256, 5), (512, 9), (576, 5), (1152, 5)):
            H = psd(n, dev, b)
            print(f"eigh batched {b}x{n}: {timed(lambda: torch.linalg.eigh(H)):.2f} ms", flush=True)
    for n, b in ((64, 5), (128, 5), (128, 32)):
        H = psd(n, dev, b)
        print(f"jacobi kernel {b}x{n}: {timed(lambda: K.eigh_jacobi(H)):.2f} ms", flush=True)
    # the real thing
    model = models.make("resnet18").to(dev)
    torch.manual_seed(0)
    X, y = torch.randn(1024, 3, 32, 32, device=dev), torch.randint(10, (1024,), device=dev)
    be = B200GGN(model, "classification", precision="bf16x3")
    _, k = be.kron(X, y, N=50000)
    torch.cuda.synchronize()
    for compact in (False, True):
        for pad in (False, True):
            matrix.COMPACT_DEAD_COORDINATES, matrix.PAD_EIGH = compact, pad
            print(f"decompose compact={compact} pad={pad}: {timed(lambda: k.decompose(), 2):.1f} ms", flush=True)
    matrix.COMPACT_DEAD_COORDINATES, matrix.PAD_EIGH = True, True
    for nt in (2, 3, 4, 6, 8):
        matrix.N_EIGH_THREADS = nt
        print(f"decompose compact+pad, {nt} threads: {timed(lambda: k.decompose(), 3):.1f} ms", flush=True)
    kd1 = None
    matrix.N_EIGH_THREADS = 1
    kd1 = k.decompose()
    matrix.N_EIGH_THREADS = 4
    kd4 = k.decompose()
    torch.cuda.synchronize()
    worst = max(float((a - b).abs().max() / a.abs().max().clamp_min(1e-30)) for La, Lb in zip(kd1.eigenvalues, kd4.eigenvalues) for a, b in zip(La, Lb))
    print("threads vs serial eigenvalues max rel diff:", worst)
    live = [(tuple(H.shape), int((H.diagonal() != 0).sum())) for F in k.kfacs for H in F]
    print("factor sizes (n, live):", live)


if __name__ == "__main__":
    main()
