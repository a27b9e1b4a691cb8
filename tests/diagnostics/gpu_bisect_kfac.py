"""Bisect the KFAC disagreement between the bench-shaped path (fused chains, implicit A factors, side stream) and the
unfused / explicit path, per block, and against the fp64 oracle at FULL ResNet-18 width."""
import os, sys, time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from laplace_b200 import B200GGN, kernels as K, models  # noqa: E402
from oracle import curvature_oracle as co  # noqa: E402  (tools/ is test infrastructure)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def blocks(k):
    return [h for F in k.kfacs for h in F]


def names(be):
    out = []
    for L in be._plan():
        if L.has_w:
            out += [f"{L.name}.B[{L.d_out}]", f"{L.name}.A[{L.d_in}]"]
        if L.has_b:
            out += [f"{L.name}.bias[{L.d_out}]"]
    return out


def compare(tag, ka, kb, nm, thr=2e-5):
    errs = [rel(a, b) for a, b in zip(blocks(ka), blocks(kb))]
    w = max(errs)
    print(f"{tag}: worst {w:.2e}")
    for n, e in zip(nm, errs):
        if e > thr:
            print(f"    {n}: {e:.2e}")
    return errs


def main():
    dev = "cuda"
    B = int(os.environ.get("BISECT_B", "1024"))
    model = models.make("resnet18").to(dev)
    torch.manual_seed(3)
    X, y = torch.randn(B, 3, 32, 32, device=dev), torch.randint(10, (B,), device=dev)
    N = 50000

    def run(fuse, overlap, implicit_a, splits=1, dev_bound=None):
        from laplace_b200 import conv_engine as ce

        ok = K.conv_patches_ok
        old = (ce.ELEMENTWISE_MIN_BATCH, ce.ELEMENTWISE_MIN_NUMEL)
        if dev_bound is not None:
            ce.ELEMENTWISE_MIN_BATCH, ce.ELEMENTWISE_MIN_NUMEL = dev_bound
        if not implicit_a:
            K.conv_patches_ok = lambda *a: False
        try:
            be = B200GGN(model, "classification", precision="bf16x3", fuse_elementwise=fuse)
            be.overlap_factors = overlap
            H = None
            step = B // splits
            for i in range(0, B, step):
                _, k = be.kron(X[i:i + step], y[i:i + step], N=N)
                H = k if H is None else H + k
            torch.cuda.synchronize()
            return be, H
        finally:
            K.conv_patches_ok = ok
            ce.ELEMENTWISE_MIN_BATCH, ce.ELEMENTWISE_MIN_NUMEL = old

    be, a1 = run(True, True, True)
    nm = names(be)
    print("fused:", be._fused, be.last_backward_mode)
    _, a2 = run(True, True, True)
    compare("a: fused+overlap run-to-run", a1, a2, nm)
    _, b = run(True, False, True)
    compare("b: fused, no overlap            vs a", b, a1, nm)
    _, c = run(False, False, True)
    compare("c: unfused, implicit A          vs a", c, a1, nm)
    _, d = run(False, False, False)
    compare("d: unfused, explicit A          vs a", d, a1, nm)
    _, e = run(False, False, False, splits=2)
    compare("e: d in 2 halves                vs d", e, d, nm)
    _, e2 = run(False, False, False, splits=2, dev_bound=(1, 1))
    compare("e2: halves, custom elementwise  vs d", e2, d, nm)
    _, e3 = run(False, False, False, splits=1, dev_bound=(1 << 30, 1 << 40))
    compare("e3: whole, torch elementwise    vs d", e3, d, nm)
    _, f = run(True, True, True, splits=2, dev_bound=(1, 1))
    compare("f: fused halves (forced custom) vs a", f, a1, nm)
    # fp64 oracle at full width
    t0 = time.perf_counter()
    md = models.make("resnet18").double()
    md.load_state_dict({k: v.double().cpu() for k, v in model.state_dict().items()})
    ref = None
    Xc, yc = X.cpu().double(), y.cpu()
    nb = B
    torch.set_num_threads(32)
    for i in range(0, nb, 128):
        _, kf = co.kfac_factors(md, "classification", Xc[i:i + 128], yc[i:i + 128], N=N)
        ref = kf if ref is None else [[p + q for p, q in zip(Fa, Fb)] for Fa, Fb in zip(ref, kf)]
    print(f"oracle fp64 ({nb} samples): {time.perf_counter() - t0:.1f} s")
    refb = [h for F in ref for h in F]
    for tag, kk in (("a fused+overlap", a1), ("c unfused implicit", c), ("d unfused explicit", d), ("e halves", e)):
        errs = [rel(h.cpu(), r) for h, r in zip(blocks(kk), refb)]
        print(f"{tag} vs fp64 oracle: worst {max(errs):.2e}")
        for n, er in zip(nm, errs):
            if er > 2e-5:
                print(f"    {n}: {er:.2e}")


if __name__ == "__main__":
    main()
