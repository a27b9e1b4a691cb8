"""A/B timing of the factor contraction: 128x128 single-CTA tiles vs 256x256 CTA-pair tiles (cta_group::2)."""
import os, sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from laplace_b200 import kernels as K

DEV = "cuda"


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    torch.manual_seed(0)
    for rows in (True, False):
        for kind, name, nprod in ((K.BF16X3, "x3", 3), (K.BF16, "x1", 1)):
            for d, Kc in ((4608, 8192), (2304, 32768), (1152, 131072), (4608, 65536)):
                X = torch.randn(Kc, d, device=DEV)
                p = K.pack_cast(X, kind) if rows else K.pack_rows(X, kind)
                H = torch.zeros(d, d, device=DEV)
                gemm = K.gemm_tn if rows else K.gemm_nt
                res = {}
                for mode in (0, 1):
                    K.set_gemm_tile_mode(mode)
                    ms = timeit(lambda: gemm(p, p, H, 1.0, True, symmetric=True))
                    # useful tensor work: upper triangle incl. diagonal tiles, nprod products
                    flops = 2.0 * d * d * Kc * nprod / 2
                    res[mode] = (ms, flops / ms / 1e9)
                K.set_gemm_tile_mode(-1)
                print(f"{'rows' if rows else 'kmaj'} {name} d={d} K={Kc}: single {res[0][0]:.3f} ms ({res[0][1]:.0f} TF/s eff)  "
                      f"pair {res[1][0]:.3f} ms ({res[1][1]:.0f} TF/s eff)  speed-up {res[0][0] / res[1][0]:.2f}x", flush=True)
                del X, p, H


if __name__ == "__main__":
    main()
