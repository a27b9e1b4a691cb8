"""A/B timing of the factor contraction schedules: one 128x128 tile per CTA (0), 256x256 CTA-pair tiles (1,
cta_group::2), persistent CTAs with double-buffered TMEM (2) -- on the SYRK shapes of one ResNet-18 KFAC step at
batch 4096 (10 curvature columns) and a few large ones."""
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
    step_shapes = ((64, 2621440), (128, 655360), (256, 163840), (512, 40960), (576, 262144), (1152, 65536), (2304, 16384),
                   (4608, 4096), (4608, 65536))
    for rows in (True, False):
        for kind, name, nprod in ((K.BF16X3, "x3", 3), (K.BF16, "x1", 1)):
            for d, Kc in (step_shapes if rows and nprod == 3 else ((4608, 8192), (1152, 131072))):
                X = torch.randn(Kc, d, device=DEV)
                p = K.pack_cast(X, kind) if rows else K.pack_rows(X, kind)
                H = torch.zeros(d, d, device=DEV)
                gemm = K.gemm_tn if rows else K.gemm_nt
                res = {}
                for mode in (0, 1, 2):
                    K.set_gemm_tile_mode(mode)
                    ms = timeit(lambda: gemm(p, p, H, 1.0, True, symmetric=True))
                    # useful tensor work: upper triangle incl. diagonal tiles, nprod products
                    flops = 2.0 * d * d * Kc * nprod / 2
                    res[mode] = (ms, flops / ms / 1e9)
                K.set_gemm_tile_mode(-1)
                print(f"{'rows' if rows else 'kmaj'} {name} d={d} K={Kc}: single {res[0][0]:.3f} ms ({res[0][1]:.0f} TF/s eff)  "
                      f"pair {res[1][0]:.3f} ms ({res[1][1]:.0f})  persistent {res[2][0]:.3f} ms ({res[2][1]:.0f})", flush=True)
                del X, p, H


if __name__ == "__main__":
    main()
