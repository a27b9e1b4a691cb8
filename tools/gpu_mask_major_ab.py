"""A/B timing of the mask-major vs row-major ``pack_cast_fused`` / ``relu_bwd`` kernels at the ResNet-18 B = 4096 shapes
(10 folded curvature columns).  Usage on a GPU box: ``python tools/gpu_mask_major_ab.py``."""
import torch

from laplace_b200 import kernels as K

DEV = "cuda"
SHAPES = [("stem 16x16x64", 4096 * 256, 64), ("layer1 8x8x64", 4096 * 64, 64), ("layer2 4x4x128", 4096 * 16, 128),
          ("layer3 2x2x256", 4096 * 4, 256), ("layer4 1x1x512", 4096, 512)]
REPS = 10


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for name, rows_y, cols in SHAPES:
    g = torch.randn(REPS * rows_y, cols, device=DEV)
    y = torch.randn(rows_y, cols, device=DEV).clamp_min(0)
    scale = torch.rand(cols, device=DEV) + 0.5
    res = {}
    for mode, thr in (("row-major", -1), ("mask-major", 0)):
        K.set_mask_major_min(thr)
        res[mode, "pack"] = timed(lambda: K.pack_cast_fused(g, K.BF16X3, scale, y))
        res[mode, "relu"] = timed(lambda: K.relu_bwd(g, y, REPS))
    K.set_mask_major_min(4 << 20)
    gb = g.numel() * 4 / 1e9
    print(f"{name:16s} mask {rows_y * cols * 4 / 2**20:6.0f} MiB  pack: row-major {res['row-major', 'pack']:.3f} ms ({(2 * gb + gb / REPS) * 1e3 / res['row-major', 'pack']:.0f} GB/s alg.)"
          f"  mask-major {res['mask-major', 'pack']:.3f} ms ({(2 * gb + gb / REPS) * 1e3 / res['mask-major', 'pack']:.0f} GB/s)"
          f" | relu_bwd: {res['row-major', 'relu']:.3f} -> {res['mask-major', 'relu']:.3f} ms", flush=True)
    del g, y
