"""Forward-pass accuracy of the convolution engine against fp64, per ReLU input (relative to the layer's rms), and the
number of ReLU units whose mask differs from the fp64 mask -- with the engine on, with cuDNN fp32, with cuDNN TF32."""
import os, sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_b200 import B200GGN, models  # noqa: E402


def capture(model, X, ctx=None):
    zs = []
    hs = [m.register_forward_pre_hook(lambda mod, inp: zs.append(inp[0].detach().double().cpu())) for m in model.modules()
          if isinstance(m, nn.ReLU)]
    try:
        with torch.no_grad():
            if ctx is None:
                model(X)
            else:
                with ctx:
                    model(X)
    finally:
        for h in hs:
            h.remove()
    return zs


def main():
    B = 1024
    md = models.make("resnet18").double()
    torch.manual_seed(3)
    Xc = torch.randn(B, 3, 32, 32, dtype=torch.float64)
    ref = capture(md, Xc)
    model = models.make("resnet18").cuda()
    X = Xc.float().cuda()
    be = B200GGN(model, "classification", precision="bf16x3")
    runs = {"engine": lambda: capture(model, X, be._conv_patch(False))}
    def cudnn(tf32):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        return capture(model, X)
    runs["cudnn fp32"] = lambda: cudnn(False)
    runs["cudnn tf32"] = lambda: cudnn(True)
    for tag, fn in runs.items():
        zs = fn()
        print(tag)
        for i, (z, r) in enumerate(zip(zs, ref)):
            rms = r.pow(2).mean().sqrt()
            err = ((z - r).abs().max() / rms).item()
            flips = int(((z > 0) != (r > 0)).sum())
            print(f"  relu#{i:2d} shape {tuple(r.shape[1:])}: max|dz|/rms {err:.2e}, mask flips {flips} of {r.numel()}")


if __name__ == "__main__":
    main()
