"""Test helper: pick inputs on which a ReLU network's curvature is well defined in fp32.

A ReLU unit whose pre-activation is closer to zero than the rounding error of an fp32 forward pass has no defined mask:
ANY fp32 implementation (cuDNN, the reference on CPU, this backend) decides it by rounding noise, and the flipped unit
changes that sample's gradient at that unit by 100 %.  On layers with few positions per sample one flip moves the KFAC
``B`` factor by ``~1 / (batch * positions)`` -- 4e-5 rel-fro for ResNet-18's 2x2 maps at batch 1024 (measured: the
same backend run on two half batches differs from the whole batch by 1.0e-4 on exactly that factor, and by 3e-7
everywhere else; ``profiles/r02_kfac_bisect.md``).  Parity against the fp64 oracle is therefore asserted on samples
whose small-map ReLU pre-activations keep a margin from zero, the same way ``test_config1_mlp_parity_anchor`` picks its
data seed."""
import torch
from torch import nn


def relu_margin_per_sample(model64: nn.Module, X64: torch.Tensor, max_positions: int = 16, chunk: int = 256) -> torch.Tensor:
    """min over the ReLU inputs with at most ``max_positions`` spatial positions of ``|z| / rms(z)`` -- per sample."""
    margins = []
    cur = {}

    def hook(mod, inp, out=None):
        z = inp[0].detach()
        pos = 1 if z.dim() <= 2 else int(z[0, 0].numel())
        if pos > max_positions:
            return
        zz = z.reshape(z.shape[0], -1)
        m = (zz.abs() / zz.pow(2).mean().sqrt().clamp_min(1e-300)).min(1).values
        cur["m"] = m if "m" not in cur else torch.minimum(cur["m"], m)

    hs = [m.register_forward_pre_hook(hook) for m in model64.modules() if isinstance(m, nn.ReLU)]
    try:
        with torch.no_grad():
            for i in range(0, len(X64), chunk):
                cur.clear()
                model64(X64[i:i + chunk])
                margins.append(cur.get("m", torch.full((len(X64[i:i + chunk]),), float("inf"), dtype=X64.dtype)))
    finally:
        for h in hs:
            h.remove()
    return torch.cat(margins)


def safe_samples(model64: nn.Module, X64: torch.Tensor, n: int, tol: float = 2e-5, max_positions: int = 16) -> torch.Tensor:
    """Indices of the first ``n`` samples whose margin exceeds ``tol`` (raises if there are fewer)."""
    idx = torch.nonzero(relu_margin_per_sample(model64, X64, max_positions) > tol).squeeze(1)
    if idx.numel() < n:
        raise RuntimeError(f"only {idx.numel()} of {len(X64)} samples keep a ReLU margin > {tol}; generate more")
    return idx[:n]
