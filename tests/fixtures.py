"""Model / data fixtures mirroring the reference's own test fixtures
(tests/test_curv_backends_curvlinops.py:23-81: seed 711, MLP 3-20-2 with Tanh, the
``Conv2d(3,4,2,2)`` "complex model", 10 samples)."""
import torch
from torch import nn


def mlp(dtype=torch.float64):
    torch.manual_seed(711)
    return nn.Sequential(nn.Linear(3, 20), nn.Tanh(), nn.Linear(20, 2)).to(dtype)


def convnet(dtype=torch.float64):
    torch.manual_seed(711)
    return nn.Sequential(
        nn.Conv2d(3, 4, 2, 2), nn.Flatten(), nn.Tanh(), nn.Linear(16, 20), nn.Tanh(), nn.Linear(20, 2)
    ).to(dtype)


MODELS = {"mlp": mlp, "conv": convnet}


def load(golden, kind, likelihood, dtype=torch.float64):
    rec = golden[f"{kind}_{likelihood}"]
    model = MODELS[kind](dtype)
    model.load_state_dict({k: v.to(dtype) for k, v in rec["state_dict"].items()})
    X = rec["X"].to(dtype)
    y = rec["y"] if likelihood == "classification" else rec["y"].to(dtype)
    return model, X, y, rec


def rel_fro(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))
