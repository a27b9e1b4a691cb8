"""Synthetic, random-init model shapes of the BASELINE.json configs (no datasets / checkpoints exist
offline).  All curvature-bearing parameters live in explicit ``nn.Linear`` / ``nn.Conv2d`` modules;
normalisation layers carry frozen affine parameters (``requires_grad=False``) so that the reference's
KFAC path and the B200 backend cover the same parameter set (hard part 5)."""
from __future__ import annotations

import torch
from torch import nn


def mlp(d_in=784, hidden=128, n_out=10) -> nn.Module:
    """Config 1: 2-layer MLP 784 -> 128 -> 10."""
    return nn.Sequential(nn.Linear(d_in, hidden), nn.ReLU(), nn.Linear(hidden, n_out))


def _frozen_bn(ch: int) -> nn.BatchNorm2d:
    bn = nn.BatchNorm2d(ch)
    for p in bn.parameters():
        p.requires_grad_(False)
    return bn


class BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = _frozen_bn(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = _frozen_bn(cout)
        self.relu = nn.ReLU()
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), _frozen_bn(cout))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + idt)


class ResNet18(nn.Module):
    """Configs 2/3: torchvision ResNet-18 topology (7x7/2 stem + max-pool, 4x2 basic blocks, fc 512->C).
    ``cifar_stem=True`` swaps in the 3x3/1 stem without max-pool (16x more spatial positions per layer)."""

    def __init__(self, n_out=10, width=64, cifar_stem=False):
        super().__init__()
        if cifar_stem:
            self.conv1 = nn.Conv2d(3, width, 3, 1, 1, bias=False)
            self.maxpool = nn.Identity()
        else:
            self.conv1 = nn.Conv2d(3, width, 7, 2, 3, bias=False)
            self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.bn1 = _frozen_bn(width)
        self.relu = nn.ReLU()
        chans = [width, 2 * width, 4 * width, 8 * width]
        blocks, cin = [], width
        for i, c in enumerate(chans):
            blocks += [BasicBlock(cin, c, 1 if i == 0 else 2), BasicBlock(c, c, 1)]
            cin = c
        self.layers = nn.Sequential(*blocks)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(cin, n_out)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layers(x)
        return self.fc(torch.flatten(self.avgpool(x), 1))


class _WideBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.bn1 = _frozen_bn(cin)
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn2 = _frozen_bn(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.short = None if (cin == cout and stride == 1) else nn.Conv2d(cin, cout, 1, stride, 0, bias=False)
        self.relu = nn.ReLU()

    def forward(self, x):
        o = self.relu(self.bn1(x))
        y = self.conv1(o)
        y = self.conv2(self.relu(self.bn2(y)))
        return y + (x if self.short is None else self.short(o))


class WideResNet(nn.Module):
    """Config 4: WideResNet-d-k shape (d=28, k=10 -> 36.5 M parameters), pre-activation blocks as in the
    reference's ``examples/helper/wideresnet.py``."""

    def __init__(self, depth=28, widen=10, n_out=10):
        super().__init__()
        n = (depth - 4) // 6
        w = [16, 16 * widen, 32 * widen, 64 * widen]
        self.conv1 = nn.Conv2d(3, w[0], 3, 1, 1, bias=False)
        blocks, cin = [], w[0]
        for i in range(3):
            for j in range(n):
                blocks.append(_WideBlock(cin, w[i + 1], (1 if i == 0 else 2) if j == 0 else 1))
                cin = w[i + 1]
        self.blocks = nn.Sequential(*blocks)
        self.bn = _frozen_bn(cin)
        self.relu = nn.ReLU()
        self.fc = nn.Linear(cin, n_out)

    def forward(self, x):
        x = self.blocks(self.conv1(x))
        x = self.relu(self.bn(x)).mean((2, 3))
        return self.fc(x)


class _ViTBlock(nn.Module):
    def __init__(self, dim, heads, mlp_dim):
        super().__init__()
        self.heads = heads
        self.ln1 = nn.LayerNorm(dim, elementwise_affine=False)
        self.q, self.k, self.v, self.o = (nn.Linear(dim, dim) for _ in range(4))
        self.ln2 = nn.LayerNorm(dim, elementwise_affine=False)
        self.fc1, self.fc2 = nn.Linear(dim, mlp_dim), nn.Linear(mlp_dim, dim)

    def forward(self, x):
        B, T, D = x.shape
        h = self.ln1(x)
        q, k, v = (m(h).view(B, T, self.heads, D // self.heads).transpose(1, 2) for m in (self.q, self.k, self.v))
        att = torch.softmax(q @ k.transpose(-1, -2) / (D // self.heads) ** 0.5, dim=-1)
        x = x + self.o((att @ v).transpose(1, 2).reshape(B, T, D))
        return x + self.fc2(torch.nn.functional.gelu(self.fc1(self.ln2(x))))


class ViT(nn.Module):
    """Config 5: ViT-B/16 shape with explicit q/k/v/o and MLP ``nn.Linear`` layers (86 M parameters at
    the default sizes); patch embedding, position embedding and head are frozen except the linears."""

    def __init__(self, image=224, patch=16, dim=768, depth=12, heads=12, mlp_dim=3072, n_out=10):
        super().__init__()
        self.patch = patch
        self.embed = nn.Linear(3 * patch * patch, dim)
        self.cls = nn.Parameter(torch.zeros(1, 1, dim), requires_grad=False)
        self.pos = nn.Parameter(torch.randn(1, (image // patch) ** 2 + 1, dim) * 0.02, requires_grad=False)
        self.blocks = nn.Sequential(*[_ViTBlock(dim, heads, mlp_dim) for _ in range(depth)])
        self.ln = nn.LayerNorm(dim, elementwise_affine=False)
        self.head = nn.Linear(dim, n_out)

    def forward(self, x):
        B, C, H, W = x.shape
        p = self.patch
        x = x.unfold(2, p, p).unfold(3, p, p).permute(0, 2, 3, 1, 4, 5).reshape(B, -1, C * p * p)
        x = torch.cat([self.cls.expand(B, -1, -1), self.embed(x)], 1) + self.pos
        return self.head(self.ln(self.blocks(x))[:, 0])


def make(name: str, **kw) -> nn.Module:
    torch.manual_seed(0)
    table = {"mlp": mlp, "resnet18": ResNet18, "wrn28_10": WideResNet, "vit_b16": ViT}
    return table[name](**kw).eval()
