"""Where does the e2e gap come from?  B200Laplace.fit over (a) pinned host batches with the copy-stream prefetch,
(b) device-resident batches through the same fit loop, (c) the bare backend loop of bench.py's `value` leg."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_b200 import B200GGN, models
from laplace_b200.posterior import B200Laplace

dev = torch.device("cuda", 0)
model = models.make("resnet18").to(dev)
B, n = 4096, 8
torch.manual_seed(1)
Xs = [torch.randn(B, 3, 32, 32, device=dev) for _ in range(n)]
ys = [torch.randint(10, (B,), device=dev) for _ in range(n)]
Xh = [x.cpu().pin_memory() for x in Xs]
yh = [y.cpu().pin_memory() for y in ys]


class DS(torch.utils.data.Dataset):
    def __len__(self):
        return 50000


class Loader:
    dataset = DS()

    def __init__(self, Xl, yl, k):
        self.Xl, self.yl, self.k = Xl, yl, k

    def __iter__(self):
        for i in range(self.k):
            yield self.Xl[i % n], self.yl[i % n]


def timed(fn, steps):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


la = B200Laplace(model, "classification", "all", "kron", backend=B200GGN, backend_kwargs={"precision": "bf16x3"})
la.fit(Loader(Xh, yh, 3), decompose=False)
print(f"fit, pinned host batches (prefetch):   {timed(lambda: la.fit(Loader(Xh, yh, 12), decompose=False), 12):.2f} ms/step")
print(f"fit, device-resident batches:          {timed(lambda: la.fit(Loader(Xs, ys, 12), decompose=False), 12):.2f} ms/step")
be = B200GGN(model, "classification", precision="bf16x3")
H = None


def bare():
    global H
    for i in range(12):
        _, kr = be.kron(Xs[i % n], ys[i % n], N=50000)
        if H is None:
            H = kr
        else:
            H += kr


bare()
print(f"bare backend loop (bench value leg):   {timed(bare, 12):.2f} ms/step")
# H2D alone
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); s.record()
for i in range(8):
    Xh[i].to(dev, non_blocking=True)
e.record(); torch.cuda.synchronize()
print(f"H2D of one batch alone:                {s.elapsed_time(e) / 8:.2f} ms")
