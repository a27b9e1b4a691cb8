import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_b200 import B200GGN, B200EF, models
dev = "cuda"
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n
m = models.make("wrn28_10").to(dev)
print("WRN-28-10 params", sum(p.numel() for p in m.parameters() if p.requires_grad))
X = torch.randn(256, 3, 32, 32, device=dev); y = torch.randint(10, (256,), device=dev)
ef = B200EF(m, "classification")
t = timeit(lambda: ef.diag(X[:64], y[:64]), 2); print(f"WRN diag EF: {64/t:.0f} samples/s (B=64)", torch.cuda.max_memory_allocated() / 2**30, "GiB")
g = B200GGN(m, "classification")
t = timeit(lambda: g.kron(X, y, N=50000), 2); print(f"WRN KFAC-GGN: {256/t:.0f} samples/s (B=256)", torch.cuda.max_memory_allocated() / 2**30, "GiB")
t = timeit(lambda: ef.kron(X, y, N=50000), 2); print(f"WRN KFAC-EF: {256/t:.0f} samples/s (B=256)")
