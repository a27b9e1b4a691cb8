import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_b200 import B200GGN, models, kernels as K
from laplace_b200.matrix import B200Kron, symeig_large
DEV = "cuda"
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
for name, kw in (("resnet18", {"width": 16}), ("wrn28_10", {"depth": 10, "widen": 2}), ("resnet18", {})):
    model = models.make(name, **kw).to(DEV)
    torch.manual_seed(3)
    X = torch.randn(16, 3, 32, 32, device=DEV)
    be = B200GGN(model, "classification")
    f = be._forward(X); cols = be._hessian_sqrt_cols(f.detach()); grads = be._backward(f, cols)
    md = models.make(name, **kw).double().to(DEV); md.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    b64 = B200GGN(md, "classification", conv_engine=False)
    f64 = b64._forward(X.double()); g64 = b64._backward(f64, cols.double())
    b32 = B200GGN(model, "classification", conv_engine=False)
    f32 = b32._forward(X); g32 = b32._backward(f32, cols)
    print(name, kw, "f err engine", rel(f, f64), "cudnn fp32", rel(f32, f64))
    print("  layer grad errs engine:", " ".join(f"{rel(a,b):.1e}" for a, b in zip(grads, g64)))
    print("  layer grad errs cudnn :", " ".join(f"{rel(a,b):.1e}" for a, b in zip(g32, g64)))
# decompose timing
model = models.make("resnet18").to(DEV)
be = B200GGN(model, "classification", precision="bf16x3")
X = torch.randn(512, 3, 32, 32, device=DEV); y = torch.randint(10, (512,), device=DEV)
H = None
for i in range(3):
    _, kr = be.kron(X, y, N=50000)
    if H is None: H = kr
    else: H += kr
torch.cuda.synchronize()
print("finite:", all(bool(torch.isfinite(h).all()) for F in H.kfacs for h in F))
for rep in range(2):
    t = time.perf_counter(); kd = H.decompose(); torch.cuda.synchronize(); print("decompose total s", time.perf_counter() - t)
sizes = {}
for F in H.kfacs:
    for h in F:
        n = h.shape[0]
        if n in sizes: continue
        t = time.perf_counter()
        if n <= 128: K.eigh_jacobi(h.unsqueeze(0))
        else: symeig_large(h)
        torch.cuda.synchronize(); sizes[n] = time.perf_counter() - t
print("per-size first-call seconds:", {k: round(v, 4) for k, v in sorted(sizes.items())})
for n in (64, 128):
    hs = torch.stack([h for F in H.kfacs for h in F if h.shape[0] == n])
    t = time.perf_counter(); K.eigh_jacobi(hs); torch.cuda.synchronize(); print("jacobi batch", n, hs.shape[0], time.perf_counter() - t)
    t = time.perf_counter(); torch.linalg.eigh(hs); torch.cuda.synchronize(); print("cusolver batch", n, time.perf_counter() - t)
