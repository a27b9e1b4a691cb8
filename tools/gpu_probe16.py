import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_b200 import B200GGN, models
from oracle import curvature_oracle as co
def rel(a, b): return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())
torch.manual_seed(1)
X, y = torch.randn(1000, 784), torch.randint(10, (1000,))
md = models.make("mlp").double()
model = models.make("mlp").cuda()
for label, kw in (("default", {}), ("fp32", {"precision": "fp32"}), ("no engine", {"conv_engine": False})):
    be = B200GGN(model, "classification", **kw)
    tot, toto = None, None
    for i in range(0, 1000, 128):
        _, kf = co.kfac_factors(md, "classification", X[i:i+128].double(), y[i:i+128], N=1000)
        _, kr = be.kron(X[i:i+128].cuda(), y[i:i+128].cuda(), N=1000)
        print(label, "batch", i // 128, [[f"{rel(h, ho):.1e}" for h, ho in zip(F, Fo)] for F, Fo in zip(kr.kfacs, kf)])
        if tot is None: tot, toto = kr, kf
        else:
            tot += kr; toto = [[a + b for a, b in zip(Fa, Fb)] for Fa, Fb in zip(toto, kf)]
    print(label, "TOTAL", [[f"{rel(h, ho):.1e}" for h, ho in zip(F, Fo)] for F, Fo in zip(tot.kfacs, toto)])
# mask agreement of the hidden layer vs fp64
with torch.no_grad():
    h64 = md[0](X.double()); h32 = model[0](X.cuda()).cpu().double()
print("mask disagreements cuBLAS fp32 vs fp64:", int(((h64 > 0) != (h32 > 0)).sum()), "min |pre-act|", float(h64.abs().min()))
