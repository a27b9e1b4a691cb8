import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_b200 import B200GGN, models, kernels as K, conv_engine
from oracle import curvature_oracle as co
def rel(a, b): return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())
torch.manual_seed(1)
X, y = torch.randn(128, 784), torch.randint(10, (128,))
md = models.make("mlp").double()
_, kf = co.kfac_factors(md, "classification", X.double(), y, N=1000)
model = models.make("mlp").cuda()
for label, kw in (("default", {}), ("fp32", {"precision": "fp32"}), ("no engine", {"conv_engine": False}), ("bf16x3 no-engine", {"conv_engine": False, "precision": "bf16x3"})):
    be = B200GGN(model, "classification", **kw)
    _, kr = be.kron(X.cuda(), y.cuda(), N=1000)
    print(label, [[f"{rel(h, ho):.1e}" for h, ho in zip(F, Fo)] for F, Fo in zip(kr.kfacs, kf)])
# direct kernel check on the actual gradient rows
be = B200GGN(model, "classification")
f = be._forward(X.cuda()); cols = be._hessian_sqrt_cols(f.detach()); grads = be._backward(f, cols)
g = grads[0].reshape(-1, 128).contiguous()
ref = g.double().t() @ g.double()
for kind, name in ((K.BF16X3, "bf16x3"), (K.F16X3, "fp16x3")):
    P = K.pack_cast(g, kind); out = torch.zeros(128, 128, device="cuda"); K.gemm_tn(P, P, out, 1.0, True, symmetric=True)
    print("gemm_tn on grads", name, rel(out, ref), "absmax", float(g.abs().max()), "min nonzero", float(g[g != 0].abs().min()))
    Pt = K.pack_rows(g, kind); out2 = torch.zeros(128, 128, device="cuda"); K.gemm_nt(Pt, Pt, out2, 1.0, True, symmetric=True)
    print("gemm_nt on grads", name, rel(out2, ref))
st = conv_engine.STASH.get(id(model[0]), {})
print("stash keys", list(st.keys()), st["G"].rows if "G" in st else None, g.shape)
if "G" in st:
    G = st["G"]; dense = G.hi[:, :128].float() + G.lo[:, :128].float()
    print("stash G vs grads rel", rel(dense, g))
