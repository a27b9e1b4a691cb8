import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_b200 import B200GGN, models, conv_engine
DEV = "cuda"
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
name, kw = "wrn28_10", {"depth": 10, "widen": 2}
model = models.make(name, **kw).to(DEV)
torch.manual_seed(3)
X = torch.randn(16, 3, 32, 32, device=DEV)
md = models.make(name, **kw).double().to(DEV); md.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
b64 = B200GGN(md, "classification", conv_engine=False)
f64 = b64._forward(X.double()); cols = b64._hessian_sqrt_cols(f64.detach()); g64 = b64._backward(f64, cols)
acts64 = dict(b64._acts)
for label, kwargs, implicit in (("engine batched implicit", {}, True), ("engine batched explicit", {}, False), ("engine loop implicit", {"batched_backward": False}, True),
                               ("cudnn fp32 functorch?", {"conv_engine": False}, True)):
    conv_engine.USE_IMPLICIT = implicit
    be = B200GGN(model, "classification", **kwargs)
    f = be._forward(X); acts = dict(be._acts); grads = be._backward(f, cols.float())
    print(label, be.last_backward_mode if hasattr(be, "last_backward_mode") else "-", "f", f"{rel(f, f64):.1e}")
    print("   acts :", " ".join(f"{rel(acts[L.name], acts64[L.name]):.1e}" for L in be._layers))
    print("   grads:", " ".join(f"{rel(a, b):.1e}" for a, b in zip(grads, g64)))
