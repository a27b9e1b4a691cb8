"""Where does a KFAC step go?  CUDA-event timing of the phases of B200GGN.kron on the bench workload.
(diagnostic only -- numbers taken with per-call events are never bench values)"""
import argparse, collections, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_b200 import B200GGN, kernels as K, models

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--model", default="resnet18")
ap.add_argument("--tf32", action="store_true")
ap.add_argument("--loop-backward", action="store_true")
ap.add_argument("--no-engine", action="store_true")
a = ap.parse_args()
dev = "cuda"
model = models.make(a.model).to(dev)
be = B200GGN(model, "classification", precision=a.precision, model_tf32=a.tf32, batched_backward=not a.loop_backward, conv_engine=not a.no_engine)
X = torch.randn(a.batch, 3, 224 if a.model == "vit_b16" else 32, 224 if a.model == "vit_b16" else 32, device=dev); y = torch.randint(10, (a.batch,), device=dev)
rec = collections.defaultdict(list)

def wrap(obj, name, label=None):
    f = getattr(obj, name)
    def g(*args, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = f(*args, **kw); e.record()
        rec[label or name].append((s, e)); return r
    setattr(obj, name, g)

for n in ("pack_rows", "pack_conv", "pack_nchw", "gemm_nt", "pack_conv_rows", "pack_nchw_rows", "pack_cast", "col2im", "conv_nhwc", "gemm_tn", "scale_channels", "relu_bwd", "maxpool2d_bwd", "syrk_conv_patches", "col2im_nhwc", "pack_cast_fused", "conv_bwd_strided", "diag_conv_sq", "maxpool2d_bwd_pack"):
    wrap(K, n)
wrap(be, "_forward"); wrap(be, "_backward")
for i in range(3):
    be.kron(X, y, N=50000)
torch.cuda.synchronize(); rec.clear()
t0 = time.perf_counter()
s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
R = 5
for i in range(R):
    be.kron(X, y, N=50000)
e0.record(); torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / R * 1e3
print(be.last_backward_mode)
print(f"batch {a.batch} precision {a.precision} tf32={a.tf32}: {s0.elapsed_time(e0)/R:.2f} ms/step device, {wall:.2f} ms wall")
tot = 0
for k, v in rec.items():
    ms = sum(s.elapsed_time(e) for s, e in v) / R
    tot += ms
    print(f"  {k:12s} {ms:8.3f} ms/step  ({len(v)//R} calls)")
print(f"  accounted   {tot:8.3f} ms/step")
