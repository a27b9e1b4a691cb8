#!/usr/bin/env python
"""Benchmark of the curvature hot path (BASELINE.json metric: KFAC-GGN fit() samples/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
                    [--precision bf16x3|bf16|fp32|auto] [--model resnet18|mlp|...]

Workload (N=1 default, configs[1] of BASELINE.json): ResNet-18-shaped CNN (torchvision topology, random
init, frozen norm affines), 32x32x3 synthetic inputs, C=10, KFAC-GGN over all conv/linear weights.
A "step" = one pass of the hot path over one batch of B samples: forward + C batched reverse passes,
packing, the A/B factor SYRKs on the tensor cores and the accumulation into the running factor buffers
(i.e. one iteration of the ``fit()`` loop, baselaplace.py:969-985).  Multi-GPU: every rank runs the same
per-batch loop on its own shard (weak scaling) and the factors are all-reduced ONCE after the K steps
(inside the timed region).

One JSON line is printed by rank 0 (see the task contract): `value` = samples/s with inputs resident in
HBM, `e2e` = the same through the public API with pinned-host batches copied inside the timed region,
`roofline` = the tcgen05 SYRK kernel against the measured bf16 peak, `cpu_baseline` = the CPU oracle port of
the reference's KFAC path timed on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "kfac_ggn_fit_samples_per_sec"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16", "bf16x3"])
    ap.add_argument("--model", default="resnet18")
    ap.add_argument("--structure", default="kron", choices=["kron", "diag_ef"],
                    help="kron: KFAC-GGN (the headline metric, BASELINE configs[1]); diag_ef: diagonal empirical Fisher with a "
                         "vector all-reduce (BASELINE configs[3], e.g. --model wrn28_10)")
    ap.add_argument("--cpu-samples", type=int, default=256, help="samples of the bounded CPU-baseline run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-predictive", action="store_true", help="skip the GLM-predictive legs")
    ap.add_argument("--no-cuda-graph", action="store_true", help="run every kron() step eagerly (default: the step is captured "
                    "into a CUDA graph after two eager calls and replayed, backend_kwargs={'cuda_graph': True})")
    ap.add_argument("--model-tf32", action="store_true", help="let cuDNN/cuBLAS use TF32 in the model's own passes")
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU arm (0 = min(cores, 16), the measured optimum)")
    return ap.parse_args()


def workload_name(args):
    what = "KFAC-GGN all weights" if args.structure == "kron" else "diagonal empirical Fisher all weights"
    return f"{args.model} (random init, frozen norm), {'x'.join(map(str, input_shape(args.model)[::-1]))} synthetic, C=10, {what}"


def make_model(name):
    from laplace_b200 import models

    return models.make(name)


def input_shape(name):
    return (784,) if name == "mlp" else ((3, 224, 224) if name == "vit_b16" else (3, 32, 32))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(args, steps, warmup, samples_per_step):
    """The reference's CPU implementation of the path: curvlinops is not installable offline, so this times the
    oracle PORT of CurvlinopsGGN.kron (oracle/curvature_oracle.py:kfac_factors) with all host threads."""
    from oracle import curvature_oracle as co

    torch.set_num_threads(cpu_threads(args))
    model = make_model(args.model)
    torch.manual_seed(1)
    X = torch.randn(samples_per_step, *input_shape(args.model))
    y = torch.randint(10, (samples_per_step,))
    acc = None
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, kf = co.kfac_factors(model, "classification", X, y, N=50000)
        acc = kf if acc is None else [[a.add_(b) for a, b in zip(Fa, Fb)] for Fa, Fb in zip(acc, kf)]
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return steps * samples_per_step / total, total / steps * 1e3


def cpu_threads(args):
    """Threads of the CPU arm: measured on the 128-core GPU host the oracle port peaks at 16 torch threads
    (8: 240, 16: 293, 32: 193, 64: 85 samples/s) -- more threads make the many small ops of the per-class
    reverse passes slower; the count actually used is reported as `cores`."""
    return args.cpu_threads if args.cpu_threads > 0 else min(os.cpu_count() or 1, 16)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spp = 128
    steps, warmup = max(1, min(args.steps, 4)), max(1, min(args.warmup, 1))
    v, ms = cpu_reference_run(args, steps, warmup, spp)
    cores = cpu_threads(args)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "batch_per_step": spp, "note": "CPU oracle port of CurvlinopsGGN.kron "
                   "(curvlinops itself is not installable offline); bounded sample per step"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{steps} steps x {spp} samples"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_burst": p.get("bf16_tflops"), "bf16_sustained": p.get("bf16_tflops_sustained"), "hbm": p.get("hbm_gbs"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def run_ours(args):
    import torch.distributed as dist

    from laplace_b200 import B200GGN
    from laplace_b200 import kernels as K
    from laplace_b200 import matrix
    from laplace_b200.data import PrefetchLoader
    from laplace_b200.distributed import allreduce_curvature, decompose_sharded
    from laplace_b200.interface import HAVE_REFERENCE
    from laplace_b200.posterior import B200Laplace

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, Ksteps, W = args.batch, args.steps, max(3, args.warmup)
    N_total = 50000
    model = make_model(args.model).to(dev)
    be_kwargs = {"precision": args.precision, "model_tf32": args.model_tf32, "cuda_graph": not args.no_cuda_graph}
    be = B200GGN(model, "classification", **be_kwargs)
    shape = input_shape(args.model)
    torch.manual_seed(1 + rank)
    n_batches = min(8, W + Ksteps)  # pool of distinct batches; factor buffers (376 MB) dwarf the 126 MB L2 anyway
    Xs = [torch.randn(B, *shape, device=dev) for _ in range(n_batches)]
    ys = [torch.randint(10, (B,), device=dev) for _ in range(n_batches)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn):
        """CUDA-event time of fn() on the current stream, barrier + synchronize on both sides, max over ranks."""
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = fn()
        e.record()
        barrier()
        return max_over_ranks(s.elapsed_time(e)), out

    # ---------------- device-resident run ----------------
    H = None

    def step(i):
        nonlocal H
        _, kr = be.kron(Xs[i % n_batches], ys[i % n_batches], N=N_total)
        if H is None:        # own accumulator: with cuda_graph the returned factors are the graph's static output buffers
            from laplace_b200 import B200Kron

            H = B200Kron.zeros(kr.dims(), dev, torch.float32)
        H += kr

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for i in range(W):
        step(i)
    # the exchange is part of the timed region: warm the very collective it uses (NCCL builds its channels / registers
    # the buffer on the first call of a given size -- 75 ms for the 376 MB factor buffer when it was left cold)
    scratch = None
    if world > 1:
        scratch = torch.zeros_like(H._flat)
        for _ in range(2):
            dist.all_reduce(scratch)
    barrier()
    l0 = K.LAUNCHES

    def timed_steps():
        for i in range(Ksteps):
            step(W + i)
        allreduce_curvature(H)

    ms_total, _ = timed(timed_steps)
    launches = K.LAUNCHES - l0
    clk = clocks.stop() if rank == 0 else None
    value = world * Ksteps * B / (ms_total / 1e3)
    exchange = {"call": "none (single process)", "bytes": 0, "ms": 0.0}
    if world > 1:
        ms_coll, _ = timed(lambda: dist.all_reduce(scratch))
        exchange = {"call": "ncclAllReduce(sum, fp32) of the flat Kronecker-factor buffer, once after the K steps, inside the "
                            "timed region", "bytes": int(scratch.numel() * 4), "ms": round(ms_coll, 3),
                    "busbw_GBps": round(2 * (world - 1) / world * scratch.numel() * 4 / (ms_coll / 1e3) / 1e9, 1)}
        del scratch

    extras = {}
    # ---------------- end to end through the public API with pinned host batches (right after the device-resident
    # ---------------- run: same allocator / clock state) ----------------
    Xh = [x.cpu().pin_memory() for x in Xs]
    yh = [y.cpu().pin_memory() for y in ys]

    class HostSet(torch.utils.data.Dataset):
        def __len__(self):
            return N_total

    class HostLoader:
        dataset = HostSet()

        def __init__(self, n, off):
            self.n, self.off = n, off

        def __len__(self):
            return self.n

        def __iter__(self):
            for i in range(self.n):
                j = (self.off + i) % n_batches
                yield Xh[j], yh[j]

    if HAVE_REFERENCE:
        # the UNMODIFIED reference front end: Laplace(...) constructs the backend (baselaplace.py:179-194); the per-batch
        # loop is ParametricLaplace.fit (baselaplace.py:904-987) -- KronLaplace.fit adds only the decomposition, which is
        # timed above -- fed by a loader that stages batch i+1 on a copy stream (laplace_b200/data.py)
        import laplace
        from laplace.baselaplace import ParametricLaplace

        la = laplace.Laplace(model, "classification", "all", "kron", backend=B200GGN, backend_kwargs=be_kwargs)
        api = "laplace.Laplace(model, 'classification', 'all', 'kron', backend=B200GGN) [unmodified reference front end, " \
              "baseline/_ref]; timed: ParametricLaplace.fit(la, PrefetchLoader(host batches)) + all-reduce + loss read"

        def run_fit(n, off):
            ParametricLaplace.fit(la, PrefetchLoader(HostLoader(n, off), dev))
            return la.H
    else:
        la = B200Laplace(model, "classification", "all", "kron", backend=B200GGN, backend_kwargs=be_kwargs)
        api = "laplace_b200.posterior.B200Laplace.fit (host-side mirror: the reference package is not importable here)"

        def run_fit(n, off):
            la.fit(HostLoader(n, off), decompose=False)
            return la.H_facs

    run_fit(W, 0)
    state = {}

    def e2e_steps():
        Hk = run_fit(Ksteps, W)
        allreduce_curvature(Hk)
        state["loss"] = float(la.loss)  # device -> host read of the step result

    ms_e2e, _ = timed(e2e_steps)
    e2e_value = world * Ksteps * B / (ms_e2e / 1e3)
    h2d = Xh[0].numel() * 4 + yh[0].numel() * 8
    # what this box's host link delivers for exactly these copies (the e2e leg is copy-bound when a step's input takes longer
    # to arrive than to process: 50 MB per 20 ms step needs 2.5 GB/s)
    xdst = torch.empty_like(Xs[0])

    def h2d_copies():
        for j in range(8):
            xdst.copy_(Xh[j % n_batches], non_blocking=True)

    h2d_copies()
    ms_h2d, _ = timed(h2d_copies)
    h2d_gbps = 8 * Xh[0].numel() * 4 / (ms_h2d / 1e3) / 1e9
    del xdst

    # ---------------- once-per-fit eigendecomposition (inside KronLaplace.fit, baselaplace.py:1809), sharded over ranks ----
    def decompose_once():
        return decompose_sharded(H) if world > 1 else H.decompose()

    t0 = time.perf_counter()
    ms_cold, Hd = timed(decompose_once)
    del Hd
    ms_dec, Hd = timed(decompose_once)
    extras["decompose_ms_first_call_cold"] = round(ms_cold, 1)
    extras["decompose_ms_once_per_fit"] = round(ms_dec, 1)
    extras["decompose_how"] = (f"factors sharded over {world} ranks (greedy n^3 balance), local eigh, one all-gather of Q/lambda"
                               if world > 1 else "single process") + \
        f"; n<=128 hand-written Jacobi kernel; live blocks of 129..513 rows: one batched library call (cusolverDnXsyevBatched) " \
        f"per size class {'on' if matrix.BATCHED_MID_SIZES else 'off'}; larger: library syevd from {matrix.N_EIGH_THREADS} host " \
        f"threads; dead-coordinate compaction {'on' if matrix.COMPACT_DEAD_COORDINATES else 'off'}; library start-up overlapped with the data pass"
    fit_s = N_total / value + ms_dec / 1e3
    extras["fit_50k_samples_per_sec_incl_decompose"] = round(N_total / fit_s, 1)
    del Hd

    # ---------------- roofline of the dominant kernel (separate short pass, CUDA events around each launch) -----
    # per-launch CUDA events need eager launches: same backend settings without the graph replay
    roof = measure_roofline(B200GGN(model, "classification", **{**be_kwargs, "cuda_graph": False}), K, Xs, ys, N_total, args, dev)

    # The legs below annotate the line (predictive, small-batch line, the contraction kernel alone); a failure in one of
    # them must not lose the headline measurement, so each is recorded as an error string instead.
    def leg(name, fn):
        try:
            fn()
        except Exception as e:  # noqa: BLE001 -- reported, not swallowed
            extras[name + "_error"] = f"{type(e).__name__}: {str(e)[:300]}"

    pred = {}
    if not args.no_predictive:
        leg("predictive", lambda: pred.update(measure_predictive(model, dev, world, timed, args)))
    if rank == 0 and world == 1 and args.model == "resnet18":
        def b512():
            res = {"cuda_graph" if be.cuda_graph else "eager": measure_small_batch(be, dev, shape, N_total)}
            other = B200GGN(model, "classification", **{**be_kwargs, "cuda_graph": not be.cuda_graph})
            res["cuda_graph" if other.cuda_graph else "eager"] = measure_small_batch(other, dev, shape, N_total)
            extras["batch_512"] = res

        leg("b512", b512)
    if rank == 0 and world == 1:
        leg("jtj_syrk_kernel", lambda: extras.__setitem__("jtj_syrk_kernel", measure_syrk_probe(K, dev)))

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, ms = cpu_reference_run(args, 2, 1, 128)
            cpu = {"value": v, "unit": UNIT, "cores": cpu_threads(args), "kind": "port",
                   "sample": "2 steps x 128 samples of the same workload (oracle port of CurvlinopsGGN.kron, "
                             f"{cpu_threads(args)} torch threads of {os.cpu_count()} host cores)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": Ksteps, "warmup": W,
            "ms_per_step": ms_total / Ksteps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16x3": "bf16 (hi/lo split, 3 products, fp32 accumulate)", "bf16": "bf16", "fp32": "f32",
                      "auto": "fp16/bf16 hi/lo split operands, fp32 accumulate (3 tensor-core products; input factors summed over "
                              ">= max(16384, 128 x d_in) rows: 1 fp16 product)"}[args.precision],
            "data": "synthetic",
            "config": {"workload": workload_name(args), "batch_per_gpu": B, "global_batch": B * world, "N_dataset": N_total,
                       "parallelism": f"dp{world}", "precision": args.precision,
                       "cuda_graph": "kron() step captured after 2 eager calls and replayed (backend_kwargs cuda_graph=True)"
                       if not args.no_cuda_graph else "off (eager launches)",
                       "model_passes": "tf32 (PyTorch default)" if args.model_tf32 else "fp32 (TF32 disabled in cuDNN/cuBLAS)",
                       "l2": "distinct batch every step; per-step working set (factor buffers 376 MB + staging) exceeds the 126 MB L2",
                       "exchange": exchange, "e2e_api": api, **extras},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / Ksteps, "loss": state.get("loss"),
                    "h2d_GBps_measured": round(h2d_gbps, 1)},
            "gpu_launches": launches,
            "roofline": roof,
            "predictive": pred or None,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_small_batch(be, dev, shape, N_total, B=512, steps=10):
    """SURVEY 8(d) quotes cfg2 at B = 512: the same step at that batch size (host-bound: ~1500 launches per step)."""
    torch.manual_seed(11)
    Xs = [torch.randn(B, *shape, device=dev) for _ in range(4)]
    ys = [torch.randint(10, (B,), device=dev) for _ in range(4)]
    from laplace_b200 import B200Kron

    H = None
    for i in range(4):
        _, k = be.kron(Xs[i % 4], ys[i % 4], N=N_total)
        if H is None:
            H = B200Kron.zeros(k.dims(), dev, torch.float32)
        H += k
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        _, k = be.kron(Xs[i % 4], ys[i % 4], N=N_total)
        H += k
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    return {"ms_per_step": round(ms, 3), "samples_per_sec": round(B / (ms / 1e3), 1)}


def measure_roofline(be, K, Xs, ys, N_total, args, dev, step_fn=None):
    """Times every native kernel family of two steps with CUDA events on the launching stream and reports the
    roofline of the DOMINANT one (largest share of device time), plus a per-family table under `families`.

    Algorithmic work per launch (dense conventions of SURVEY 8(d), stated in DESIGN.md section 3):
      gemm_nt (factor SYRKs / engine GEMMs)  2*M*N*K FLOPs (SYRK counted dense: 2*d^2*K_rows)
      conv_nhwc (implicit-GEMM convolution)  2*rows*N*K*taps FLOPs
      pack_* / col2im                        bytes read + written once
    """
    rec = {}

    def wrap(name, work):
        orig = getattr(K, name)

        def timed(*a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig(*a, **kw)
            e.record()
            rec.setdefault(name, []).append((work(r, a, kw), s, e))
            return r

        setattr(K, name, timed)
        return orig

    def bytes_packed(p):
        per = {K.F32: 4, K.BF16: 2, K.BF16X3: 4, K.F16X3: 4}[p.kind]
        return p.rows * p.K * (4 + per)

    works = {
        "gemm_nt": lambda r, a, kw: ("flops", 2.0 * a[0].rows * a[1].rows * a[0].K, a[0].kind != K.F32),
        "gemm_tn": lambda r, a, kw: ("flops", 2.0 * a[0].K * a[1].K * a[0].rows, True),
        "conv_nhwc": lambda r, a, kw: ("flops", 2.0 * a[0].rows * a[5] * a[0].K * a[6] * a[7], True),
        # implicit patch SYRK: dense 2 * d^2 * rows convention, d = C_in * kh * kw
        "syrk_conv_patches": lambda r, a, kw: ("flops", 2.0 * float(a[5].shape[0]) ** 2 * a[0].rows, True),
        "conv_bwd_strided": lambda r, a, kw: ("flops", 2.0 * a[0].rows * a[0].K * a[4].rows, True),
        "col2im_nhwc": lambda r, a, kw: ("bytes", 4.0 * (a[0].shape[0] * a[0].shape[1] + r.numel()), False),
        "pack_rows": lambda r, a, kw: ("bytes", bytes_packed(r), False),
        "pack_conv": lambda r, a, kw: ("bytes", bytes_packed(r[0]), False),
        "pack_nchw": lambda r, a, kw: ("bytes", bytes_packed(r), False),
        "pack_conv_rows": lambda r, a, kw: ("bytes", bytes_packed(r), False),
        "pack_nchw_rows": lambda r, a, kw: ("bytes", bytes_packed(r), False),
        "pack_cast": lambda r, a, kw: ("bytes", bytes_packed(r), False),
        "pack_cast_fused": lambda r, a, kw: ("bytes", bytes_packed(r) + (4.0 * a[3].numel() if len(a) > 3 and a[3] is not None else 0.0), False),
        "scale_channels": lambda r, a, kw: ("bytes", 8.0 * r.numel(), False),
        "relu_bwd": lambda r, a, kw: ("bytes", 8.0 * r.numel() + 4.0 * a[1].numel(), False),
        "maxpool2d_bwd": lambda r, a, kw: ("bytes", 4.0 * (r.numel() + a[0].numel()) + 8.0 * a[1].numel(), False),
        "col2im": lambda r, a, kw: ("bytes", 4.0 * (a[0].shape[0] * a[0].shape[1] + r.numel()), False),
        "maxpool2d_bwd_pack": lambda r, a, kw: ("bytes", bytes_packed(r) - 4.0 * r.rows * r.K + 4.0 * a[0].numel() + 8.0 * a[1].numel()
                                                + (4.0 * a[7].numel() if a[7] is not None else 0.0), False),
        # tensor-core diagonal of a convolution weight: per (sample, column) one T-deep [C_out x 9 C_in] product
        "diag_conv_sq": lambda r, a, kw: ("flops", 2.0 * a[0].rows * a[0].K * a[1].K * a[5].kernel_size[0] * a[5].kernel_size[1], True),
        "shared_weight_contract": lambda r, a, kw: ("flops", 2.0 * a[3] * a[4] * a[5] * a[6] * a[7], False),
    }
    def one(i):
        if step_fn is not None:
            step_fn(i)
        else:
            be.kron(Xs[i % len(Xs)], ys[i % len(ys)], N=N_total)

    for i in range(2):      # untimed: this (eager) backend's buffers come out of the caching allocator, not cudaMalloc
        one(i)
    torch.cuda.synchronize()
    origs = {n: wrap(n, w) for n, w in works.items()}
    try:
        for i in range(2):
            if step_fn is not None:
                step_fn(i)
            else:
                be.kron(Xs[i % len(Xs)], ys[i % len(ys)], N=N_total)
        torch.cuda.synchronize()
    finally:
        for n, o in origs.items():
            setattr(K, n, o)
    pk = peaks()
    fam = {}
    for name, lst in rec.items():
        ms = sum(s.elapsed_time(e) for _, s, e in lst) / 2
        unit = lst[0][0][0]
        work = sum(w[1] for w, _, _ in lst) / 2
        fam[name] = {"ms_per_step": ms, "launches_per_step": len(lst) // 2, "kind": unit,
                     "achieved": (work / (ms / 1e3) / 1e12) if unit == "flops" else (work / (ms / 1e3) / 1e9),
                     "unit": "TFLOP/s" if unit == "flops" else "GB/s"}
    if not fam:
        return None
    dom = max(fam, key=lambda n: fam[n]["ms_per_step"])
    d = fam[dom]
    tensor = d["kind"] == "flops"
    peak = (pk["bf16_sustained"] or 1400.0) if tensor else (pk["hbm"] or 6650.0)
    names = {"gemm_nt": "tc::gemm_tc_persistent_kernel<.,0> (tcgen05 GEMM-NT/SYRK, K-major operands: KFAC factor contractions + explicit-engine GEMMs)",
             "gemm_tn": "tc::gemm_tc_persistent_kernel<.,1> (tcgen05 SYRK on row-major operands, MN-major descriptors: KFAC factor contractions)",
             "syrk_conv_patches": "tc::gemm_tc_persistent_kernel<.,2> (tcgen05 SYRK on implicit convolution patches, shifted 4-D TMA boxes: KFAC input factors)",
             "conv_nhwc": "tc::conv_nhwc_tc_persistent_kernel (tcgen05 implicit-GEMM convolution, forward + backward-data)"}
    # DRAM bytes per launch of the dominant family from the committed ncu launch list of the same workload (a number
    # taken under a profiler is never a bench value; it only annotates the roofline entry)
    traffic = None
    tfile = {"auto": "r02_ncu_traffic.json", "bf16x3": "r01_ncu_traffic_run26.json"}.get(args.precision, "")
    tpath = os.path.join(ROOT, "profiles", tfile)
    if args.model == "resnet18" and args.batch == 4096 and tfile and os.path.exists(tpath) and step_fn is None:
        traffic = json.load(open(tpath))["families"].get(dom, {}).get("dram_bytes_per_launch")
    return {"bound": "tensor" if tensor else "hbm", "kernel": names.get(dom, "lpb::" + dom + "_kernel"),
            "achieved": d["achieved"], "peak": peak, "unit": d["unit"], "frac": d["achieved"] / peak, "traffic": traffic,
            "traffic_note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu launch list of the same workload "
                            f"(profiles/{tfile}, profiles/r02_launches_b4096_auto.md)" if traffic else "",
            "peak_source": pk["source"] + (", bf16 sustained" if tensor else ", copy bandwidth"),
            "note": "algorithmic FLOPs (one product per MAC) over CUDA-event time; precision bf16x3 issues 3 tensor-core "
                    "products per algorithmic MAC, so the tensor pipe does 3x the counted work" if tensor and args.precision == "bf16x3" else "",
            "launches_per_step": d["launches_per_step"], "ms_per_step_in_kernel": d["ms_per_step"],
            "families": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in fam.items()}}


def measure_syrk_probe(K, dev):
    """The J^T J contraction kernel alone (tensor-pipe fraction of the factor SYRK): d = 4608
    (ResNet-18 layer-4 input factor), 65 536 sample rows, row-major operands, timed with CUDA events after warm-up.
    `tensor_tflops` counts issued products (3 per MAC in the bf16x3 mode), `algorithmic_tflops` one per MAC of the
    upper triangle."""
    torch.manual_seed(7)
    d, rows = 4608, 65536
    X = torch.randn(rows, d, device=dev)
    out = {}
    pk = peaks()
    peak = pk["bf16_sustained"] or 1400.0
    for kind, name, nprod in ((K.BF16X3, "bf16x3", 3), (K.BF16, "bf16", 1)):
        P = K.pack_cast(X, kind)
        H = torch.zeros(d, d, device=dev)
        for _ in range(2):
            K.gemm_tn(P, P, H, 1.0, True, symmetric=True)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            K.gemm_tn(P, P, H, 1.0, True, symmetric=True)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 5
        alg = (d * (d + 128) / 2) * 2.0 * rows / (ms / 1e3) / 1e12     # tiles on/above the diagonal
        out[name] = {"ms": round(ms, 3), "algorithmic_tflops": round(alg, 1), "tensor_tflops": round(alg * nprod, 1),
                     "tensor_frac_of_peak": round(alg * nprod / peak, 3)}
        del P, H
    out["shape"] = f"SYRK d={d}, K={rows} sample rows (working set 1.2 GB > L2), peak {peak:.0f} TFLOP/s ({pk['source']})"
    return out


def measure_predictive(model, dev, world, timed, args):
    """GLM-predictive samples/s (the metric's second half).  Every rank predicts its own shard of test points, no
    collective; the aggregate is world x points / max-over-ranks time.

    * BASELINE configs[2]: ResNet-18 shape, last-layer full GGN posterior (P = 5130), probit predictive of 10 000 test
      points per GPU in batches of 1000, after a fit on 2048 points;
    * BASELINE configs[0]: MLP 784-128-10, all-weights Kron posterior, 1000 points (fit incl. decomposition timed too).

    Through the unmodified reference front end when it is importable (``la(x, pred_type="glm", link_approx="probit")``),
    through the host-side mirror otherwise -- `api` says which."""
    from laplace_b200 import B200GGN, models
    from laplace_b200.interface import HAVE_REFERENCE
    from laplace_b200.posterior import B200Laplace

    TD, DL = torch.utils.data.TensorDataset, torch.utils.data.DataLoader
    if HAVE_REFERENCE:
        import laplace

        def make(m, sub, hs):
            return laplace.Laplace(m, "classification", sub, hs, backend=B200GGN)

        def predict(la, x):
            return la(x, pred_type="glm", link_approx="probit")
        api = "laplace.Laplace(..., backend=B200GGN): la.fit(loader); la(x, pred_type='glm', link_approx='probit')"
    else:
        def make(m, sub, hs):
            return B200Laplace(m, "classification", sub, hs, backend=B200GGN)

        def predict(la, x):
            return la(x)
        api = "laplace_b200.posterior.B200Laplace (mirror; reference not importable)"
    out = {"api": api}
    if args.model == "resnet18":
        torch.manual_seed(5)
        shape = input_shape(args.model)
        Xf, yf = torch.randn(2048, *shape, device=dev), torch.randint(10, (2048,), device=dev)
        n_test, bs = 10000, 1000
        Xt = torch.randn(n_test, *shape, device=dev)
        la = make(model, "last_layer", "full")
        la.fit(DL(TD(Xf, yf), batch_size=512))
        ms_llfit, _ = timed(lambda: la.fit(DL(TD(Xf, yf), batch_size=512)))   # structured last-layer GGN, 4 batches of 512
        predict(la, Xt[:bs])                       # warm-up: posterior covariance, gathered blocks, allocator
        predict(la, Xt[:bs])

        def run():
            acc = 0.0
            for i in range(0, n_test, bs):
                acc = acc + predict(la, Xt[i:i + bs]).sum()
            return float(acc)                      # device -> host read of the result

        ms, _ = timed(run)
        rate = world * n_test / (ms / 1e3)
        D, C = 512, 10
        flops = 2.0 * (D + 1) * (C * C * (D + 1)) + 2.0 * C * C * (D + 1)   # [phi;1] against the gathered blocks + pair dots
        peak = 148 * 128 * 2 * 1.965e9 / 1e12
        out["ll_full_resnet18"] = {
            "samples_per_sec": round(rate, 1), "test_points_per_gpu": n_test, "batch": bs, "ms_total": round(ms, 2),
            "fit_samples_per_sec": round(world * 2048 / (ms_llfit / 1e3), 1),
            "includes": "backbone forward (convolution engine) + structured J Sigma J^T + probit link, result read back",
            "roofline": {"bound": "fp32 SIMT FMA (gemm_nt_f32: [phi;1] x gathered covariance blocks)",
                         "algorithmic_flops_per_sample": flops, "achieved": round(flops * rate / world / 1e12, 2),
                         "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(flops * rate / world / 1e12 / peak, 3),
                         "peak_source": "nominal: 148 SMs x 128 fp32 lanes x 2 x 1.965 GHz (MEASURED_PEAKS.json has no fp32 entry)",
                         "dense_equivalent_flops_per_sample": 2.0 * C * 5130 ** 2 + 2.0 * C * C * 5130}}
        del la, Xt
    # BASELINE configs[0]
    mlp = models.make("mlp").to(dev)
    torch.manual_seed(6)
    Xm, ym = torch.randn(1000, 784, device=dev), torch.randint(10, (1000,), device=dev)
    lam = make(mlp, "all", "kron")
    loader = DL(TD(Xm, ym), batch_size=128)
    lam.fit(loader)
    predict(lam, Xm[:500])
    ms_fit, _ = timed(lambda: lam.fit(loader))

    def run_mlp():
        acc = 0.0
        for i in range(0, 1000, 500):
            acc = acc + predict(lam, Xm[i:i + 500]).sum()
        return float(acc)

    ms_pred, _ = timed(run_mlp)
    out["mlp_kron"] = {"fit_1k_samples_per_sec_incl_decompose": round(world * 1000 / (ms_fit / 1e3), 1),
                       "glm_predictive_samples_per_sec": round(world * 1000 / (ms_pred / 1e3), 1), "batch": 500}
    return out


def run_diag_ef(args):
    """BASELINE configs[3]: diagonal empirical Fisher over all weights (e.g. ``--model wrn28_10``), one vector all-reduce at
    the end.  Same timing contract as the headline run; ``value`` = samples/s of ``B200EF.diag`` steps."""
    import torch.distributed as dist

    from laplace_b200 import B200EF
    from laplace_b200 import kernels as K

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, Ksteps, W = args.batch, args.steps, max(3, args.warmup)
    N_total = 50000
    model = make_model(args.model).to(dev)
    be = B200EF(model, "classification", precision=args.precision, model_tf32=args.model_tf32)
    shape = input_shape(args.model)
    torch.manual_seed(1 + rank)
    nb = min(4, W + Ksteps)
    Xs = [torch.randn(B, *shape, device=dev) for _ in range(nb)]
    ys = [torch.randint(10, (B,), device=dev) for _ in range(nb)]
    state = {"H": None}

    def step(i):
        _, d = be.diag(Xs[i % nb], ys[i % nb], N=N_total)
        state["H"] = d if state["H"] is None else state["H"].add_(d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for i in range(W):
        step(i)
    if world > 1:
        scratch = torch.zeros_like(state["H"])
        for _ in range(2):
            dist.all_reduce(scratch)
    barrier()
    l0 = K.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(Ksteps):
        step(W + i)
    if world > 1:
        dist.all_reduce(state["H"])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = K.LAUNCHES - l0
    clk = clocks.stop() if rank == 0 else None
    roof = measure_roofline(be, K, Xs, ys, N_total, args, dev, step_fn=step)
    if rank == 0:
        P = int(state["H"].numel())
        print(json.dumps({
            "metric": "diag_ef_fit_samples_per_sec", "value": world * Ksteps * B / (ms / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": Ksteps, "warmup": W, "ms_per_step": ms / Ksteps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 hi/lo split operands, fp32 accumulate", "data": "synthetic",
            "config": {"workload": workload_name(args), "batch_per_gpu": B, "parallelism": f"dp{world}", "parameters": P,
                       "exchange": f"one all-reduce of the {4 * P / 1e6:.0f} MB diagonal after the K steps, inside the timed region"},
            "clocks": clk, "gpu_launches": launches, "roofline": roof}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.structure == "diag_ef":
        run_diag_ef(a)
    else:
        run_ours(a)
