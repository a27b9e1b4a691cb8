"""Where decompose() spends its time, phase by phase, and whether host threads really overlap the library eigensolver."""
import os, sys, threading, time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_b200 import B200GGN, kernels as K, matrix, models  # noqa: E402


def wall(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    for n in (1152, 2304):
        Hs = []
        for _ in range(3):
            X = torch.randn(n, 2 * n, device=dev, generator=g)
            Hs.append(X @ X.T / (2 * n))
        seq = wall(lambda: [torch.linalg.eigh(H) for H in Hs])

        def threaded():
            streams = [torch.cuda.Stream() for _ in Hs]
            cur = torch.cuda.current_stream()
            def work(H, s):
                torch.cuda.set_device(0)
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    torch.linalg.eigh(H)
            th = [threading.Thread(target=work, args=(H, s)) for H, s in zip(Hs, streams)]
            [t.start() for t in th]
            [t.join() for t in th]
            [cur.wait_stream(s) for s in streams]
        par = wall(threaded)
        print(f"3 x eigh({n}): sequential {seq:.1f} ms, 3 threads {par:.1f} ms", flush=True)
    model = models.make("resnet18").to(dev)
    torch.manual_seed(0)
    X, y = torch.randn(1024, 3, 32, 32, device=dev), torch.randint(10, (1024,), device=dev)
    _, k = B200GGN(model, "classification", precision="bf16x3").kron(X, y, N=50000)
    mats = [H for F in k.kfacs for H in F]
    print(f"live_sizes (one host read): {wall(lambda: matrix.live_sizes(mats)):.2f} ms")
    small = {}
    for H in mats:
        if H.shape[0] <= 128:
            small.setdefault(H.shape[0], []).append(H)
    for n, grp in small.items():
        st = torch.stack(grp)
        print(f"jacobi kernel {len(grp)} x {n}: {wall(lambda: K.eigh_jacobi(st)):.2f} ms")
    large = [(0, 0, H) for H in mats if H.shape[0] > 128]
    for nt in (1, 2, 4):
        matrix.N_EIGH_THREADS = nt

        def run():
            ev = [[None] for _ in large]
            vv = [[None] for _ in large]
            items = [(i, 0, H) for i, (_, _, H) in enumerate(large)]
            matrix._symeig_concurrent(items, ev, vv)
        print(f"library part ({len(large)} factors), {nt} thread(s): {wall(run, 2):.1f} ms", flush=True)
    matrix.N_EIGH_THREADS = 4
    print(f"whole decompose: {wall(lambda: k.decompose(), 3):.1f} ms")
    by = {}
    live = matrix.live_sizes(mats)
    for H, l in zip(mats, live):
        if H.shape[0] > 128:
            by.setdefault((H.shape[0], l if l != H.shape[0] else None), 0)
    print("estimated serial library ms:", round(sum(matrix.eigh_cost_ms(l) for H, l in zip(mats, live) if H.shape[0] > 128), 1))


if __name__ == "__main__":
    main()
