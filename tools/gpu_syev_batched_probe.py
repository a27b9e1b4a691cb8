"""Probe: cusolverDnXsyevBatched (cuSOLVER >= 11.7) on groups of equally sized symmetric matrices vs a loop of
torch.linalg.eigh -- does the batched library call overlap the factors of one size group?"""
import ctypes as C, os, sys, time

import torch


def load_cusolver():
    torch.linalg.eigh(torch.eye(4, device="cuda"))      # makes torch load its cusolver
    for name in ("libcusolver.so.11", "libcusolver.so"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    raise RuntimeError("no cusolver")


class Syev:
    def __init__(self):
        self.lib = load_cusolver()
        self.h = C.c_void_p()
        assert self.lib.cusolverDnCreate(C.byref(self.h)) == 0
        self.params = C.c_void_p()
        assert self.lib.cusolverDnCreateParams(C.byref(self.params)) == 0

    def batched(self, A):            # A [b, n, n] fp32 contiguous (symmetric): overwritten with eigenvectors (column-major)
        b, n, _ = A.shape
        W = torch.empty(b, n, device=A.device, dtype=torch.float32)
        info = torch.zeros(b, device=A.device, dtype=torch.int32)
        self.lib.cusolverDnSetStream(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        wd, wh = C.c_size_t(), C.c_size_t()
        args = (self.h, self.params, C.c_int(1), C.c_int(1), C.c_int64(n), C.c_int(0), C.c_void_p(A.data_ptr()), C.c_int64(n), C.c_int(0),
                C.c_void_p(W.data_ptr()), C.c_int(0))
        rc = self.lib.cusolverDnXsyevBatched_bufferSize(*args, C.byref(wd), C.byref(wh), C.c_int64(b))
        assert rc == 0, rc
        dbuf = torch.empty(max(1, wd.value), device=A.device, dtype=torch.uint8)
        hbuf = (C.c_char * max(1, wh.value))()
        rc = self.lib.cusolverDnXsyevBatched(*args, C.c_void_p(dbuf.data_ptr()), C.c_size_t(wd.value), hbuf, C.c_size_t(wh.value),
                                             C.c_void_p(info.data_ptr()), C.c_int64(b))
        assert rc == 0, rc
        return W, A, info


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
    s = Syev()
    g = torch.Generator(device="cuda").manual_seed(0)
    for n, b in ((64, 6), (128, 6), (256, 6), (513, 17), (576, 5), (1152, 4), (2304, 3)):
        X = torch.randn(b, n, 2 * n, device="cuda", generator=g)
        H = X @ X.transpose(1, 2) / (2 * n)
        loop = wall(lambda: [torch.linalg.eigh(h) for h in H])
        bat = wall(lambda: s.batched(H.clone()))
        W, V, info = s.batched(H.clone())
        Vc = V.transpose(1, 2)          # column-major eigenvectors -> rows = vectors^T
        resid = float((H @ Vc - Vc * W.unsqueeze(1)).norm() / H.norm())
        orth = float((Vc.transpose(1, 2) @ Vc - torch.eye(n, device="cuda")).norm())
        print(f"n={n} b={b}: eigh loop {loop:.1f} ms, XsyevBatched {bat:.1f} ms, info {info.tolist()[:3]}, resid {resid:.1e}, orth {orth:.1e}", flush=True)


if __name__ == "__main__":
    main()
