"""ctypes binding of ONE cuSOLVER entry point, ``cusolverDnXsyevBatched`` (cuSOLVER >= 11.7 / CUDA 12.6.2): the batched
symmetric eigensolver for groups of equally sized matrices.  LIBRARY call, declared as such in DESIGN.md -- PyTorch does
not expose it (``torch.linalg.eigh`` loops over the batch: 17 matrices of 513 rows take 96 ms in the loop, 30 ms here; 6 of
256 rows 32 vs 3.8 ms; profiles/r02_eigh.md).  The library torch already loaded is re-used (no second copy of cuSOLVER)."""
from __future__ import annotations

import ctypes as C

import torch

_state = {}


def _lib():
    if "lib" in _state:
        return _state["lib"]
    lib = None
    try:
        torch.linalg.eigh(torch.eye(2, device="cuda"))     # makes torch load (and initialise) its cuSOLVER
        for name in ("libcusolver.so.11", "libcusolver.so.12", "libcusolver.so"):
            try:
                cand = C.CDLL(name)
            except OSError:
                continue
            if hasattr(cand, "cusolverDnXsyevBatched") and hasattr(cand, "cusolverDnXsyevBatched_bufferSize"):
                lib = cand
                break
    except Exception:  # noqa: BLE001 -- optional acceleration: absence is reported by available()
        lib = None
    _state["lib"] = lib
    return lib


def available() -> bool:
    return torch.cuda.is_available() and _lib() is not None


def _handle(device_index: int):
    key = ("h", device_index)
    if key not in _state:
        lib = _lib()
        h, p = C.c_void_p(), C.c_void_p()
        if lib.cusolverDnCreate(C.byref(h)) != 0 or lib.cusolverDnCreateParams(C.byref(p)) != 0:
            raise RuntimeError("cusolverDnCreate failed")
        _state[key] = (h, p)
    return _state[key]


def syev_batched(A: torch.Tensor):
    """``A [b, n, n]`` fp32, symmetric, contiguous -- OVERWRITTEN.  Returns ``(W [b, n] ascending, Q [b, n, n] with eigenvectors
    as columns, info [b] int32 on the device)``; asynchronous on the current stream."""
    assert A.is_cuda and A.dtype == torch.float32 and A.is_contiguous() and A.dim() == 3 and A.shape[1] == A.shape[2]
    lib = _lib()
    b, n, _ = A.shape
    h, params = _handle(A.device.index or 0)
    W = torch.empty(b, n, device=A.device, dtype=torch.float32)
    info = torch.zeros(b, device=A.device, dtype=torch.int32)
    if lib.cusolverDnSetStream(h, C.c_void_p(torch.cuda.current_stream(A.device).cuda_stream)) != 0:
        raise RuntimeError("cusolverDnSetStream failed")
    wd, wh = C.c_size_t(), C.c_size_t()
    # jobz = CUSOLVER_EIG_MODE_VECTOR (1), uplo = CUBLAS_FILL_MODE_UPPER (1) of the column-major view = lower of ours: symmetric
    # input, either triangle holds the same numbers; CUDA_R_32F = 0
    head = (h, params, C.c_int(1), C.c_int(1), C.c_int64(n), C.c_int(0), C.c_void_p(A.data_ptr()), C.c_int64(n), C.c_int(0),
            C.c_void_p(W.data_ptr()), C.c_int(0))
    rc = lib.cusolverDnXsyevBatched_bufferSize(*head, C.byref(wd), C.byref(wh), C.c_int64(b))
    if rc != 0:
        raise RuntimeError(f"cusolverDnXsyevBatched_bufferSize failed ({rc})")
    dbuf = torch.empty(max(1, wd.value), device=A.device, dtype=torch.uint8)
    hbuf = (C.c_char * max(1, wh.value))()
    rc = lib.cusolverDnXsyevBatched(*head, C.c_void_p(dbuf.data_ptr()), C.c_size_t(wd.value), hbuf, C.c_size_t(wh.value),
                                    C.c_void_p(info.data_ptr()), C.c_int64(b))
    if rc != 0:
        raise RuntimeError(f"cusolverDnXsyevBatched failed ({rc})")
    # column-major eigenvectors: the memory of A now holds V^T in our row-major reading
    return W, A.transpose(1, 2), info
