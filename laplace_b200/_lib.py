"""ctypes binding of ``liblaplace_b200.so`` (C ABI declared in ``include/laplace_b200.h``).

The product path has no fallback: if the shared library is missing or a call fails the
wrapper raises (``RuntimeError``), mirroring the reference's "Python exceptions only" error
convention.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblaplace_b200.so")

c_i64, c_int, c_f32, c_vp = C.c_int64, C.c_int, C.c_float, C.c_void_p

# name -> argtypes ; every entry point of include/laplace_b200.h returning int
SIGNATURES = {
    "lpb_device_info": [C.POINTER(c_int)] * 3,
    "lpb_set_gemm_tile_mode": [c_int],
    "lpb_set_mask_major_min": [c_i64],
    "lpb_pack_rows_t": [c_vp, c_i64, c_i64, c_i64, c_vp, c_int, c_f32, c_int, c_vp, c_vp, c_int, c_i64, c_i64, c_vp],
    "lpb_pack_conv2d_t": [c_vp] + [c_int] * 12 + [c_f32, c_int, c_int, c_vp, c_vp, c_int, c_i64, c_i64, c_vp],
    "lpb_pack_nchw_t": [c_vp, c_i64, c_int, c_int, c_f32, c_int, c_int, c_vp, c_vp, c_int, c_i64, c_i64, c_vp],
    "lpb_pack_conv2d_rows": [c_vp] + [c_int] * 12 + [c_vp, c_vp, c_int, c_i64, c_vp],
    "lpb_pack_nchw_rows": [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_int, c_i64, c_vp],
    "lpb_pack_cast": [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_int, c_i64, c_vp],
    "lpb_col2im": [c_vp, c_i64] + [c_int] * 12 + [c_vp, c_vp],
    "lpb_syrk_conv_patches_tc": [c_vp, c_vp, c_i64, c_i64] + [c_int] * 7 + [c_f32, c_int, c_vp, c_i64, c_int, c_vp],
    "lpb_conv_live_taps": [c_int] * 6,
    "lpb_taps_to_param_accumulate": [c_vp, c_i64] + [c_int] * 8 + [c_vp, c_i64, c_vp],
    "lpb_diag_conv_sq_tc": [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64] + [c_int] * 8 + [c_f32, c_int, c_vp, c_i64, c_vp],
    "lpb_taps_to_param_rect": [c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp, c_i64, c_vp],
    "lpb_conv_bwd_strided_tc": [c_vp, c_vp, c_i64, c_int, c_int, c_i64, c_i64, c_vp, c_vp, c_i64] + [c_int] * 9 + [c_vp, c_i64, c_vp],
    "lpb_pack_cast_fused": [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_int, c_i64, c_vp],
    "lpb_col2im_nhwc": [c_vp, c_i64] + [c_int] * 12 + [c_vp, c_vp],
    "lpb_gemm_nt_f32": [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, c_f32, c_int, c_vp, c_i64, c_int, c_vp],
    "lpb_gemm_nt_tc": [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_f32, c_int, c_vp, c_i64, c_int, c_int,
                       c_vp],
    "lpb_scale_channels": [c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_vp],
    "lpb_relu_bwd": [c_vp, c_vp, c_vp, c_i64, c_int, c_vp],
    "lpb_maxpool2d_bwd": [c_vp, c_vp, c_vp, c_i64] + [c_int] * 9 + [c_vp],
    "lpb_maxpool2d_bwd_pack_nhwc": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64] + [c_int] * 9 + [c_vp],
    "lpb_maxpool2d_bwd_nhwc": [c_vp, c_vp, c_vp, c_i64] + [c_int] * 9 + [c_vp],
    "lpb_gemm_tn_tc": [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_f32, c_int, c_vp, c_i64, c_int, c_int,
                       c_vp],
    "lpb_conv_nhwc_tc": [c_vp, c_vp, c_i64, c_int, c_int, c_i64, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int,
                         c_int, c_f32, c_vp, c_i64, c_int, c_vp],
    "lpb_shared_weight_contract": [c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp,
                                   c_i64, c_i64, c_i64, c_vp],
    "lpb_kfac_accum_rows": [c_vp, c_i64, c_i64, c_i64, c_f32, c_int, c_vp, c_i64, c_vp, c_i64, c_vp],
    "lpb_kfac_accum_conv_input": [c_vp] + [c_int] * 12 + [c_f32, c_vp, c_i64, c_vp, c_i64, c_vp],
    "lpb_kron_conv_quadform": [c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_f32, c_int, c_vp,
                               c_vp],
    "lpb_jac_linear_write": [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp],
    "lpb_ll_jacobian_write": [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp],
    "lpb_batched_pair_dot": [c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_i64, c_i64, c_i64, c_i64, c_int, c_vp,
                             c_vp],
    "lpb_ll_ggn_expand": [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp],
    "lpb_ll_sigma_gather": [c_vp, c_int, c_int, c_int, c_vp, c_vp],
    "lpb_eigh_jacobi": [c_vp, c_int, c_int, c_vp, c_vp, c_int, c_vp],
}
EXPORTS = ["lpb_version", "lpb_last_error", "lpb_workspace_bytes"] + list(SIGNATURES)

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the native library (once).  Raises ``NativeLibraryError`` when it is absent --
    there is deliberately no Python/torch fallback for the kernels."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `bash laplace_b200/csrc/build.sh` (needs nvcc, sm_100a)."
        )
    lib = C.CDLL(LIB_PATH)
    lib.lpb_version.restype = c_int
    lib.lpb_version.argtypes = []
    lib.lpb_last_error.restype = C.c_char_p
    lib.lpb_last_error.argtypes = []
    lib.lpb_workspace_bytes.restype = c_i64
    lib.lpb_workspace_bytes.argtypes = [c_i64, c_i64]
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.lpb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{name} failed: {msg}")
