"""The C-ABI library loads on a GPU-less machine and exports every symbol declared in include/laplace_b200.h
(no compute calls here)."""
import ctypes
import os
import re

from laplace_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "laplace_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lpb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header_symbols():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"missing export {n}"
    assert sorted(_lib.EXPORTS) == names
    assert lib.lpb_version() >= 100
    assert isinstance(lib.lpb_last_error(), bytes)


def test_argument_validation_without_gpu():
    """Host-side argument checks fail with a message before any launch."""
    lib = _lib.load()
    rc = lib.lpb_gemm_nt_f32(None, 1, None, 1, 4, 4, 8, ctypes.c_float(1.0), 1, None, 4, 0, None)
    assert rc != 0 and b"leading dimension" in lib.lpb_last_error()


def test_product_path_refuses_cpu_tensors():
    import pytest
    import torch

    from laplace_b200 import B200GGN

    model = torch.nn.Sequential(torch.nn.Linear(3, 2))
    with pytest.raises(RuntimeError):
        B200GGN(model, "classification").kron(torch.randn(4, 3), torch.tensor([0, 1, 0, 1]), N=4)


def test_live_tap_rule_matches_python_wrapper():
    """``lpb_conv_live_taps`` (host-only helper): kernel positions whose window overlaps the image at all -- the rule the
    Python wrapper of the im2col-free factor uses to size its scratch buffer."""
    lib = _lib.load()
    for kh, kw, ph, pw, H, W in [(3, 3, 1, 1, 1, 1), (3, 3, 1, 1, 2, 2), (3, 3, 1, 1, 8, 8), (1, 1, 0, 0, 1, 1), (3, 1, 1, 0, 1, 4),
                                 (3, 3, 1, 1, 1, 5)]:
        want = sum(1 for a in range(kh) for b in range(kw) if abs(a - ph) < H and abs(b - pw) < W)
        assert lib.lpb_conv_live_taps(kh, kw, ph, pw, H, W) == want
    assert lib.lpb_conv_live_taps(3, 3, 1, 1, 1, 1) == 1 and lib.lpb_conv_live_taps(3, 3, 1, 1, 2, 2) == 9
