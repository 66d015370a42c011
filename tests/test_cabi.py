"""CPU: the C-ABI library builds, loads, and exports every symbol include/layoutdetr_sm100.h declares
(no compute calls — there is no GPU in this container)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "layoutdetr_sm100.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(ld_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    assert "ld_gemm_bf16" in names and "ld_bias_act" in names and "ld_upfirdn2d" in names and len(names) >= 30


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 on its own (no C++ / CUDA / torch types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no gcc on this box")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "layoutdetr_sm100.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_library_builds_and_exports_all_declared_symbols():
    from layoutdetr_b200 import build
    lib_path = build.build(verbose=False)
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, "declared in the header but not exported: %s" % missing
    lib.ld_version.restype = ctypes.c_int
    assert lib.ld_version() >= 100
    lib.ld_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.ld_last_error(), bytes)


def test_argument_validation_without_gpu():
    """Bad arguments are rejected on the host before any CUDA call."""
    from layoutdetr_b200 import _lib
    lib = _lib.lib()
    rc = lib.ld_gemm_bf16(None, None)
    assert rc == -1 and b"null descriptor" in lib.ld_last_error()
    rc = lib.ld_bias_act(None, None, None, None, None, None, 0, 0, 1, ctypes.c_float(0), ctypes.c_float(1), ctypes.c_float(-1),
                         ctypes.c_int64(0), 0, ctypes.c_int64(1), None)
    assert rc == -1


def test_new_entry_points_validate_arguments_without_gpu():
    """Box-loss / eval-metric / loader entry points reject null pointers and bad sizes on the host, before any CUDA call."""
    from layoutdetr_b200 import _lib
    lib = _lib.lib()
    i64, f32 = ctypes.c_int64, ctypes.c_float
    assert lib.ld_layout_losses(None, None, i64(1), 9, None, None, None, None, None) == -1
    buf = (ctypes.c_float * 64)()
    b8 = (ctypes.c_uint8 * 64)()
    assert lib.ld_layout_losses(buf, b8, i64(1), 65, buf, buf, None, None, None) == -1 and b"slots" in lib.ld_last_error()
    assert lib.ld_layout_losses(buf, b8, i64(0), 9, buf, buf, None, None, None) == 0            # empty batch: nothing launched
    assert lib.ld_giou_loss(buf, buf, i64(0), buf, None, None) == -1
    assert lib.ld_rows_scale(buf, buf, buf, i64(4), i64(0), 0, None) == -1
    assert lib.ld_rows_scale(buf, buf, buf, i64(0), i64(1), 0, None) == 0
    assert lib.ld_layout_pair_metrics(buf, None, b8, i64(1), 9, buf, buf, None) == -1
    assert lib.ld_normalize_u8_image(b8, buf, i64(1), i64(3), i64(3), buf, buf, None) == -1 and b"multiple of 4" in lib.ld_last_error()
    assert lib.ld_cross_entropy(None, 1, i64(8), None, None, None, 1, i64(8), i64(1), 8, f32(0.0), i64(-100), f32(1.0), None, None) == -1


def test_product_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from layoutdetr_b200.torch_utils.ops import bias_act, upfirdn2d
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bias_act.bias_act(torch.zeros(2, 3), torch.zeros(3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        upfirdn2d.upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(2, 2))
