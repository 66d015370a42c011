"""ctypes loader for liblayoutdetr_sm100.so.

The product path never falls back to PyTorch/CPU math: if the library is missing the import of any
compute module raises (build it with `python -m layoutdetr_b200.build`).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LD_LIB_PATH_OVERRIDE") or os.path.join(_HERE, "liblayoutdetr_sm100.so")   # override: A/B of two builds
_lib = None


class LayoutDetrKernelError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LayoutDetrKernelError(
                "liblayoutdetr_sm100.so not found at %s — build it with `python -m layoutdetr_b200.build` "
                "(there is no CPU / PyTorch fallback for the hot path)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.ld_last_error.restype = ctypes.c_char_p
        _lib.ld_launch_count.restype = ctypes.c_int64
    return _lib


def check(code, what=""):
    if code != 0:
        msg = lib().ld_last_error().decode("utf-8", "replace")
        raise LayoutDetrKernelError("%s failed (code %d): %s" % (what or "kernel call", code, msg))


def launch_count():
    return int(lib().ld_launch_count())


def launch_count_reset():
    lib().ld_launch_count_reset()
