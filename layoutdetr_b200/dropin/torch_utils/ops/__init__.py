"""Namespace overlay for `torch_utils.ops`: bias_act / upfirdn2d / conv2d_resample / conv2d_gradfix / fma are the
sm_100a ops; grid_sample_gradfix and filtered_lrelu (unused on the LayoutDETR path) still come from the reference."""
import os

_ref = os.environ.get("LAYOUTDETR_REFERENCE", "/root/reference")
_ref_pkg = os.path.join(_ref, "torch_utils", "ops")
if os.path.isdir(_ref_pkg) and _ref_pkg not in __path__:
    __path__.append(_ref_pkg)
