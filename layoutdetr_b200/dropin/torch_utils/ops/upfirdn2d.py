"""Overlay: `torch_utils.ops.upfirdn2d` -> layoutdetr_b200.torch_utils.ops.upfirdn2d (ld_* sm_100a kernels behind the same API)."""
from layoutdetr_b200.torch_utils.ops import upfirdn2d as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
