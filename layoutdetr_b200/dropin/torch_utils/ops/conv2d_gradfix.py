"""Overlay: `torch_utils.ops.conv2d_gradfix` -> layoutdetr_b200.torch_utils.ops.conv2d_gradfix (ld_* sm_100a kernels behind the same API)."""
from layoutdetr_b200.torch_utils.ops import conv2d_gradfix as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
