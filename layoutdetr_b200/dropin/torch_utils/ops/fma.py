"""Overlay: `torch_utils.ops.fma` -> layoutdetr_b200.torch_utils.ops.fma (ld_* sm_100a kernels behind the same API)."""
from layoutdetr_b200.torch_utils.ops import fma as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
