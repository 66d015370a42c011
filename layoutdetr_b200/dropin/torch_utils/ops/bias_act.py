"""Overlay: `torch_utils.ops.bias_act` -> layoutdetr_b200.torch_utils.ops.bias_act (ld_* sm_100a kernels behind the same API)."""
from layoutdetr_b200.torch_utils.ops import bias_act as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
