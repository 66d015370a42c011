"""Overlay: `training.loss.StyleGAN2Loss` (the class name train.py passes to construct_class_by_name) ->
layoutdetr_b200.training.loss: same constructor arguments, fused box-loss kernels, lane-aware backward."""
from layoutdetr_b200.training.loss import *  # noqa: F401,F403
from layoutdetr_b200.training import loss as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
