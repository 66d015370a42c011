"""Overlay: `training.detr_transformer` -> layoutdetr_b200.training.detr_transformer (sm_100a implementation, same public names)."""
from layoutdetr_b200.training.detr_transformer import *  # noqa: F401,F403
from layoutdetr_b200.training import detr_transformer as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
