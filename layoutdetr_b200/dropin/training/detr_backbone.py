"""Overlay: `training.detr_backbone` -> layoutdetr_b200.training.detr_backbone (sm_100a implementation, same public names)."""
from layoutdetr_b200.training.detr_backbone import *  # noqa: F401,F403
from layoutdetr_b200.training import detr_backbone as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
