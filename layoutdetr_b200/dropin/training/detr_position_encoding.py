"""Overlay: `training.detr_position_encoding` -> layoutdetr_b200.training.detr_position_encoding (sm_100a implementation, same public names)."""
from layoutdetr_b200.training.detr_position_encoding import *  # noqa: F401,F403
from layoutdetr_b200.training import detr_position_encoding as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
