"""Overlay: `training.networks_detr` -> layoutdetr_b200.training.networks_detr (sm_100a implementation, same public names)."""
from layoutdetr_b200.training.networks_detr import *  # noqa: F401,F403
from layoutdetr_b200.training import networks_detr as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
