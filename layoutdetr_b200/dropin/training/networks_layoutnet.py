"""Overlay: `training.networks_layoutnet` -> layoutdetr_b200.training.networks_layoutnet (sm_100a implementation, same public names)."""
from layoutdetr_b200.training.networks_layoutnet import *  # noqa: F401,F403
from layoutdetr_b200.training import networks_layoutnet as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
