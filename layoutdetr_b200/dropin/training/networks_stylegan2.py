"""Overlay: `training.networks_stylegan2` -> layoutdetr_b200.training.networks_stylegan2 (sm_100a implementation, same public names)."""
from layoutdetr_b200.training.networks_stylegan2 import *  # noqa: F401,F403
from layoutdetr_b200.training import networks_stylegan2 as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
