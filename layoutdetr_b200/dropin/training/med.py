"""Overlay: `training.med` -> layoutdetr_b200.training.med (sm_100a implementation, same public names)."""
from layoutdetr_b200.training.med import *  # noqa: F401,F403
from layoutdetr_b200.training import med as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
