"""Overlay: `training.training_loop.training_loop` -> layoutdetr_b200.training.training_loop (same keyword arguments and run-dir
artefacts; the iteration is this package's Trainer / GraphedStep)."""
from layoutdetr_b200.training.training_loop import *  # noqa: F401,F403
from layoutdetr_b200.training import training_loop as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
