"""Overlay: `training.dataset_layoutganpp` -> layoutdetr_b200.training.dataset_layoutganpp (same zip format, class and sample
dict; `lean=True` in `training_set_kwargs` skips the PNGs the training path never reads)."""
from layoutdetr_b200.training.dataset_layoutganpp import *  # noqa: F401,F403
from layoutdetr_b200.training import dataset_layoutganpp as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
