"""Overlay: `metrics.layout_frechet_inception_distance.compute_layout_fid` -> layoutdetr_b200.metrics.sweep_entry."""
from layoutdetr_b200.metrics.sweep_entry import compute_layout_fid  # noqa: F401
