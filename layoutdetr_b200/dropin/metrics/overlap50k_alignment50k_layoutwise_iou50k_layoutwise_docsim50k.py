"""Overlay: `metrics.overlap50k_..._docsim50k.compute_overlap_alignment_laywise_IoU_layerwise_DocSim` -> layoutdetr_b200.metrics.sweep_entry."""
from layoutdetr_b200.metrics.sweep_entry import compute_overlap_alignment_laywise_IoU_layerwise_DocSim  # noqa: F401
