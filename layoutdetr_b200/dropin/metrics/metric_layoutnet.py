"""Overlay: `metrics.metric_layoutnet` -> the reference's module (LayoutFID, compute_iou*, compute_docsim*, ... stay as they
are) with the box losses of the training objective and the Hungarian max-IoU metric replaced by the sm_100a kernels."""
import importlib.util
import os

_ref = os.path.join(os.environ.get("LAYOUTDETR_REFERENCE", "/root/reference"), "metrics", "metric_layoutnet.py")
if os.path.exists(_ref):
    try:
        _spec = importlib.util.spec_from_file_location("_layoutdetr_reference_metric_layoutnet", _ref)
        _mod = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_mod)
        globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
    except ImportError:            # optional host-side dependencies of the reference file (pytorch_fid, ...) are missing
        pass
from layoutdetr_b200.metrics.metric_layoutnet import (compute_alignment, compute_maximum_iou, compute_overlap,  # noqa: E402,F401
                                                      generalized_iou_loss, layout_overlap_alignment)
