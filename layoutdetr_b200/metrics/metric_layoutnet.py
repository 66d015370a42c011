"""Box losses of the LayoutDETR training objective: mirror of the functions training/loss.py imports from the
reference's metrics/metric_layoutnet.py (generalized_iou_loss :245-275, compute_overlap :153-179,
compute_alignment :182-201) plus `compute_maximum_iou` (:100-150, the Hungarian max-IoU metric — dead code in the
reference, served here by the batched ld_lsap kernel).

The three losses act on `[B, 9, 4]` boxes: negligible FLOPs, so they are evaluated by fused CUDA kernels
(csrc/box_loss.cu: value + analytic Jacobian in one launch, backward = one scaling launch) instead of ~60 eager launches.
"""
import torch

from .. import box_ops


def convert_xywh_to_ltrb(bbox):
    xc, yc, w, h = bbox
    return [xc - w / 2, yc - h / 2, xc + w / 2, yc + h / 2]


def generalized_iou_loss(bbox_fake, bbox_real):
    return box_ops.giou_loss(bbox_fake, bbox_real)


def compute_overlap(bbox, mask):
    return box_ops.overlap(bbox, mask)


def compute_alignment(bbox, mask):
    return box_ops.alignment(bbox, mask)


def layout_overlap_alignment(bbox, mask):
    """(compute_overlap(bbox, mask), compute_alignment(bbox, mask)) from a single launch."""
    return box_ops.layout_losses(bbox, mask)
