"""Layout box losses on `[B, N, 4]` (xc, yc, w, h) boxes: generalized-IoU, pairwise overlap, alignment.

Semantics: reference metrics/metric_layoutnet.py:153-201, 245-275 (line-by-line restatement: oracle/layoutdetr_oracle.py).
Each loss is ONE launch of csrc/box_loss.cu, which also writes the analytic Jacobian rows; backward is one
`ld_rows_scale` launch (upstream gradient x stored Jacobian).  CUDA tensors only — there is no CPU / eager path.
"""
import torch

from . import kernels as K


class _LayoutLossesFn(torch.autograd.Function):
    """(overlap [B], alignment [B]) of a batch of layouts; `valid` = ~padding_mask."""

    @staticmethod
    def forward(ctx, bbox, valid):
        want = ctx.needs_input_grad[0]
        ov, al, j_ov, j_al = K.layout_losses(bbox.detach().float(), valid, want)
        if want:
            ctx.save_for_backward(j_ov, j_al)
        return ov, al

    @staticmethod
    def backward(ctx, g_ov, g_al):
        j_ov, j_al = ctx.saved_tensors
        d = K.rows_scale(j_ov, g_ov.float())
        K.rows_scale(j_al, g_al.float(), out=d, accumulate=True)
        return d, None


class _GiouLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fake, real):
        want = ctx.needs_input_grad[0]
        loss, jac = K.giou_loss(fake.detach().float(), real.detach().float(), want)
        if want:
            ctx.save_for_backward(jac)
        return loss

    @staticmethod
    def backward(ctx, g):
        (jac,) = ctx.saved_tensors
        return K.rows_scale(jac, g.float()), None


def giou_loss(fake, real):
    """mean over index-paired rows of 1 - GIoU; fake/real: [M, 4]."""
    return _GiouLossFn.apply(fake, real)


def layout_losses(bbox, mask):
    """(overlap [B], alignment [B]) in one launch; mask: [B, N] bool, True = real element."""
    return _LayoutLossesFn.apply(bbox, mask)


def overlap(bbox, mask):
    """[B]: sum_{i != j} area(i ∩ j) / area(i) over valid slots, divided by the number of valid slots."""
    return _LayoutLossesFn.apply(bbox, mask)[0]


def alignment(bbox, mask):
    """[B]: -log(1 - min over other slots and the 6 edge/centre coordinates of |delta|), summed over valid slots."""
    return _LayoutLossesFn.apply(bbox, mask)[1]
