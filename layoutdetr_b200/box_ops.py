"""Layout box losses on `[B, N, 4]` (xc, yc, w, h) boxes: generalized-IoU, pairwise overlap, alignment.

Semantics: reference metrics/metric_layoutnet.py:153-201, 245-275 (see the oracle restatement for the
line-by-line citation).  These are O(B * 81) scalar operations on the device — glue next to the 50 TFLOP step.
"""
import torch


def _ltrb(b):
    xc, yc, w, h = b.unbind(-1)
    return xc - w / 2, yc - h / 2, xc + w / 2, yc + h / 2


def giou_loss(fake, real):
    """mean over index-paired rows of 1 - GIoU; fake/real: [M, 4]."""
    l1, t1, r1, b1 = _ltrb(fake)
    l2, t2, r2, b2 = _ltrb(real)
    a1 = (r1 - l1) * (b1 - t1)
    a2 = (r2 - l2) * (b2 - t2)
    iw = torch.minimum(r1, r2) - torch.maximum(l1, l2)
    ih = torch.minimum(b1, b2) - torch.maximum(t1, t2)
    inter = torch.where((iw > 0) & (ih > 0), iw * ih, torch.zeros_like(iw))
    union = a1 + a2 - inter
    hull = (torch.maximum(r1, r2) - torch.minimum(l1, l2)) * (torch.maximum(b1, b2) - torch.minimum(t1, t2))
    giou = inter / union - (hull - union) / hull
    return (1 - giou).mean()


def overlap(bbox, mask):
    """[B]: sum_{i != j} area(i ∩ j) / area(i) over valid slots, divided by the number of valid slots."""
    bbox = bbox.masked_fill(~mask.unsqueeze(-1), 0)          # masked_fill (not *0): its backward stops the 0/0 NaNs of padded slots
    l, t, r, b = _ltrb(bbox)
    area = (r - l) * (b - t)
    iw = torch.minimum(r[:, :, None], r[:, None, :]) - torch.maximum(l[:, :, None], l[:, None, :])
    ih = torch.minimum(b[:, :, None], b[:, None, :]) - torch.maximum(t[:, :, None], t[:, None, :])
    inter = torch.where((iw > 0) & (ih > 0), iw * ih, torch.zeros_like(iw))
    n = bbox.shape[1]
    inter = inter.masked_fill(torch.eye(n, dtype=torch.bool, device=bbox.device), 0)
    ratio = torch.nan_to_num(inter / area[:, :, None])
    return ratio.sum(dim=(1, 2)) / mask.float().sum(-1)


def alignment(bbox, mask):
    """[B]: -log(1 - min over other slots and the 6 edge/centre coordinates of |delta|), summed over valid slots."""
    l, t, r, b = _ltrb(bbox)
    xc, yc = bbox[..., 0], bbox[..., 1]
    X = torch.stack([l, xc, r, t, yc, b], dim=1)                       # [B, 6, N]
    D = (X.unsqueeze(-1) - X.unsqueeze(-2)).abs()                      # [B, 6, N, N]
    n = bbox.shape[1]
    D = D.masked_fill(torch.eye(n, dtype=torch.bool, device=bbox.device), 1.0)
    D = D.permute(0, 2, 1, 3)                                          # [B, N, 6, N]
    D = torch.where(mask[:, :, None, None], D, torch.ones_like(D))
    m = D.amin(dim=(-1, -2))
    m = torch.where(m == 1.0, torch.zeros_like(m), m)
    return (-torch.log(1 - m)).sum(-1) / mask.float().sum(-1)
