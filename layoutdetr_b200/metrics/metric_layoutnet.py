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


# ------------------------------------------------------------------------------------------------
# Hungarian max-IoU metric (reference :100-150; dead code there — no caller — but part of the module's public surface).
def _pairwise_iou(a, b):
    """IoU of every box of a [..., n, 4] with every box of b [..., m, 4] (xc, yc, w, h) -> [..., n, m], NaN -> 0 (reference :66-92)."""
    al, at, ar, ab = a[..., 0] - a[..., 2] / 2, a[..., 1] - a[..., 3] / 2, a[..., 0] + a[..., 2] / 2, a[..., 1] + a[..., 3] / 2
    bl, bt, br, bb = b[..., 0] - b[..., 2] / 2, b[..., 1] - b[..., 3] / 2, b[..., 0] + b[..., 2] / 2, b[..., 1] + b[..., 3] / 2
    lm = torch.maximum(al[..., :, None], bl[..., None, :]); rm = torch.minimum(ar[..., :, None], br[..., None, :])
    tm = torch.maximum(at[..., :, None], bt[..., None, :]); bm = torch.minimum(ab[..., :, None], bb[..., None, :])
    inter = torch.where((lm < rm) & (tm < bm), (rm - lm) * (bm - tm), torch.zeros_like(lm))
    area_a = ((ar - al) * (ab - at))[..., :, None]
    area_b = ((br - bl) * (bb - bt))[..., None, :]
    return torch.nan_to_num(inter / (area_a + area_b - inter))


def compute_maximum_iou(layouts_1, layouts_2, n_jobs=None, device="cuda"):
    """Mean, over the layouts matched between two sets, of the label-aware maximum IoU (reference compute_maximum_iou :140-150).

    layouts_*: lists of (boxes [n, 4] array, labels [n] array).  Layouts are grouped by their sorted label multiset; inside a
    group every layout of set 1 is scored against every layout of set 2 — per label, the optimal one-to-one assignment of the
    boxes carrying that label under IoU (`linear_sum_assignment(maximize=True)`), summed and divided by the element count
    (:100-113) — and the groups' score matrices are assigned once more at layout level (:116-126).  All assignments of a group
    run as batched launches of ld_lsap (bit-identical to scipy, tests/test_lsap.py); the reference walks the pairs one by one in
    a multiprocessing pool.  The reference reshapes the list of pair scores (ordered set-2-major) as [len(set 1), len(set 2)];
    that ordering quirk is reproduced, so results agree for rectangular groups too."""
    import numpy as np
    from .. import kernels as K

    def groups(layouts):
        out = {}
        for b, l in layouts:
            l = np.asarray(l)
            out.setdefault(str(sorted(l.tolist())), []).append((np.asarray(b, dtype=np.float64), l))
        return out

    g1, g2 = groups(layouts_1), groups(layouts_2)
    matched = []
    for key in g1.keys() & g2.keys():
        L1, L2 = g1[key], g2[key]
        N, M = len(L1), len(L2)
        n_el = len(L1[0][1])
        if n_el == 0:
            continue
        pair_scores = torch.zeros((M, N), dtype=torch.float64, device=device)           # [j over set 2, i over set 1]
        for lab in sorted(set(L1[0][1].tolist())):
            A = torch.from_numpy(np.stack([b[l == lab] for b, l in L1])).to(device)     # [N, n, 4]
            Bx = torch.from_numpy(np.stack([b[l == lab] for b, l in L2])).to(device)    # [M, n, 4]
            n = A.shape[1]
            # cost[j, i][r, c] = IoU(box c of layout i (set 1), box r of layout j (set 2)) — the reference's meshgrid orientation
            cost = _pairwise_iou(Bx[:, None], A[None, :]).reshape(M * N, n, n).contiguous()
            rows, cols, status = K.lsap(cost, maximize=True)
            if int(status.abs().sum()) != 0:
                raise RuntimeError("ld_lsap reported an infeasible / non-finite IoU matrix")
            pair_scores += cost[torch.arange(M * N, device=device)[:, None], rows, cols].sum(1).reshape(M, N)
        scores = (pair_scores / n_el).reshape(-1).reshape(N, M)                          # the reference's reshape(N, M) of an M-major list
        r, c, st = K.lsap(scores.contiguous(), maximize=True)
        matched.append(scores[r, c])
    if not matched:
        return float("nan")
    return float(torch.cat(matched).mean())
