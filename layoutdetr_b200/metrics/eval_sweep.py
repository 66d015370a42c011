"""Evaluation sweep of the layout metrics on the GPU (BASELINE configs[4]):
`layout_fid50k_val` + `overlap50k_alignment50k_layoutwise_iou50k_layoutwise_docsim50k`.

Reference flow (metrics/metric_utils_layout.py:255-340, metrics/layout_frechet_inception_distance.py:22-41,
metrics/overlap50k_alignment50k_layoutwise_iou50k_layoutwise_docsim50k.py:17-47): for every batch of 8, run G_ema,
broadcast each rank's tensors `num_gpus` times, copy them to the host, append to NumPy lists; at the end loop over the
layouts in Python for IoU / DocSim and build the FID moments on the host.

Here: batches of 64 per GPU; G_ema forward, LayoutNet features of the real and generated layouts, overlap / alignment
and layout-wise IoU / DocSim are kernel launches on device tensors; per-rank results are ACCUMULATED on the device
(sums and fp64 raw moments — the same statistics FeatureStats keeps) and exchanged ONCE at the end with a single
all-reduce.  Only the 256 x 256 matrix square root of the FID stays on the host (scipy, fp64), as in the reference.
"""
import numpy as np
import torch

from .. import kernels as K
from .. import box_ops


class LayoutEvalAccumulator:
    """Device-side running statistics of one evaluation sweep."""

    def __init__(self, feature_dim=256, device="cuda"):
        dev = torch.device(device)
        self.n = torch.zeros((), dtype=torch.float64, device=dev)
        self.sums = torch.zeros(4, dtype=torch.float64, device=dev)             # overlap, alignment, IoU, DocSim
        self.raw_mean = torch.zeros((2, feature_dim), dtype=torch.float64, device=dev)      # [real | generated]
        self.raw_cov = torch.zeros((2, feature_dim, feature_dim), dtype=torch.float64, device=dev)

    @torch.no_grad()
    def update(self, bbox_real, bbox_fake, mask, feat_real=None, feat_fake=None):
        """bbox_* [B, N, 4] fp32, mask [B, N] bool (True = real element), feat_* fp32 [B, F] (LayoutNet.extract_features)."""
        overlap, alignment = box_ops.layout_losses(bbox_fake, mask)
        iou, docsim = K.layout_pair_metrics(bbox_real, bbox_fake, mask)
        self.n += bbox_real.shape[0]
        self.sums += torch.stack([overlap, alignment, iou, docsim]).double().sum(dim=1)
        for i, f in enumerate((feat_real, feat_fake)):
            if f is not None:
                x64 = f.double()                                             # FeatureStats.append: fp64 raw moments (:108-111)
                self.raw_mean[i] += x64.sum(dim=0)
                self.raw_cov[i] += x64.t() @ x64
        return overlap, alignment, iou, docsim

    def all_reduce(self):
        """The sweep's single exchange step: SUM over ranks of the running statistics (no-op without a process group)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            flat = torch.cat([self.n.reshape(1), self.sums, self.raw_mean.reshape(-1), self.raw_cov.reshape(-1)])
            dist.all_reduce(flat)
            F = self.raw_mean.shape[1]
            self.n = flat[0]
            self.sums = flat[1:5].clone()
            self.raw_mean = flat[5:5 + 2 * F].reshape(2, F).clone()
            self.raw_cov = flat[5 + 2 * F:].reshape(2, F, F).clone()
        return self

    def result(self, with_fid=True):
        n = float(self.n.item())
        s = (self.sums / self.n).cpu().numpy()
        out = dict(num_items=int(n), overlap=float(s[0]), alignment=float(s[1]), layoutwise_iou=float(s[2]), layoutwise_docsim=float(s[3]))
        if with_fid:
            mean = (self.raw_mean / self.n).cpu().numpy()
            cov = (self.raw_cov / self.n).cpu().numpy()
            cov = cov - np.einsum("ki,kj->kij", mean, mean)                     # get_mean_cov (:133-138)
            out["layout_fid"] = frechet_distance(mean[1], cov[1], mean[0], cov[0])
        return out


def frechet_distance(mu_gen, sigma_gen, mu_real, sigma_real):
    """metrics/layout_frechet_inception_distance.py:36-39 (host, fp64)."""
    import scipy.linalg
    m = np.square(mu_gen - mu_real).sum()
    s = scipy.linalg.sqrtm(np.dot(sigma_gen, sigma_real))
    if isinstance(s, tuple):                                                    # scipy < 1.16 with disp=False semantics
        s = s[0]
    return float(np.real(m + np.trace(sigma_gen + sigma_real - s * 2)))


@torch.no_grad()
def run_sweep(G, layoutnet, batches, z_seed=0, label_idx_replace=False, label_idx_replace_2=False, G_kwargs=None):
    """`batches`: iterable of dicts with the reference loader's keys on the device (bboxes/bbox_real, labels/bbox_class,
    texts/bbox_text, patches/bbox_patch, mask or padding_mask, background).  Returns the metric dict of this rank's
    share after the final all-reduce (identical on every rank)."""
    acc = None
    gen = None
    for bt in batches:
        bbox_real = bt.get("bbox_real", bt.get("bboxes")).float()
        bbox_class = bt.get("bbox_class", bt.get("labels")).long()
        padding_mask = bt["padding_mask"] if "padding_mask" in bt else ~bt["mask"].bool()
        mask = ~padding_mask
        dev = bbox_real.device
        if acc is None:
            acc = LayoutEvalAccumulator(device=dev)
            gen = torch.Generator(device=dev).manual_seed(z_seed)
        z = torch.randn((bbox_class.shape[0], bbox_class.shape[1], G.z_dim), device=dev, generator=gen)
        bbox_fake = G(z=z, bbox_class=bbox_class, bbox_real=bbox_real, bbox_text=bt.get("bbox_text", bt.get("texts")),
                      bbox_patch=bt.get("bbox_patch", bt.get("patches")), padding_mask=padding_mask, background=bt["background"],
                      c=bt.get("c"), **(G_kwargs or {}))
        f_real = f_fake = None
        if layoutnet is not None:
            f_real = layoutnet.extract_features(bbox_real, bbox_class, padding_mask, label_idx_replace, label_idx_replace_2)
            f_fake = layoutnet.extract_features(bbox_fake, bbox_class, padding_mask, label_idx_replace, label_idx_replace_2)
        acc.update(bbox_real, bbox_fake, mask, f_real, f_fake)
    return acc.all_reduce().result(with_fid=layoutnet is not None)
