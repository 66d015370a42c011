"""One data-parallel training iteration of LayoutDETR (the body of the reference's hot loop,
training/training_loop.py:274-328): phases Gmain / Greg / Dmain / Dreg, gradient all-reduce, Adam, G_ema.

Differences from the reference are mechanical, not numerical:
  * parameters, gradients and Adam state live in flat buffers (flat.FlatParams): the all-reduce runs in place
    on the gradient buffer (no torch.cat / split) and nan_to_num + Adam + bf16 refresh is one kernel;
  * the EMA skips parameters that cannot change (the frozen text encoder) — lerp(p, p, beta) == p;
  * Greg / Dreg are no-ops at the reference defaults (pl_weight = r1_gamma = 0: zero_grad + an optimizer step
    over parameters without gradients), so they launch nothing.
"""
import copy

import torch

from ..flat import FlatParams
from .loss import StyleGAN2Loss


class Trainer:
    def __init__(self, G, D, device, batch_size, num_gpus=1, lr=1e-5, betas=(0.0, 0.99), eps=1e-8,
                 G_reg_interval=4, D_reg_interval=16, ema_kimg=10.0, ema_rampup=0.05, loss_kwargs=None, process_group=None):
        self.G, self.D = G, D
        self.G_ema = copy.deepcopy(G).eval()
        self.device = device
        self.batch_size = batch_size              # total batch over all ranks
        self.num_gpus = num_gpus
        self.pg = process_group
        self.loss = StyleGAN2Loss(device=device, G=G, D=D, **(loss_kwargs or {}))
        self.flat = {"G": FlatParams(G, beta1=betas[0]), "D": FlatParams(D, beta1=betas[0])}
        self.flat_ema = FlatParams(self.G_ema, beta1=0.0, with_grad=False)
        self.opt = {}
        for name, interval in (("G", G_reg_interval), ("D", D_reg_interval)):
            r = interval / (interval + 1) if interval is not None else 1.0        # lazy-regularisation rescale, :191-195
            self.opt[name] = dict(lr=lr * r, beta1=betas[0] ** r, beta2=betas[1] ** r, eps=eps)
        self.ema_kimg, self.ema_rampup = ema_kimg, ema_rampup
        self.cur_nimg = 0
        for m in (G, D, self.G_ema):
            m.requires_grad_(False)

    def _phase(self, name, batch, gen_z):
        mod = self.G if name == "G" else self.D
        flat = self.flat[name]
        flat.zero_grad()
        mod.requires_grad_(True)
        mod.text_encoder.requires_grad_(False)
        self.loss.accumulate_gradients(phase=name + "main", bbox_real=batch["bbox_real"], bbox_class=batch["bbox_class"],
                                       bbox_text=batch["bbox_text"], bbox_patch=batch["bbox_patch"],
                                       padding_mask=batch["padding_mask"], background=batch["background"], real_c=batch["c"],
                                       gen_z=gen_z, gen_c=batch["c"], gain=1, cur_nimg=self.cur_nimg)
        mod.requires_grad_(False)
        scale = 1.0
        if self.num_gpus > 1:
            torch.distributed.all_reduce(flat.g, group=self.pg)
            scale = 1.0 / self.num_gpus
        o = self.opt[name]
        flat.adam_step(o["lr"], o["beta1"], o["beta2"], o["eps"], grad_scale=scale)

    def iteration(self, batch, z_g, z_d):
        self._phase("G", batch, z_g)
        self._phase("D", batch, z_d)
        ema_nimg = self.ema_kimg * 1000
        if self.ema_rampup is not None:
            ema_nimg = min(ema_nimg, self.cur_nimg * self.ema_rampup)
        ema_beta = 0.5 ** (self.batch_size / max(ema_nimg, 1e-8))
        self.flat_ema.ema_from(self.flat["G"], ema_beta)
        self.cur_nimg += self.batch_size
        return self.loss.last
