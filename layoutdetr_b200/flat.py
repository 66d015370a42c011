"""Flat parameter storage for the data-parallel optimizer step.

The reference concatenates all gradients into a temporary, all-reduces it, splits it back and then runs a
per-tensor Adam (training/training_loop.py:303-313): ~16 B/param of copies before the optimizer even starts.
Here the trainable parameters of a module live in ONE fp32 buffer; `.grad` of every parameter is a view of ONE
gradient buffer (the backward kernels accumulate straight into it), so the all-reduce needs no gather/scatter
and nan_to_num + Adam + the bf16 tensor-core shadow refresh are a single kernel over the flat range.
"""
import torch

from . import engine as E
from . import kernels as K


class FlatParams:
    def __init__(self, module, exclude_prefixes=("text_encoder.",), beta1=0.0, with_grad=True):
        named = [(n, p) for n, p in module.named_parameters() if not n.startswith(tuple(exclude_prefixes))]
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 7) // 8 * 8            # 32-byte aligned fp32 / 16-byte aligned bf16 views
        self.offsets, self.numel = offs, total
        self.p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.g = torch.zeros(total, dtype=torch.float32, device=dev) if with_grad else None
        self.v = torch.zeros(total, dtype=torch.float32, device=dev) if with_grad else None
        self.m = torch.zeros(total, dtype=torch.float32, device=dev) if (beta1 != 0.0 and with_grad) else None
        self.p16 = torch.empty(total, dtype=torch.bfloat16, device=dev)
        self.step = 0
        self.hyper = torch.zeros(3, dtype=torch.float32, device=dev)     # {lr, 1 - beta1^t, 1 - beta2^t} read by the kernel
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                n = p.numel()
                self.p[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.p[o:o + n].view(p.shape)
                p.grad = self.g[o:o + n].view(p.shape) if with_grad else None
            self.p16.copy_(K.to_bf16(self.p))
        for p, o in zip(self.params, offs):
            if p.ndim == 2 and p.shape[1] % 8 == 0:
                E.register_managed(p, self.p16[o:o + p.numel()].view(p.shape))
        E.bump_generation(self.params)

    def zero_grad(self):
        self.g.zero_()
        for p, o in zip(self.params, self.offsets):       # the loop may have dropped the views (set_to_none)
            if p.grad is None or p.grad.data_ptr() != self.g.data_ptr() + 4 * o:
                p.grad = self.g[o:o + p.numel()].view(p.shape)

    def set_hyper(self, lr, beta1, beta2):
        """Advance the step counter and publish the step-dependent scalars on the device (outside any graph capture)."""
        self.step += 1
        h = torch.tensor([lr, 1.0 - beta1 ** self.step, 1.0 - beta2 ** self.step], dtype=torch.float32)
        if self.hyper.is_cuda:
            # pinned source + asynchronous copy: the host must not wait for the previous replay here (a fresh pinned tensor per
            # call — the caching host allocator keeps it alive until the copy has run)
            self.hyper.copy_(h.pin_memory(), non_blocking=True)
        else:
            self.hyper.copy_(h)

    def adam_step(self, lr, beta1, beta2, eps, grad_scale=1.0):
        if not torch.cuda.is_current_stream_capturing():
            self.set_hyper(lr, beta1, beta2)          # a captured graph re-reads `hyper`, refreshed by its owner before replay
        K.adam_flat(self.p, self.g, self.m, self.v, self.p16, lr, beta1, beta2, eps, max(1, self.step), grad_scale,
                    hyper_dev=self.hyper)
        E.bump_generation(self.params)

    def ema_from(self, src, beta):
        """self = lerp(src, self, beta) over the flat range (same parameter order required)."""
        assert self.numel == src.numel
        K.ema_flat(self.p, src.p, self.p16, beta)
        E.bump_generation(self.params)
