"""Drop-in for the reference's torch_utils/ops/fma.py: fma(a, b, c) = a * b + c as one kernel with
broadcasting, with the reference's gradient rules (:27-60)."""
import torch

from ... import kernels as K


def fma(a, b, c):
    return _FusedMultiplyAdd.apply(a, b, c)


class _FusedMultiplyAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        if not a.is_cuda:
            raise RuntimeError("layoutdetr_b200 fma: CUDA tensors only (no CPU fallback)")
        shape = torch.broadcast_shapes(a.shape, b.shape, c.shape)
        out = K.fma_f32(a.float().broadcast_to(shape).contiguous(), b.float(), c.float()).to(a.dtype)
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        da = db = dc = None
        if ctx.needs_input_grad[0]:
            da = _unbroadcast(dout * b, a.shape)
        if ctx.needs_input_grad[1]:
            db = _unbroadcast(dout * a, b.shape)
        if ctx.needs_input_grad[2]:
            dc = _unbroadcast(dout, ctx.c_shape)
        return da, db, dc


def _unbroadcast(x, shape):
    extra_dims = x.ndim - len(shape)
    assert extra_dims >= 0
    dim = [i for i in range(x.ndim) if x.shape[i] > 1 and (i < extra_dims or shape[i - extra_dims] == 1)]
    if len(dim):
        x = x.sum(dim=dim, keepdim=True)
    if extra_dims:
        x = x.reshape(-1, *x.shape[extra_dims + 1:])
    assert x.shape == shape
    return x
