"""Drop-in for the reference's torch_utils/ops/fma.py (`fma(a, b, c) = a * b + c`, reference :16-60): one broadcasting kernel
(ld_fma_f32) forward; the gradients are products reduced back to each operand's shape."""
import torch

from ... import kernels as K


def _sum_to_shape(t, shape):
    """Reduce a broadcast result back to the operand shape `shape`: sum over the leading axes the operand does not have and
    over every axis where it has extent 1."""
    shape = tuple(shape)
    lead = t.ndim - len(shape)
    if lead < 0:
        raise ValueError("cannot reduce %s to %s" % (tuple(t.shape), shape))
    if lead:
        t = t.sum(dim=tuple(range(lead)))
    ones = tuple(i for i, (have, want) in enumerate(zip(t.shape, shape)) if want == 1 and have != 1)
    if ones:
        t = t.sum(dim=ones, keepdim=True)
    if tuple(t.shape) != shape:
        raise ValueError("gradient of shape %s does not reduce to %s" % (tuple(t.shape), shape))
    return t


class _Fma(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        if not a.is_cuda:
            raise RuntimeError("layoutdetr_b200 fma: CUDA tensors only (no CPU fallback)")
        full = torch.broadcast_shapes(a.shape, b.shape, c.shape)
        y = K.fma_f32(a.float().broadcast_to(full).contiguous(), b.float(), c.float()).to(a.dtype)
        ctx.save_for_backward(a, b)
        ctx.shapes = (a.shape, b.shape, c.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        a, b = ctx.saved_tensors
        sa, sb, sc = ctx.shapes
        need_a, need_b, need_c = ctx.needs_input_grad
        return (_sum_to_shape(dy * b, sa) if need_a else None,
                _sum_to_shape(dy * a, sb) if need_b else None,
                _sum_to_shape(dy, sc) if need_c else None)


def fma(a, b, c):
    return _Fma.apply(a, b, c)
