"""Drop-in for the reference's torch_utils/ops/bias_act.py: same public API
(`bias_act(x, b, dim, act, alpha, gain, clamp, impl)`, `activation_funcs[act].def_gain`, ...), backed by
the ld_bias_act sm_100a kernel instead of the JIT-compiled plugin (torch_utils/custom_ops.py:62).
First- and second-order gradients are supported where the reference supports them.  CUDA tensors only:
there is no CPU fallback (the reference's `_bias_act_ref` restatement lives in oracle/ as test code).
"""
import numpy as np
import torch

from ... import kernels as K


class EasyDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


activation_funcs = {
    'linear':   EasyDict(def_alpha=0,   def_gain=1,          cuda_idx=1, ref='',  has_2nd_grad=False),
    'relu':     EasyDict(def_alpha=0,   def_gain=np.sqrt(2), cuda_idx=2, ref='y', has_2nd_grad=False),
    'lrelu':    EasyDict(def_alpha=0.2, def_gain=np.sqrt(2), cuda_idx=3, ref='y', has_2nd_grad=False),
    'tanh':     EasyDict(def_alpha=0,   def_gain=1,          cuda_idx=4, ref='y', has_2nd_grad=True),
    'sigmoid':  EasyDict(def_alpha=0,   def_gain=1,          cuda_idx=5, ref='y', has_2nd_grad=True),
    'elu':      EasyDict(def_alpha=0,   def_gain=1,          cuda_idx=6, ref='y', has_2nd_grad=True),
    'selu':     EasyDict(def_alpha=0,   def_gain=1,          cuda_idx=7, ref='y', has_2nd_grad=True),
    'softplus': EasyDict(def_alpha=0,   def_gain=1,          cuda_idx=8, ref='y', has_2nd_grad=True),
    'swish':    EasyDict(def_alpha=0,   def_gain=np.sqrt(2), cuda_idx=9, ref='x', has_2nd_grad=True),
}


def _launch(x, b, xref, yref, dy, grad, dim, spec, alpha, gain, clamp):
    y = torch.empty_like(x)
    sizeB = b.numel() if b is not None else 0
    stepB = x.stride(dim) if b is not None else 1
    K.bias_act_raw(x, b, xref, yref, dy, y, grad, spec.cuda_idx, alpha, gain, clamp, sizeB, stepB)
    return y


def _fmt(t):
    return torch.channels_last if t.ndim == 4 and t.stride(1) == 1 and t.shape[1] > 1 else torch.contiguous_format


_cache = dict()


def _make(dim, act, alpha, gain, clamp):
    key = (dim, act, alpha, gain, clamp)
    if key in _cache:
        return _cache[key]
    spec = activation_funcs[act]

    class BiasActGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y):
            dy = dy.contiguous(memory_format=_fmt(dy))
            dx = _launch(dy, b, x, y, None, 1, dim, spec, alpha, gain, clamp)
            ctx.save_for_backward(dy if spec.has_2nd_grad else None, x, b, y)
            return dx

        @staticmethod
        def backward(ctx, d_dx):
            dy, x, b, y = ctx.saved_tensors
            d_dx = d_dx.contiguous(memory_format=_fmt(d_dx))
            d_dy = d_x = d_b = None
            if ctx.needs_input_grad[0]:
                d_dy = BiasActGrad.apply(d_dx, x, b, y)
            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                d_x = _launch(d_dx, b, x, y, dy, 2, dim, spec, alpha, gain, clamp)
            if spec.has_2nd_grad and ctx.needs_input_grad[2]:
                d_b = d_x.sum([i for i in range(d_x.ndim) if i != dim])
            return d_dy, d_x, d_b, None

    class BiasAct(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            x = x.contiguous(memory_format=_fmt(x))
            b = b.contiguous() if b is not None else None
            y = x
            if act != 'linear' or gain != 1 or clamp >= 0 or b is not None:
                y = _launch(x, b, None, None, None, 0, dim, spec, alpha, gain, clamp)
            keep_x = 'x' in spec.ref or spec.has_2nd_grad
            ctx.save_for_backward(x if keep_x else None, b if keep_x else None, y if ('y' in spec.ref or clamp >= 0) else None)
            return y

        @staticmethod
        def backward(ctx, dy):
            x, b, y = ctx.saved_tensors
            dx = db = None
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                dx = dy
                if act != 'linear' or gain != 1 or clamp >= 0:
                    dx = BiasActGrad.apply(dy, x, b, y)
            if ctx.needs_input_grad[1]:
                db = dx.sum([i for i in range(dx.ndim) if i != dim])
            return dx, db

    _cache[key] = BiasAct
    return BiasAct


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    assert clamp is None or clamp >= 0
    if not x.is_cuda:
        raise RuntimeError("layoutdetr_b200 bias_act: CUDA tensors only (no CPU fallback; the reference restatement "
                           "is test code under oracle/)")
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1
        assert 0 <= dim < x.ndim and b.shape[0] == x.shape[dim]
        b = b.to(x.dtype)
    return _make(dim, act, alpha, gain, clamp).apply(x, b)
