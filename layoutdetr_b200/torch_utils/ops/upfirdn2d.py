"""Drop-in for the reference's torch_utils/ops/upfirdn2d.py (`setup_filter`, `upfirdn2d`, `upsample2d`, `downsample2d`,
`filter2d`; reference :119-384) on the ld_upfirdn2d sm_100a kernel.  CUDA tensors only.

The op is  y = decimate_down( FIR_f( pad( zero_insert_up(x) ) ) ) * gain  per axis.  Everything here is expressed through
one per-axis record `_Axis(up, down, lo, hi)`:
  * output length        (n * up + lo + hi - taps) // down + 1
  * its adjoint          swap up / down, flip the filter, lo' = taps - lo - 1, hi' = n*up - m*down + lo - up + 1
                         (so the backward pass — of any order — is the same kernel with another `_Axis`; reference :248-270)
  * "centred" padding    lo = (taps + up - down) // 2, hi = (taps - up - down + 1) // 2, which keeps an up-sampled /
                         filtered / down-sampled image aligned with its input; it is the one formula behind the three
                         convenience wrappers (reference :274-384 states it three times, once per wrapper).
"""
import collections

import numpy as np
import torch

from ... import kernels as K

_Axis = collections.namedtuple("_Axis", "up down lo hi")


def _pair(v, what):
    """int | [x, y] -> (x, y) of python ints."""
    if isinstance(v, (int, np.integer)):
        return int(v), int(v)
    if not (isinstance(v, (list, tuple)) and len(v) == 2 and all(isinstance(t, (int, np.integer)) for t in v)):
        raise AssertionError("%s must be an int or a pair of ints, got %r" % (what, v))
    return int(v[0]), int(v[1])


def _parse_scaling(scaling):
    sx, sy = _pair(scaling, "scaling")
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    """int | [x, y] | [x0, x1, y0, y1] -> (x0, x1, y0, y1)."""
    if isinstance(padding, (list, tuple)) and len(padding) == 4:
        assert all(isinstance(t, (int, np.integer)) for t in padding)
        return tuple(int(t) for t in padding)
    px, py = _pair(padding, "padding")
    return px, px, py, py


def _get_filter_size(f):
    """(taps along x, taps along y) of a separable (1-D) or full (2-D) filter; None is the identity."""
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in (1, 2)
    return int(f.shape[-1]), int(f.shape[0])


def _centred(taps, up=1, down=1):
    return (taps + up - down) // 2, (taps - up - down + 1) // 2


def _axes(up, down, padding):
    (ux, uy), (dx, dy) = _parse_scaling(up), _parse_scaling(down)
    x0, x1, y0, y1 = _parse_padding(padding)
    return _Axis(ux, dx, x0, x1), _Axis(uy, dy, y0, y1)


def _adjoint(ax, n_in, n_out, taps):
    return _Axis(ax.down, ax.up, taps - ax.lo - 1, n_in * ax.up - n_out * ax.down + ax.lo - ax.up + 1)


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """FIR taps as the fp32 tensor `upfirdn2d` expects (reference :59-107): short 1-D tap lists become their outer product,
    long ones (>= 8 taps) stay separable; unit DC gain unless normalize=False; `gain` is spread evenly over the axes."""
    taps = torch.as_tensor(1 if f is None else f, dtype=torch.float32)
    if taps.ndim == 0:
        taps = taps.reshape(1)
    assert taps.ndim in (1, 2) and taps.numel() > 0
    keep_1d = (taps.ndim == 1 and taps.numel() >= 8) if separable is None else bool(separable)
    if taps.ndim == 1 and not keep_1d:
        taps = torch.outer(taps, taps)
    assert taps.ndim == (1 if keep_1d else 2)
    if normalize:
        taps = taps / taps.sum()
    if flip_filter:
        taps = taps.flip(tuple(range(taps.ndim)))
    return (taps * gain ** (taps.ndim / 2)).to(device=device)


def _run(x, f, ax, ay, flip, gain):
    """One (2-D filter) or two (separable filter) launches of ld_upfirdn2d."""
    if f.ndim == 2:
        return K.upfirdn2d_raw(x, f, ax.up, ay.up, ax.down, ay.down, ax.lo, ax.hi, ay.lo, ay.hi, flip, gain)
    y = K.upfirdn2d_raw(x, f.unsqueeze(0), ax.up, 1, ax.down, 1, ax.lo, ax.hi, 0, 0, flip, 1.0)
    return K.upfirdn2d_raw(y, f.unsqueeze(1), 1, ay.up, 1, ay.down, 0, 0, ay.lo, ay.hi, flip, gain)


class _UpfirdnFn(torch.autograd.Function):
    """The geometry travels in `ctx`; backward applies the same Function with the adjoint geometry, so higher-order
    gradients (R1 / path-length regularisation) need nothing extra."""

    @staticmethod
    def forward(ctx, x, f, ax, ay, flip, gain):
        assert isinstance(x, torch.Tensor) and x.ndim == 4
        if f is None:
            f = torch.ones((1, 1), dtype=torch.float32, device=x.device)
        elif f.ndim == 1 and f.shape[0] == 1:            # a single separable tap acts on both axes
            f = f.square().unsqueeze(0)
        assert f.ndim in (1, 2) and f.dtype == torch.float32
        f = f.to(x.device)
        ctx.geom = (ax, ay, flip, gain, x.shape[2], x.shape[3])
        ctx.save_for_backward(f)
        return _run(x, f, ax, ay, flip, gain)

    @staticmethod
    def backward(ctx, dy):
        (f,) = ctx.saved_tensors
        ax, ay, flip, gain, ih, iw = ctx.geom
        if not ctx.needs_input_grad[0]:
            return (None,) * 6
        fw, fh = _get_filter_size(f)
        bx = _adjoint(ax, iw, dy.shape[3], fw)
        by = _adjoint(ay, ih, dy.shape[2], fh)
        return (_UpfirdnFn.apply(dy, f, bx, by, not flip, gain),) + (None,) * 5


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    assert isinstance(x, torch.Tensor)
    assert impl in ('ref', 'cuda')
    if not x.is_cuda:
        raise RuntimeError("layoutdetr_b200 upfirdn2d: CUDA tensors only (no CPU fallback)")
    ax, ay = _axes(up, down, padding)
    return _UpfirdnFn.apply(x, f, ax, ay, bool(flip_filter), float(gain))


def _resample(x, f, up, down, padding, flip_filter, gain, impl):
    """Shared body of filter2d / upsample2d / downsample2d: user padding on top of the centred padding of the filter."""
    (ux, uy), (dx, dy) = _parse_scaling(up), _parse_scaling(down)
    x0, x1, y0, y1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    cx, cy = _centred(fw, ux, dx), _centred(fh, uy, dy)
    pads = [x0 + cx[0], x1 + cx[1], y0 + cy[0], y1 + cy[1]]
    return upfirdn2d(x, f, up=[ux, uy], down=[dx, dy], padding=pads, flip_filter=flip_filter, gain=gain * ux * uy, impl=impl)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    return _resample(x, f, 1, 1, padding, flip_filter, gain, impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    return _resample(x, f, up, 1, padding, flip_filter, gain, impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    return _resample(x, f, 1, down, padding, flip_filter, gain, impl)
