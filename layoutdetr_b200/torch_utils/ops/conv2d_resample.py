"""Drop-in for the reference's torch_utils/ops/conv2d_resample.py (`conv2d_resample`, reference :47-144): a 2-D convolution
with an optional FIR up- / down-sampling stage, built from this package's conv2d_gradfix (tcgen05 GEMM path) and upfirdn2d
kernels.

Instead of the reference's chain of early returns, the call is first turned into a PLAN — a short list of
("fir", kwargs) / ("conv", kwargs) steps — by `resample_plan`, a pure function of the geometry that the CPU tests check
against the reference's output sizes, and then executed.  User padding is meant w.r.t. the up-sampled image; per axis the FIR
stages add their centred padding (upfirdn2d._centred) on top of it.
"""
import torch

from . import conv2d_gradfix
from . import upfirdn2d
from .upfirdn2d import _centred, _get_filter_size, _parse_padding


def _get_weight_shape(w):
    return [int(sz) for sz in w.shape]


def _conv(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """conv2d_gradfix call; the reference's `flip_weight=True` means plain correlation (what conv2d computes)."""
    kh, kw = int(w.shape[2]), int(w.shape[3])
    if not flip_weight and (kh > 1 or kw > 1):
        w = w.flip([2, 3])
    fn = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
    return fn(x, w, stride=stride, padding=padding, groups=groups)


def resample_plan(kw, kh, fw, fh, up, down, padding):
    """-> [(op, kwargs), ...] with op in {"fir", "pad", "conv"}.

    Per axis (taps t of the FIR, kernel extent k): lo / hi = user padding + centred padding of the up-sampling FIR
    + centred padding of the down-sampling FIR.  Cases:
      1x1 kernel, down only      FIR (decimating) first, then the pointwise conv on the small image
      1x1 kernel, up only        pointwise conv on the small image, then the interpolating FIR
      k x k, down only           FIR at full resolution, strided conv
      k x k, up (+ down)         transposed strided conv (the up-sampling and the conv in one pass), FIR to remove the
                                 imaging, optional decimating FIR
      no resampling              plain conv when the padding is symmetric and non-negative, otherwise an explicit
                                 pad / crop (upfirdn2d with the identity filter) followed by the conv
    """
    x0, x1, y0, y1 = _parse_padding(padding)
    if up > 1:
        cx, cy = _centred(fw, up, 1), _centred(fh, up, 1)
        x0, x1, y0, y1 = x0 + cx[0], x1 + cx[1], y0 + cy[0], y1 + cy[1]
    if down > 1:
        cx, cy = _centred(fw, 1, down), _centred(fh, 1, down)
        x0, x1, y0, y1 = x0 + cx[0], x1 + cx[1], y0 + cy[0], y1 + cy[1]
    pointwise = (kw == 1 and kh == 1)
    pads = [x0, x1, y0, y1]

    if pointwise and down > 1 and up == 1:
        return [("fir", dict(down=down, padding=pads)), ("conv", dict())]
    if pointwise and up > 1 and down == 1:
        return [("conv", dict()), ("fir", dict(up=up, padding=pads, gain=up ** 2))]
    if down > 1 and up == 1:
        return [("fir", dict(padding=pads)), ("conv", dict(stride=down))]
    if up > 1:
        # the transposed conv already grows the image by (k - 1) on the low side and (k - up) on the high side
        x0, x1, y0, y1 = x0 - (kw - 1), x1 - (kw - up), y0 - (kh - 1), y1 - (kh - up)
        tx = max(min(-x0, -x1), 0)           # crop that the transposed conv can do itself (its `padding`)
        ty = max(min(-y0, -y1), 0)
        plan = [("conv", dict(stride=up, padding=[ty, tx], transpose=True)),
                ("fir", dict(padding=[x0 + tx, x1 + tx, y0 + ty, y1 + ty], gain=up ** 2))]
        if down > 1:
            plan.append(("fir", dict(down=down)))
        return plan
    if x0 == x1 and y0 == y1 and x0 >= 0 and y0 >= 0:
        return [("conv", dict(padding=[y0, x0]))]
    return [("pad", dict(padding=pads)), ("conv", dict())]


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in (1, 2) and f.dtype == torch.float32)
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    _, _, kh, kw = _get_weight_shape(w)
    fw, fh = _get_filter_size(f)
    for op, kw_args in resample_plan(kw, kh, fw, fh, up, down, padding):
        kw_args = dict(kw_args)
        if op == "fir":
            x = upfirdn2d.upfirdn2d(x=x, f=f, flip_filter=flip_filter, **kw_args)
        elif op == "pad":
            x = upfirdn2d.upfirdn2d(x=x, f=None, **kw_args)
        else:
            transpose = kw_args.pop("transpose", False)
            if transpose:
                assert groups == 1, "grouped (fused-modconv) transposed conv is not on the LayoutDETR path"
                x = _conv(x, w.transpose(0, 1), groups=groups, transpose=True, flip_weight=not flip_weight, **kw_args)
            else:
                x = _conv(x, w, groups=groups, flip_weight=flip_weight, **kw_args)
    return x
