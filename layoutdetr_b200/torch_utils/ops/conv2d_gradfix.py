"""Drop-in for the reference's torch_utils/ops/conv2d_gradfix.py API (`conv2d`, `conv_transpose2d`, `enabled`,
`weight_gradients_disabled`, `no_weight_gradients()`; reference :37-42) on the sm_100a conv path (im2col / col2im + tcgen05 GEMM,
bf16 operands, fp32 accumulation).  CUDA tensors only.

Arbitrary-order gradients, as the reference's op provides them for R1 / path-length regularisation (reference :103-176,
training/loss.py:132,210): the three bilinear maps of a convolution geometry

    conv   (x, w)  -> y        y  = x * w                      forward
    convT  (y, w)  -> x        dx = y (*)^T w                  data gradient   = transposed convolution
    wgrad  (x, y)  -> w        dw = sum_b x (*) y              weight gradient

form a closed set under differentiation — the backward of each is made of the other two — so every one of them is an
autograd Function whose backward calls the others through `.apply`, and autograd can differentiate through a backward pass as
often as it likes.  Each map is one GEMM around the patch-gather kernels; nothing here is a cuDNN / cuBLAS call.
"""
import collections
import contextlib

import torch
import torch.nn.functional as F

from ... import kernels as K

enabled = False
weight_gradients_disabled = False

_Geom = collections.namedtuple("_Geom", "B Cin H W Cout KH KW stride pad Ho Wo")       # x-space [B, Cin, H, W], y-space [B, Cout, Ho, Wo]


@contextlib.contextmanager
def no_weight_gradients(disable=True):
    global weight_gradients_disabled
    old = weight_gradients_disabled
    if disable:
        weight_gradients_disabled = True
    yield
    weight_gradients_disabled = old


def _pad8(n):
    return (n + 7) // 8 * 8


def _rows(t, C, Cp):
    """[B, C, H, W] (any float dtype) -> bf16 NHWC rows [B*H*W, Cp], channels zero-padded to a multiple of 8 (TMA stride rule)."""
    t = t.detach()
    if Cp != C:
        t = F.pad(t, (0, 0, 0, 0, 0, Cp - C))
    return K.nchw_to_nhwc(t if t.dtype in (torch.float32, torch.bfloat16) else t.float(), torch.bfloat16)


def _nchw(rows, B, C, Cp, H, W, dtype):
    out = K.nhwc_to_nchw(rows, B, Cp, H, W, torch.float32)
    return (out[:, :C] if Cp != C else out).to(dtype).contiguous()


def _w2d(w, Cop, Cip):
    """OIHW -> bf16 [Cop, pad8(KH*KW*Cip)], K ordered (kh, kw, ci) to match the patch matrix; channels zero-padded."""
    w = w.detach().float()
    Cout, Cin, KH, KW = w.shape
    w = F.pad(w, (0, 0, 0, 0, 0, Cip - Cin, 0, Cop - Cout))
    w2 = w.permute(0, 2, 3, 1).reshape(Cop, KH * KW * Cip).contiguous()
    return K.cast_pad(w2, torch.bfloat16, _pad8(w2.shape[1]))


def _conv_raw(x, w, g):
    Cip, Cop = _pad8(g.Cin), _pad8(g.Cout)
    xr = _rows(x, g.Cin, Cip)
    if g.KH == 1 and g.KW == 1 and g.stride == 1 and g.pad == 0:
        cols = xr
    else:
        cols, _, _ = K.im2col(xr, g.B, g.H, g.W, Cip, g.KH, g.KW, g.stride, g.pad)
    w2 = _w2d(w, Cop, Cip)
    M = g.B * g.Ho * g.Wo
    y = torch.empty((M, Cop), dtype=torch.bfloat16, device=x.device)
    K.gemm(M, Cop, cols.shape[1], K.Op(cols, cols.stride(0)), K.Op(w2, w2.stride(0)), K.Out(y, Cop))
    return _nchw(y, g.B, g.Cout, Cop, g.Ho, g.Wo, x.dtype)


def _convT_raw(y, w, g):
    Cip, Cop = _pad8(g.Cin), _pad8(g.Cout)
    yr = _rows(y, g.Cout, Cop)
    w2 = _w2d(w, Cop, Cip)
    M = g.B * g.Ho * g.Wo
    dcols = torch.empty((M, w2.shape[1]), dtype=torch.bfloat16, device=y.device)
    K.gemm(M, w2.shape[1], Cop, K.Op(yr, Cop), K.Op(w2, w2.stride(0), mn=True), K.Out(dcols, dcols.stride(0)))
    if g.KH == 1 and g.KW == 1 and g.stride == 1 and g.pad == 0:
        xr = dcols[:, :Cip].contiguous()
    else:
        xr = K.col2im(dcols, g.B, g.H, g.W, Cip, g.Ho, g.Wo, g.KH, g.KW, g.stride, g.pad)
    return _nchw(xr, g.B, g.Cin, Cip, g.H, g.W, y.dtype)


def _wgrad_raw(x, y, g, dtype):
    Cip, Cop = _pad8(g.Cin), _pad8(g.Cout)
    xr = _rows(x, g.Cin, Cip)
    yr = _rows(y, g.Cout, Cop)
    if g.KH == 1 and g.KW == 1 and g.stride == 1 and g.pad == 0:
        cols = xr
    else:
        cols, _, _ = K.im2col(xr, g.B, g.H, g.W, Cip, g.KH, g.KW, g.stride, g.pad)
    Kreal = g.KH * g.KW * Cip
    M = g.B * g.Ho * g.Wo
    dw = torch.empty((Cop, Kreal), dtype=torch.float32, device=x.device)
    K.gemm(Cop, Kreal, M, K.Op(yr, Cop, mn=True), K.Op(cols, cols.stride(0), mn=True), K.Out(dw, Kreal))
    return dw.view(Cop, g.KH, g.KW, Cip).permute(0, 3, 1, 2)[:g.Cout, :g.Cin].to(dtype).contiguous()


class _Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, g):
        ctx.g = g
        ctx.save_for_backward(x, w)
        return _conv_raw(x, w, g)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = _ConvT.apply(dy, w, ctx.g) if ctx.needs_input_grad[0] else None
        dw = _WGrad.apply(x, dy, ctx.g, w.dtype) if (ctx.needs_input_grad[1] and not weight_gradients_disabled) else None
        return dx, dw, None


class _ConvT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, w, g):
        ctx.g = g
        ctx.save_for_backward(y, w)
        return _convT_raw(y, w, g)

    @staticmethod
    def backward(ctx, dx):
        y, w = ctx.saved_tensors
        dy = _Conv.apply(dx, w, ctx.g) if ctx.needs_input_grad[0] else None
        dw = _WGrad.apply(dx, y, ctx.g, w.dtype) if (ctx.needs_input_grad[1] and not weight_gradients_disabled) else None
        return dy, dw, None


class _WGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, g, dtype):
        ctx.g = g
        ctx.save_for_backward(x, y)
        return _wgrad_raw(x, y, g, dtype)

    @staticmethod
    def backward(ctx, dw):
        x, y = ctx.saved_tensors
        dx = _ConvT.apply(y, dw, ctx.g) if ctx.needs_input_grad[0] else None
        dy = _Conv.apply(x, dw, ctx.g) if ctx.needs_input_grad[1] else None
        return dx, dy, None, None


def _square(v, what):
    a, b = (v, v) if isinstance(v, int) else tuple(v)
    if a != b:
        raise NotImplementedError("layoutdetr_b200 conv2d_gradfix: %s must be the same on both axes (got %r)" % (what, v))
    return int(a)


def _check(input, groups, dilation):
    if not input.is_cuda:
        raise RuntimeError("layoutdetr_b200 conv2d_gradfix: CUDA tensors only (no CPU fallback)")
    if groups != 1 or _square(dilation, "dilation") != 1:
        raise NotImplementedError("layoutdetr_b200 conv2d_gradfix: groups = 1 and no dilation (the LayoutDETR path runs "
                                  "fused_modconv=False, training/networks_detr.py:261)")


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    _check(input, groups, dilation)
    s, p = _square(stride, "stride"), _square(padding, "padding")
    B, Cin, H, W = input.shape
    Cout, Cin_w, KH, KW = weight.shape
    assert Cin_w == Cin
    g = _Geom(B, Cin, H, W, Cout, KH, KW, s, p, K.conv_out_size(H, KH, s, p), K.conv_out_size(W, KW, s, p))
    y = _Conv.apply(input, weight, g)
    return y if bias is None else y + bias.to(y.dtype).reshape(1, -1, 1, 1)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    """weight [C_in, C_out, kh, kw] as in torch: the adjoint of conv2d(C_out -> C_in) with that very tensor as OIHW weight."""
    _check(input, groups, dilation)
    s, p, op = _square(stride, "stride"), _square(padding, "padding"), _square(output_padding, "output_padding")
    B, Cy, Ho, Wo = input.shape
    Cy_w, Cx, KH, KW = weight.shape
    assert Cy_w == Cy
    H, W = (Ho - 1) * s - 2 * p + KH + op, (Wo - 1) * s - 2 * p + KW + op
    g = _Geom(B, Cx, H, W, Cy, KH, KW, s, p, Ho, Wo)
    assert K.conv_out_size(H, KH, s, p) == Ho and K.conv_out_size(W, KW, s, p) == Wo
    x = _ConvT.apply(input, weight, g)
    return x if bias is None else x + bias.to(x.dtype).reshape(1, -1, 1, 1)
