"""Drop-in for the reference's torch_utils/ops/conv2d_gradfix.py API (`conv2d`, `conv_transpose2d`,
`enabled`, `weight_gradients_disabled`, `no_weight_gradients()`), routed to the sm_100a conv path
(im2col / col2im + tcgen05 GEMM, bf16 operands, fp32 accumulation) for CUDA tensors.

On torch >= 1.11 the reference's wrappers are plain F.conv2d / F.conv_transpose2d calls (:53-55), so the only
contract is shape/dtype semantics; `no_weight_gradients` is honoured by detaching the weight."""
import contextlib

import torch

from ... import functional as Fn
from ... import kernels as K

enabled = False
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients(disable=True):
    global weight_gradients_disabled
    old = weight_gradients_disabled
    if disable:
        weight_gradients_disabled = True
    yield
    weight_gradients_disabled = old


class _NCHWToRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape = x.shape
        ctx.dtype = x.dtype
        return K.nchw_to_nhwc(x.float() if x.dtype not in (torch.float32, torch.bfloat16) else x, torch.bfloat16)

    @staticmethod
    def backward(ctx, g):
        B, C, H, W = ctx.shape
        return K.nhwc_to_nchw(g.contiguous(), B, C, H, W, torch.float32).to(ctx.dtype)


class _RowsToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, B, C, H, W, dtype):
        ctx.geom = (B, C, H, W)
        return K.nhwc_to_nchw(y.contiguous(), B, C, H, W, torch.float32).to(dtype)

    @staticmethod
    def backward(ctx, g):
        return K.nchw_to_nhwc(g.float(), torch.bfloat16), None, None, None, None, None


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    if not input.is_cuda:
        raise RuntimeError("layoutdetr_b200 conv2d: CUDA tensors only (no CPU fallback)")
    sh, sw = _pair(stride); ph, pw = _pair(padding); dh, dw = _pair(dilation)
    if groups != 1 or sh != sw or ph != pw or (dh, dw) != (1, 1):
        raise NotImplementedError("layoutdetr_b200 conv2d: groups=1, square stride/padding, no dilation "
                                  "(the LayoutDETR path uses fused_modconv=False, networks_detr.py:261)")
    B, Cin, H, W = input.shape
    Cout, _, KH, KW = weight.shape
    w = weight.detach() if weight_gradients_disabled else weight
    rows = _NCHWToRows.apply(input)
    y = Fn.conv2d(rows, w, None, bias.float() if bias is not None else None, None, B, H, W, sh, ph, K.ACT_NONE)
    Ho, Wo = K.conv_out_size(H, KH, sh, ph), K.conv_out_size(W, KW, sw, pw)
    return _RowsToNCHW.apply(y, B, Cout, Ho, Wo, input.dtype)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    if not input.is_cuda:
        raise RuntimeError("layoutdetr_b200 conv_transpose2d: CUDA tensors only (no CPU fallback)")
    if groups != 1 or _pair(stride) != (2, 2) or _pair(padding) != (0, 0) or _pair(output_padding) != (0, 0) \
            or _pair(dilation) != (1, 1) or tuple(weight.shape[2:]) != (3, 3):
        raise NotImplementedError("layoutdetr_b200 conv_transpose2d: only the stride-2 3x3 unpadded form used by "
                                  "conv2d_resample's up=2 branch (torch_utils/ops/conv2d_resample.py:113-130)")
    B, Cin, H, W = input.shape
    w = weight.transpose(0, 1)                      # [in, out, kh, kw] -> OIHW view expected by the kernel path
    w = w.detach() if weight_gradients_disabled else w
    rows = _NCHWToRows.apply(input)
    y = _ConvTUp2Shim.apply(rows, w.contiguous(), B, H, W)
    out = _RowsToNCHW.apply(y, B, w.shape[0], 2 * H + 1, 2 * W + 1, input.dtype)
    if bias is not None:
        out = out + bias.reshape(1, -1, 1, 1)
    return out


class _ConvTUp2Shim(torch.autograd.Function):
    """ConvTransposeUp2Fn accumulates weight gradients into `.grad` of a Parameter; the public op API must
    return them through autograd instead (weights here are arbitrary tensors)."""

    @staticmethod
    def forward(ctx, rows, w, B, H, W):
        ctx.geom = (B, H, W)
        ctx.save_for_backward(rows, w)
        with torch.no_grad():
            return Fn.ConvTransposeUp2Fn.apply(rows, w, B, H, W)

    @staticmethod
    def backward(ctx, dy):
        rows, w = ctx.saved_tensors
        B, H, W = ctx.geom
        Cout, Cin, KH, KW = w.shape
        dcols, _, _ = K.im2col(dy.contiguous(), B, 2 * H + 1, 2 * W + 1, Cout, KH, KW, 2, 0)
        wt = K.cast_pad(w.detach().permute(2, 3, 0, 1).reshape(KH * KW * Cout, Cin).contiguous(), torch.bfloat16)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((dcols.shape[0], Cin), dtype=torch.bfloat16, device=dy.device)
            K.gemm(dcols.shape[0], Cin, dcols.shape[1], K.Op(dcols, dcols.stride(0)), K.Op(wt, wt.stride(0), mn=True), K.Out(dx, Cin))
        if ctx.needs_input_grad[1]:
            tmp = torch.empty((KH * KW * Cout, Cin), dtype=torch.float32, device=dy.device)
            K.gemm(KH * KW * Cout, Cin, dcols.shape[0], K.Op(dcols, dcols.stride(0), mn=True), K.Op(rows, rows.stride(0), mn=True), K.Out(tmp, Cin))
            dw = tmp.view(KH, KW, Cout, Cin).permute(2, 3, 0, 1).contiguous()
        return dx, dw, None, None, None
