"""Tensor -> raw-pointer wrappers over the C-ABI in include/layoutdetr_sm100.h.

PyTorch is used here only for device memory (caching allocator) and the current stream; every
function launches hand-written sm_100a kernels from liblayoutdetr_sm100.so.  No fallbacks.
"""
import ctypes
import os
from ctypes import c_void_p, c_int, c_int32, c_int64, c_float

import torch

from ._lib import lib, check

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_LRELU, ACT_SIGMOID = 0, 1, 2, 3, 4


def dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError("unsupported dtype %s (fp32 / bf16 only)" % t.dtype)


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return c_void_p(0 if t is None else t.data_ptr())


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("layoutdetr_b200 kernels need CUDA tensors (no CPU fallback); got %s" % t.device)


class _Operand(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("ld", c_int64), ("sb1", c_int64), ("sb2", c_int64),
                ("mn_major", c_int32), ("_pad", c_int32)]


class _GemmDesc(ctypes.Structure):
    _fields_ = [("M", c_int32), ("N", c_int32), ("K", c_int32), ("nb1", c_int32), ("nb2", c_int32),
                ("act", c_int32), ("accumulate", c_int32), ("split_k", c_int32),
                ("d_dtype", c_int32), ("r_dtype", c_int32), ("block_n", c_int32), ("_pad", c_int32),
                ("alpha", c_float), ("post_gain", c_float),
                ("A", _Operand), ("B", _Operand),
                ("D", c_void_p), ("ldd", c_int64), ("d_sb1", c_int64), ("d_sb2", c_int64),
                ("aux", c_void_p),
                ("R", c_void_p), ("ldr", c_int64), ("r_sb1", c_int64), ("r_sb2", c_int64),
                ("col_scale", c_void_p), ("col_bias", c_void_p), ("col_sb1", c_int64), ("col_sb2", c_int64),
                ("alpha_dev", c_void_p),
                ("softmax", c_int32), ("causal", c_int32), ("mask_value", c_float), ("_pad2", c_int32), ("key_mask", c_void_p)]


class _ConvGeom(ctypes.Structure):
    _fields_ = [("img", c_void_p), ("mode", c_int32), ("B", c_int32), ("H", c_int32), ("W", c_int32), ("C", c_int32),
                ("Ho", c_int32), ("Wo", c_int32), ("KH", c_int32), ("KW", c_int32), ("stride", c_int32), ("pad", c_int32),
                ("_pad", c_int32)]


def conv_box_ok(Ho, Wo, stride, pix):
    """Can `pix` consecutive output pixels be loaded as ONE TMA box (whole rows / whole images)?  Mirrors make_conv_map."""
    P = Ho * Wo
    if Wo >= pix:
        ok, bw, bh = Wo % pix == 0, pix, 1
    else:
        if pix % Wo:
            return False
        bw = Wo
        if P >= pix:
            ok, bh = P % pix == 0, pix // Wo
        else:
            ok, bh = pix % P == 0, Ho
    return bool(ok and bw * stride <= 256 and bh * stride <= 256)


class Op:
    """A GEMM operand view: base tensor (bf16) + element offset, leading dim, batch strides, major-ness."""
    __slots__ = ("t", "off", "ld", "sb1", "sb2", "mn")

    def __init__(self, t, ld, off=0, sb1=0, sb2=0, mn=False):
        self.t, self.off, self.ld, self.sb1, self.sb2, self.mn = t, off, ld, sb1, sb2, mn

    def fill(self, o):
        if self.t.dtype != torch.bfloat16:
            raise TypeError("GEMM operands must be bf16")
        o.ptr = self.t.data_ptr() + 2 * self.off
        o.ld, o.sb1, o.sb2, o.mn_major = self.ld, self.sb1, self.sb2, 1 if self.mn else 0


class Out:
    """Output / residual view: tensor + element offset, leading dim, batch strides."""
    __slots__ = ("t", "off", "ld", "sb1", "sb2")

    def __init__(self, t, ld, off=0, sb1=0, sb2=0):
        self.t, self.off, self.ld, self.sb1, self.sb2 = t, off, ld, sb1, sb2

    def ptr(self):
        return self.t.data_ptr() + self.t.element_size() * self.off


GEMM_LOG = None        # tools/lane_trace.py sets this to a list: one (M, N, K, batches, a_mn, b_mn, split_k, caller) per launch


def gemm(M, N, K, A, B, D, nb1=1, nb2=1, alpha=1.0, act=ACT_NONE, post_gain=1.0, accumulate=0, split_k=1,
         R=None, col_scale=None, col_bias=None, col_sb1=0, col_sb2=0, aux=None, block_n=0, alpha_dev=None, softmax=None, conv=None):
    """D = epilogue(A @ B^T) on tcgen05 tensor cores; see ld_gemm_bf16 in the header for semantics.
    conv = dict(img, mode, B, H, W, C, Ho, Wo, KH, KW, stride, pad): implicit-GEMM convolution (ld_conv_gemm_bf16) — the operand the
    mode replaces (A for mode 1, B for mode 2) is ignored and may be any bf16 tensor."""
    _cuda(A.t, B.t, D.t)
    if GEMM_LOG is not None:
        import sys
        f = sys._getframe(1)
        names = []
        while f is not None and len(names) < 4:
            if f.f_code.co_name not in ("gemm", "linear", "apply", "_dgrad", "_wgrad") or len(names) == 0:
                q = f.f_locals.get("ctx", None)
                cls = type(f.f_locals["self"]).__name__ + "." if "self" in f.f_locals else ""
                names.append(cls + f.f_code.co_name)
            f = f.f_back
        GEMM_LOG.append((M, N, K, nb1 * nb2, int(A.mn), int(B.mn), split_k, "<".join(names),
                         dict(nb1=nb1, nb2=nb2, alpha=alpha, act=act, post_gain=post_gain, accumulate=accumulate, d_dtype=str(D.t.dtype),
                              ldd=D.ld, d_sb=(D.sb1, D.sb2), a_ld=A.ld, a_sb=(A.sb1, A.sb2), b_ld=B.ld, b_sb=(B.sb1, B.sb2),
                              r=None if R is None else (str(R.t.dtype), R.ld), cs=col_scale is not None, cb=col_bias is not None,
                              aux=aux is not None, block_n=block_n, alpha_dev=alpha_dev is not None, softmax=softmax is not None,
                              conv=None if conv is None else conv["mode"])))
    d = _GemmDesc()
    d.M, d.N, d.K, d.nb1, d.nb2 = M, N, K, nb1, nb2
    d.act, d.accumulate, d.split_k = act, accumulate, split_k
    d.d_dtype = dt(D.t)
    d.block_n = block_n
    d.alpha, d.post_gain = alpha, post_gain
    A.fill(d.A)
    B.fill(d.B)
    d.D, d.ldd, d.d_sb1, d.d_sb2 = D.ptr(), D.ld, D.sb1, D.sb2
    if aux is not None:
        if aux.dtype != D.t.dtype:
            raise TypeError("aux must have the dtype of D")
        d.aux = aux.data_ptr() + aux.element_size() * D.off
    if R is not None:
        d.R, d.r_dtype, d.ldr, d.r_sb1, d.r_sb2 = R.ptr(), dt(R.t), R.ld, R.sb1, R.sb2
    if col_scale is not None:
        assert col_scale.dtype == torch.float32
        d.col_scale = col_scale.data_ptr()
    if col_bias is not None:
        assert col_bias.dtype == torch.float32
        d.col_bias = col_bias.data_ptr()
    d.col_sb1, d.col_sb2 = col_sb1, col_sb2
    if alpha_dev is not None:
        assert alpha_dev.dtype == torch.float32
        d.alpha_dev = alpha_dev.data_ptr()
    if softmax is not None:            # dict(key_mask=uint8 [nb1, N] or None, mask_inf=bool, causal=bool)
        d.softmax = 1
        d.causal = 1 if softmax.get("causal") else 0
        d.mask_value = float("-inf") if softmax.get("mask_inf") else -10000.0
        km = softmax.get("key_mask")
        if km is not None:
            assert km.dtype == torch.uint8 and km.is_contiguous()
            d.key_mask = km.data_ptr()
    if conv is not None:
        img = conv["img"]
        _cuda(img)
        if img.dtype != torch.bfloat16 or not img.is_contiguous():
            raise TypeError("conv gemm: the image must be a contiguous bf16 NHWC tensor")
        g = _ConvGeom()
        g.img, g.mode = img.data_ptr(), conv["mode"]
        g.B, g.H, g.W, g.C = conv["B"], conv["H"], conv["W"], conv["C"]
        g.Ho, g.Wo, g.KH, g.KW, g.stride, g.pad = conv["Ho"], conv["Wo"], conv["KH"], conv["KW"], conv["stride"], conv["pad"]
        check(lib().ld_conv_gemm_bf16(ctypes.byref(d), ctypes.byref(g), _stream()), "ld_conv_gemm_bf16")
        return
    check(lib().ld_gemm_bf16(ctypes.byref(d), _stream()), "ld_gemm_bf16")


def linear(x, w, bias=None, act=ACT_NONE, residual=None, out_dtype=torch.bfloat16, out=None, aux=None,
           alpha=1.0, post_gain=1.0, col_scale=None):
    """y[M, N] = act(x[M, K] @ w[N, K]^T * col_scale + bias + residual).  x, w bf16 row-major (K % 8 == 0)."""
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K and x.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    R = None
    if residual is not None:
        assert residual.shape == (M, N) and residual.stride(1) == 1
        R = Out(residual, residual.stride(0))
    gemm(M, N, K, Op(x, x.stride(0)), Op(w, w.stride(0)), Out(out, out.stride(0)), act=act, R=R,
         col_bias=bias, col_scale=col_scale, aux=aux, alpha=alpha, post_gain=post_gain)
    return out


# ------------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, out_bf16=True, out_f32=False, save_stats=False, residual=None):
    """LayerNorm(x [+ residual]) over the last dim; residual (bf16, same shape) is added in fp32."""
    _cuda(x, gamma, beta)
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.shape == x.shape and residual.stride(1) == 1
    rows, C = x.shape
    assert x.stride(1) == 1
    y16 = torch.empty((rows, C), dtype=torch.bfloat16, device=x.device) if out_bf16 else None
    y32 = torch.empty((rows, C), dtype=torch.float32, device=x.device) if out_f32 else None
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    check(lib().ld_layernorm_res_fwd(_p(x), c_int(dt(x)), c_int64(x.stride(0)), _p(residual),
                                     c_int64(residual.stride(0) if residual is not None else 0), _p(gamma), _p(beta), _p(y16), _p(y32),
                                     c_int64(C), _p(mean), _p(rstd), c_int(rows), c_int(C), c_float(eps), _stream()),
          "ld_layernorm_res_fwd")
    return y16, y32, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma=None, dbeta=None, out_dtype=torch.bfloat16):
    _cuda(dy, x, mean, rstd, gamma)
    rows, C = x.shape
    dx = torch.empty((rows, C), dtype=out_dtype, device=x.device)
    dx16 = dx if out_dtype == torch.bfloat16 else None
    dx32 = dx if out_dtype == torch.float32 else None
    check(lib().ld_layernorm_bwd(_p(dy), c_int(dt(dy)), c_int64(dy.stride(0)), _p(x), c_int(dt(x)), c_int64(x.stride(0)),
                                 _p(mean), _p(rstd), _p(gamma), _p(dx16), _p(dx32), c_int64(C), _p(dgamma), _p(dbeta),
                                 c_int(rows), c_int(C), _stream()), "ld_layernorm_bwd")
    return dx


def embed_ln_fwd(ids, word, pos, gamma, beta, T, eps, save=False):
    _cuda(ids, word, pos)
    rows = ids.numel()
    C = word.shape[1]
    y = torch.empty((rows, C), dtype=torch.bfloat16, device=word.device)
    pre = torch.empty((rows, C), dtype=torch.float32, device=word.device) if save else None
    mean = torch.empty(rows, dtype=torch.float32, device=word.device) if save else None
    rstd = torch.empty(rows, dtype=torch.float32, device=word.device) if save else None
    check(lib().ld_embed_ln_fwd(_p(ids), _p(word), _p(pos), _p(gamma), _p(beta), _p(y), _p(pre), _p(mean), _p(rstd),
                                c_int(rows), c_int(T), c_int(C), c_float(eps), _stream()), "ld_embed_ln_fwd")
    return y, pre, mean, rstd


def embed_bwd(ids, dpre, dword, dpos, T, pad_id):
    rows, C = dpre.shape
    check(lib().ld_embed_bwd(_p(ids), _p(dpre), _p(dword), _p(dpos), c_int64(rows), c_int(T), c_int(C), c_int64(pad_id),
                             _stream()), "ld_embed_bwd")


def _rng_ptr(dropout_p, device):
    if dropout_p <= 0.0:
        return c_void_p(0)
    from . import rng
    return c_void_p(rng.state(device).data_ptr())


def attention_fwd(q_t, q_off, k_t, k_off, v_t, v_off, B, H, Lq, Lk, d, scale, key_mask=None, mask_inf=False, causal=False,
                  lse_out=None, dropout_p=0.0, rng_site=0):
    """Fused attention forward. q_t/k_t/v_t: bf16 [B*L, ld] buffers (may alias), head h of q at columns q_off + h*d.
    Returns O bf16 [B*Lq, H*d]; fills lse_out [B*H, Lq] (log2-domain log-sum-exp, for attention_bwd) if given."""
    _cuda(q_t, k_t, v_t)
    O = torch.empty((B * Lq, H * d), dtype=torch.bfloat16, device=q_t.device)
    check(lib().ld_attention_fwd(c_void_p(q_t.data_ptr() + 2 * q_off), c_int64(q_t.stride(0)),
                                 c_void_p(k_t.data_ptr() + 2 * k_off), c_int64(k_t.stride(0)),
                                 c_void_p(v_t.data_ptr() + 2 * v_off), c_int64(v_t.stride(0)),
                                 _p(O), c_int64(H * d), _p(lse_out),
                                 c_int(B), c_int(H), c_int(Lq), c_int(Lk), c_int(d), c_float(scale), _p(key_mask),
                                 c_int(1 if mask_inf else 0), c_int(1 if causal else 0), c_float(dropout_p),
                                 _rng_ptr(dropout_p, q_t.device), ctypes.c_uint32(rng_site), _stream()), "ld_attention_fwd")
    return O


def attention_bwd(q_t, q_off, k_t, k_off, v_t, v_off, O, dO, lse, dq_t, Pd, dS, B, H, Lq, Lk, d, scale, key_mask=None,
                  mask_inf=False, causal=False, dropout_p=0.0, rng_site=0):
    """Fused attention backward: writes dQ into dq_t (same layout / offset as q_t), dropout(P) into Pd and dS into dS
    (both bf16 [B*H, Lq, pad8(Lk)]) for the transposed dV / dK products."""
    _cuda(q_t, k_t, v_t, O, dO, lse, dq_t, Pd, dS)
    assert O.stride(0) == dO.stride(0) and Pd.stride(1) == dS.stride(1) and Pd.is_contiguous() and dS.is_contiguous()
    check(lib().ld_attention_bwd(c_void_p(q_t.data_ptr() + 2 * q_off), c_int64(q_t.stride(0)),
                                 c_void_p(k_t.data_ptr() + 2 * k_off), c_int64(k_t.stride(0)),
                                 c_void_p(v_t.data_ptr() + 2 * v_off), c_int64(v_t.stride(0)),
                                 _p(O), _p(dO), c_int64(O.stride(0)), _p(lse),
                                 c_void_p(dq_t.data_ptr() + 2 * q_off), c_int64(dq_t.stride(0)), _p(Pd), _p(dS), c_int64(Pd.stride(1)),
                                 c_int(B), c_int(H), c_int(Lq), c_int(Lk), c_int(d), c_float(scale), _p(key_mask),
                                 c_int(1 if mask_inf else 0), c_int(1 if causal else 0), c_float(dropout_p),
                                 _rng_ptr(dropout_p, q_t.device), ctypes.c_uint32(rng_site), _stream()), "ld_attention_bwd")


def dropout(x, dropout_p, rng_site, out=None):
    """y = keep ? x / (1 - p) : 0 with the Philox mask of `rng_site` (contiguous bf16 / fp32, numel % 8 == 0)."""
    _cuda(x)
    assert x.is_contiguous() and x.numel() % 8 == 0
    y = torch.empty_like(x) if out is None else out
    check(lib().ld_dropout(_p(x), _p(y), c_int(dt(x)), c_int64(x.numel()), c_float(dropout_p), _rng_ptr(dropout_p, x.device),
                           ctypes.c_uint32(rng_site), _stream()), "ld_dropout")
    return y


def layernorm_res_dropout_fwd(x, residual, gamma, beta, eps, dropout_p, rng_site, save=False):
    """y = LayerNorm(dropout(x) + residual) (bf16 in / out); with `save` also returns dropout(x) + residual (fp32), mean, rstd."""
    _cuda(x, gamma, beta)
    rows, C = x.shape
    assert x.dtype == torch.bfloat16 and x.stride(1) == 1
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.shape == x.shape and residual.stride(1) == 1
    y = torch.empty((rows, C), dtype=torch.bfloat16, device=x.device)
    pre = torch.empty((rows, C), dtype=torch.float32, device=x.device) if save else None
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if save else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save else None
    check(lib().ld_layernorm_res_dropout_fwd(_p(x), c_int64(x.stride(0)), _p(residual),
                                             c_int64(residual.stride(0) if residual is not None else 0), _p(gamma), _p(beta),
                                             _p(y), c_int64(C), _p(pre), c_int64(C), _p(mean), _p(rstd), c_int(rows), c_int(C),
                                             c_float(eps), c_float(dropout_p), _rng_ptr(dropout_p, x.device),
                                             ctypes.c_uint32(rng_site), _stream()), "ld_layernorm_res_dropout_fwd")
    return y, pre, mean, rstd


def softmax_fwd(S, P, nb1, nb2, rows, cols, scale, key_mask=None, mask_inf=False, causal=False):
    """S fp32 [nb1*nb2, rows, ldS] -> P bf16 [nb1*nb2, rows, ldP] (both contiguous 3-D tensors)."""
    _cuda(S, P)
    check(lib().ld_softmax_fwd(_p(S), c_int64(S.stride(1)), c_int64(S.stride(0)), _p(P), c_int64(P.stride(1)),
                               c_int64(P.stride(0)), c_int(nb1), c_int(nb2), c_int(rows), c_int(cols), c_float(scale),
                               _p(key_mask), c_int(1 if mask_inf else 0), c_int(1 if causal else 0), _stream()),
          "ld_softmax_fwd")


def softmax_bwd(P, dP, dS, nb, rows, cols, scale):
    check(lib().ld_softmax_bwd(_p(P), c_int64(P.stride(1)), c_int64(P.stride(0)), _p(dP), c_int64(dP.stride(1)),
                               c_int64(dP.stride(0)), _p(dS), c_int64(dS.stride(1)), c_int64(dS.stride(0)),
                               c_int(nb), c_int(rows), c_int(cols), c_float(scale), _stream()), "ld_softmax_bwd")


def cross_entropy(logits, labels, label_smoothing=0.0, ignore_index=-100, want_loss=True, dlogits=None, grad_scale=1.0,
                  grad_scale_dev=None):
    _cuda(logits, labels)
    rows, V = logits.shape
    loss_rows = torch.empty(rows, dtype=torch.float32, device=logits.device) if want_loss else None
    g_dt = dt(dlogits) if dlogits is not None else BF16
    ldg = dlogits.stride(0) if dlogits is not None else 0
    check(lib().ld_cross_entropy(_p(logits), c_int(dt(logits)), c_int64(logits.stride(0)), _p(labels), _p(loss_rows),
                                 _p(dlogits), c_int(g_dt), c_int64(ldg), c_int64(rows), c_int(V),
                                 c_float(label_smoothing), c_int64(ignore_index), c_float(grad_scale), _p(grad_scale_dev), _stream()),
          "ld_cross_entropy")
    return loss_rows


# ------------------------------------------------------------------------------------------------
def bias_act_raw(x, b, xref, yref, dy, y, grad, act_idx, alpha, gain, clamp, sizeB, stepB):
    _cuda(x, y)
    check(lib().ld_bias_act(_p(x), _p(b), _p(xref), _p(yref), _p(dy), _p(y), c_int(dt(x)), c_int(grad), c_int(act_idx),
                            c_float(alpha), c_float(gain), c_float(clamp), c_int64(x.numel()), c_int(sizeB),
                            c_int64(stepB), _stream()), "ld_bias_act")


def fma_f32(a, b, c):
    """a * b + c (fp32): a contiguous (<= 4-D), b / c broadcastable to a."""
    _cuda(a, b, c)
    shape = list(a.shape)
    while len(shape) < 4:
        shape = [1] + shape
    a4 = a.contiguous().view(shape)
    be = b.broadcast_to(a.shape)
    ce = c.broadcast_to(a.shape)
    pad = 4 - be.ndim
    bs = (c_int64 * 4)(*([0] * pad + list(be.stride())))
    cs = (c_int64 * 4)(*([0] * pad + list(ce.stride())))
    sh = (c_int64 * 4)(*shape)
    y = torch.empty_like(a4)
    check(lib().ld_fma_f32(_p(a4), _p(be), _p(ce), _p(y), sh, bs, cs, _stream()), "ld_fma_f32")
    return y.view(a.shape)


def cast_pad(src, dst_dtype, cols_dst=None, cols_src=None):
    """2-D cast with optional zero padding (cols_dst > cols) or truncation (cols_src < src.shape[1]) of the
    inner dim (padding to a multiple of 8 is what TMA needs)."""
    _cuda(src)
    rows, cols = src.shape
    assert src.stride(1) == 1
    cols = cols if cols_src is None else cols_src
    cols_dst = cols if cols_dst is None else cols_dst
    dst = torch.empty((rows, cols_dst), dtype=dst_dtype, device=src.device)
    check(lib().ld_cast_pad(_p(src), c_int(dt(src)), c_int64(src.stride(0)), _p(dst), c_int(dt(dst)), c_int64(cols_dst),
                            c_int64(rows), c_int(cols), c_int(cols_dst), _stream()), "ld_cast_pad")
    return dst


def to_bf16(x):
    """Contiguous fp32 -> bf16 copy (any shape)."""
    if x.dtype == torch.bfloat16:
        return x
    x = x.contiguous()
    out = cast_pad(x.view(1, -1), torch.bfloat16)
    return out.view(x.shape)


def to_f32(x):
    if x.dtype == torch.float32:
        return x
    x = x.contiguous()
    return cast_pad(x.view(1, -1), torch.float32).view(x.shape)


def axpby_bcast(a, b, out_dtype=torch.bfloat16, alpha=1.0, beta=1.0):
    """out = alpha * a + beta * b, with b broadcast periodically over a (a.numel() % b.numel() == 0)."""
    _cuda(a, b)
    a = a.contiguous(); b = b.contiguous()
    assert a.numel() % b.numel() == 0
    out = torch.empty(a.shape, dtype=out_dtype, device=a.device)
    check(lib().ld_axpby_bcast(_p(a), c_int(dt(a)), _p(b), c_int(dt(b)), _p(out), c_int(dt(out)), c_int64(a.numel()),
                               c_int64(b.numel()), c_float(alpha), c_float(beta), _stream()), "ld_axpby_bcast")
    return out


def act_bwd(dy, ref, act, gain=1.0):
    dy = dy.contiguous(); ref = ref.contiguous()
    dx = torch.empty_like(dy)
    check(lib().ld_act_bwd(_p(dy), c_int(dt(dy)), _p(ref), c_int(dt(ref)), _p(dx), c_int(dt(dx)), c_int64(dy.numel()),
                           c_int(act), c_float(gain), _stream()), "ld_act_bwd")
    return dx


def act_bwd_colscale(dy, ref, cs, act, want_dx):
    """(dx, dx * cs[col]) with dx = dy * act'(ref); bf16 [rows, cols], cols % 8 == 0."""
    dy = dy.contiguous()
    dxs = torch.empty_like(dy)
    dx = torch.empty_like(dy) if want_dx else None
    check(lib().ld_act_bwd_colscale(_p(dy), _p(ref), _p(dx), _p(dxs), _p(cs), c_int64(dy.numel()), c_int(dy.shape[1]), c_int(act),
                                    _stream()), "ld_act_bwd_colscale")
    return dx, dxs


def act_fwd(x, act, gain=1.0):
    x = x.contiguous()
    y = torch.empty_like(x)
    check(lib().ld_act_fwd_bf16(_p(x), _p(y), c_int64(x.numel()), c_int(act), c_float(gain), _stream()), "ld_act_fwd_bf16")
    return y


def colsum_accum(x, out):
    """out[c] += sum_r x[r, c]  (bias gradients)."""
    rows, cols = x.shape
    check(lib().ld_colsum_accum(_p(x), c_int(dt(x)), c_int64(x.stride(0)), _p(out), c_int64(rows), c_int(cols), _stream()),
          "ld_colsum_accum")


def scale_channels(x, s, out_dtype, per_sample, C):
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    check(lib().ld_scale_channels(_p(x), c_int(dt(x)), _p(s), _p(y), c_int(dt(y)), c_int64(x.numel()), c_int64(per_sample),
                                  c_int(C), _stream()), "ld_scale_channels")
    return y


# ------------------------------------------------------------------------------------------------
# convolution support (channels-last bf16)
def conv_out_size(n, k, stride, pad):
    return (n + 2 * pad - k) // stride + 1


def im2col(x, B, H, W, C, KH, KW, stride, pad):
    """x: bf16 NHWC (any shape with B*H*W*C elements, contiguous) -> cols bf16 [B*Ho*Wo, pad8(KH*KW*C)]."""
    _cuda(x)
    Ho, Wo = conv_out_size(H, KH, stride, pad), conv_out_size(W, KW, stride, pad)
    Kp = (KH * KW * C + 7) // 8 * 8
    cols = torch.empty((B * Ho * Wo, Kp), dtype=torch.bfloat16, device=x.device)
    check(lib().ld_im2col_nhwc(_p(x), _p(cols), c_int(B), c_int(H), c_int(W), c_int(C), c_int(Ho), c_int(Wo),
                               c_int(KH), c_int(KW), c_int(stride), c_int(pad), c_int(Kp), _stream()), "ld_im2col_nhwc")
    return cols, Ho, Wo


def col2im(cols, B, H, W, C, Ho, Wo, KH, KW, stride, pad):
    """Gather-form inverse of im2col: cols [B*Ho*Wo, Kp] -> bf16 [B*H*W, C]."""
    _cuda(cols)
    out = torch.empty((B * H * W, C), dtype=torch.bfloat16, device=cols.device)
    check(lib().ld_col2im_nhwc(_p(cols), _p(out), c_int(B), c_int(H), c_int(W), c_int(C), c_int(Ho), c_int(Wo),
                               c_int(KH), c_int(KW), c_int(stride), c_int(pad), c_int(cols.stride(0)), _stream()),
          "ld_col2im_nhwc")
    return out


def maxpool3s2_fwd(x, B, H, W, C, save_argmax):
    Ho, Wo = conv_out_size(H, 3, 2, 1), conv_out_size(W, 3, 2, 1)
    y = torch.empty((B * Ho * Wo, C), dtype=torch.bfloat16, device=x.device)
    arg = torch.empty((B * Ho * Wo, C), dtype=torch.uint8, device=x.device) if save_argmax else None
    check(lib().ld_maxpool3s2_fwd(_p(x), _p(y), _p(arg), c_int(B), c_int(H), c_int(W), c_int(C), _stream()), "ld_maxpool3s2_fwd")
    return y, arg, Ho, Wo


def maxpool3s2_bwd(dy, arg, B, H, W, C):
    dx = torch.empty((B * H * W, C), dtype=torch.bfloat16, device=dy.device)
    check(lib().ld_maxpool3s2_bwd(_p(dy), _p(arg), _p(dx), c_int(B), c_int(H), c_int(W), c_int(C), _stream()), "ld_maxpool3s2_bwd")
    return dx


def nchw_to_nhwc(x, out_dtype):
    """[B, C, H, W] contiguous -> [B*H*W, C] contiguous (dtype conversion fused)."""
    _cuda(x)
    x = x.contiguous()
    B, C, H, W = x.shape
    out = torch.empty((B * H * W, C), dtype=out_dtype, device=x.device)
    check(lib().ld_layout_convert(_p(x), c_int(dt(x)), _p(out), c_int(dt(out)), c_int(B), c_int(C), c_int64(H * W), c_int(0),
                                  _stream()), "ld_layout_convert")
    return out


def nhwc_to_nchw(x, B, C, H, W, out_dtype):
    """[B*H*W, C] contiguous -> [B, C, H, W] contiguous."""
    _cuda(x)
    x = x.contiguous()
    out = torch.empty((B, C, H, W), dtype=out_dtype, device=x.device)
    check(lib().ld_layout_convert(_p(x), c_int(dt(x)), _p(out), c_int(dt(out)), c_int(B), c_int(C), c_int64(H * W), c_int(1),
                                  _stream()), "ld_layout_convert")
    return out


def upfirdn2d_raw(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain, channels_last_out=None):
    """x: 4-D [N, C, H, W] view with arbitrary strides (NCHW or channels-last), fp32 or bf16; f: fp32 [fh, fw]."""
    _cuda(x, f)
    assert x.ndim == 4 and f.ndim == 2 and f.dtype == torch.float32
    f = f.contiguous()
    N, C, H, W = x.shape
    fh, fw = f.shape
    outH = (H * upy + pady0 + pady1 - fh + downy) // downy
    outW = (W * upx + padx0 + padx1 - fw + downx) // downx
    cl = (x.stride(1) == 1 and C > 1) if channels_last_out is None else channels_last_out
    y = torch.empty((N, C, outH, outW), dtype=x.dtype, device=x.device,
                    memory_format=torch.channels_last if cl else torch.contiguous_format)
    xs = (c_int64 * 4)(*x.stride())
    ys = (c_int64 * 4)(*y.stride())
    check(lib().ld_upfirdn2d(_p(x), _p(y), c_int(dt(x)), _p(f), c_int(fh), c_int(fw), c_int(N), c_int(C), c_int(H), c_int(W),
                             c_int(outH), c_int(outW), xs, ys, c_int(upx), c_int(upy), c_int(downx), c_int(downy),
                             c_int(padx0), c_int(padx1), c_int(pady0), c_int(pady1), c_int(1 if flip_filter else 0),
                             c_float(gain), _stream()), "ld_upfirdn2d")
    return y


def demod_bias_act_fwd(x, d, bias, B, pixels, C, act, gain):
    y = torch.empty((B * pixels, C), dtype=torch.bfloat16, device=x.device)
    check(lib().ld_demod_bias_act_fwd(_p(x), c_int(dt(x)), _p(d), _p(bias), _p(y), c_int(B), c_int64(pixels), c_int(C),
                                      c_int(act), c_float(gain), _stream()), "ld_demod_bias_act_fwd")
    return y


DETERMINISTIC_STYLE_SUMS = os.environ.get("LD_DETERMINISTIC_STYLE_SUMS", "1") != "0"


def _style_ws(B, pixels, C, mult, device):
    lib().ld_style_reduce_ws_floats.restype = c_int64
    n = int(lib().ld_style_reduce_ws_floats(c_int(B), c_int64(pixels), c_int(C)))
    return torch.empty(mult * n, dtype=torch.float32, device=device) if n > 0 else None


def demod_bias_act_bwd(dy, y, x, d, dd, dbias, B, pixels, C, act, gain):
    dx = torch.empty((B * pixels, C), dtype=torch.bfloat16, device=x.device)
    ws = _style_ws(B, pixels, C, 2, x.device) if (DETERMINISTIC_STYLE_SUMS and x.dtype == torch.bfloat16) else None
    if ws is not None:          # fixed-order sums (no fp32 atomics): run-to-run bit-identical style / bias gradients
        check(lib().ld_demod_bias_act_bwd_ws(_p(dy), _p(y), _p(x), _p(d), _p(dx), _p(dd), _p(dbias), _p(ws), c_int64(ws.numel()), c_int(B),
                                             c_int64(pixels), c_int(C), c_int(act), c_float(gain), _stream()), "ld_demod_bias_act_bwd_ws")
        return dx
    check(lib().ld_demod_bias_act_bwd(_p(dy), _p(y), _p(x), c_int(dt(x)), _p(d), _p(dx), _p(dd), _p(dbias), c_int(B),
                                      c_int64(pixels), c_int(C), c_int(act), c_float(gain), _stream()), "ld_demod_bias_act_bwd")
    return dx


def channel_dot(a, g, B, pixels, C):
    out = torch.zeros((B, C), dtype=torch.float32, device=a.device)
    ws = _style_ws(B, pixels, C, 1, a.device) if (DETERMINISTIC_STYLE_SUMS and a.dtype == torch.bfloat16) else None
    if ws is not None:
        check(lib().ld_channel_dot_ws(_p(a), _p(g), _p(out), _p(ws), c_int64(ws.numel()), c_int(B), c_int64(pixels), c_int(C), _stream()),
              "ld_channel_dot_ws")
        return out
    check(lib().ld_channel_dot(_p(a), c_int(dt(a)), _p(g), _p(out), c_int(B), c_int64(pixels), c_int(C), _stream()),
          "ld_channel_dot")
    return out


def adam_flat(p, g, m, v, p16, lr, beta1, beta2, eps, step, grad_scale=1.0, hyper_dev=None):
    """Fused nan_to_num + Adam + bf16-shadow refresh over flat fp32 storage (numel % 4 == 0)."""
    _cuda(p, g, v)
    check(lib().ld_adam_flat(_p(p), _p(g), _p(m), _p(v), _p(p16), c_int64(p.numel()), c_float(lr), c_float(beta1),
                             c_float(beta2), c_float(eps), c_int(step), c_float(grad_scale), _p(hyper_dev), _stream()), "ld_adam_flat")


def ema_flat(p_ema, p, p_ema16, beta):
    _cuda(p_ema, p)
    check(lib().ld_ema_flat(_p(p_ema), _p(p), _p(p_ema16), c_int64(p.numel()), c_float(beta), _stream()), "ld_ema_flat")


def lsap(cost, maximize=False):
    """Batched linear-sum-assignment on the GPU. cost: fp64 CUDA tensor [P, nr, nc] (or [nr, nc]);
    returns (rows, cols) int64 [P, min(nr, nc)] exactly as scipy.optimize.linear_sum_assignment orders them."""
    _cuda(cost)
    single = cost.ndim == 2
    c = cost.reshape(-1, cost.shape[-2], cost.shape[-1]).to(torch.float64).contiguous()
    P, nr, nc = c.shape
    k = min(nr, nc)
    rows = torch.empty((P, k), dtype=torch.int64, device=c.device)
    cols = torch.empty((P, k), dtype=torch.int64, device=c.device)
    status = torch.empty(P, dtype=torch.int32, device=c.device)
    check(lib().ld_lsap(_p(c), c_int(nr), c_int(nc), c_int(1 if maximize else 0), c_int64(P), _p(rows), _p(cols), _p(status),
                        _stream()), "ld_lsap")
    if single:
        return rows[0], cols[0], status[0]
    return rows, cols, status


# ------------------------------------------------------------------------------------------------
# layout box losses (csrc/box_loss.cu)
def layout_losses(bbox, valid, want_jac):
    """bbox [B, N, 4] fp32, valid [B, N] bool/uint8 -> (overlap [B], alignment [B], J_overlap, J_alignment [B, N, 4] or None)."""
    _cuda(bbox, valid)
    B, N, _ = bbox.shape
    bbox = bbox.contiguous()
    v8 = valid.contiguous().view(torch.uint8) if valid.dtype == torch.bool else valid.to(torch.uint8).contiguous()
    ov = torch.empty(B, dtype=torch.float32, device=bbox.device)
    al = torch.empty(B, dtype=torch.float32, device=bbox.device)
    j_ov = torch.empty_like(bbox) if want_jac else None
    j_al = torch.empty_like(bbox) if want_jac else None
    check(lib().ld_layout_losses(_p(bbox), _p(v8), c_int64(B), c_int(N), _p(ov), _p(al), _p(j_ov), _p(j_al), _stream()), "ld_layout_losses")
    return ov, al, j_ov, j_al


def giou_loss(fake, real, want_jac):
    """fake, real [M, 4] fp32 -> (mean(1 - GIoU) as a 0-dim tensor, d loss / d fake [M, 4] or None)."""
    _cuda(fake, real)
    fake, real = fake.contiguous(), real.contiguous()
    loss = torch.empty(1, dtype=torch.float32, device=fake.device)
    jac = torch.empty_like(fake) if want_jac else None
    check(lib().ld_giou_loss(_p(fake), _p(real), c_int64(fake.shape[0]), _p(loss), _p(jac), _stream()), "ld_giou_loss")
    return loss[0], jac


def rows_scale(J, g, out=None, accumulate=False):
    """out (+)= J * g broadcast over J's trailing elements (g: one value per leading row group, or a scalar)."""
    _cuda(J, g)
    g = g.reshape(-1).contiguous()
    per = J.numel() // g.numel()
    if out is None:
        out = torch.empty_like(J)
    check(lib().ld_rows_scale(_p(J), _p(g), _p(out), c_int64(J.numel()), c_int64(per), c_int(1 if accumulate else 0), _stream()), "ld_rows_scale")
    return out


def layout_pair_metrics(real, fake, valid):
    """real, fake [B, N, 4] fp32, valid [B, N] -> (layout-wise IoU [B], layout-wise DocSim [B])."""
    _cuda(real, fake, valid)
    B, N, _ = real.shape
    real, fake = real.float().contiguous(), fake.float().contiguous()
    v8 = valid.contiguous().view(torch.uint8) if valid.dtype == torch.bool else valid.to(torch.uint8).contiguous()
    iou = torch.empty(B, dtype=torch.float32, device=real.device)
    doc = torch.empty(B, dtype=torch.float32, device=real.device)
    check(lib().ld_layout_pair_metrics(_p(real), _p(fake), _p(v8), c_int64(B), c_int(N), _p(iou), _p(doc), _stream()), "ld_layout_pair_metrics")
    return iou, doc


def normalize_u8_image(img_u8, mean, std):
    """uint8 [B, H, W, 3] (CUDA) -> fp32 [B, 3, H, W] = (x / 255 - mean) / std, bit-identical to the reference loader's NumPy."""
    _cuda(img_u8)
    assert img_u8.dtype == torch.uint8 and img_u8.ndim == 4 and img_u8.shape[-1] == 3
    img_u8 = img_u8.contiguous()
    B, H, W, _ = img_u8.shape
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=img_u8.device)
    m = (c_float * 3)(*[float(x) for x in mean])
    s = (c_float * 3)(*[float(x) for x in std])
    check(lib().ld_normalize_u8_image(_p(img_u8), _p(out), c_int64(B), c_int64(H), c_int64(W), m, s, _stream()), "ld_normalize_u8_image")
    return out
