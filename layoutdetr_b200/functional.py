"""Autograd bindings of the sm_100a kernels (forward and hand-written backward).

Conventions
  * activations are bf16, row-major 2-D `[rows, features]`; fp32 only inside fused blocks
    (pre-LayerNorm sums) and for final heads / losses;
  * parameters are fp32 masters; tensor cores read bf16 shadows (engine.w_bf16);
  * weight / bias gradients are accumulated straight into `param.grad` by the backward kernels
    (GEMM epilogue accumulate / atomics) and the autograd Function returns None for them — the
    training loop only ever reads `.grad` (reference training/training_loop.py:303-312).
"""
import math
import os

import torch

from . import engine as E
from . import kernels as K
from . import rng as RNG


def _needs(p):
    return p is not None and p.requires_grad


_TRACK = [True]


class _Fn(torch.autograd.Function):
    """Base of the hand-written Functions.  `ctx.needs_input_grad` reports `requires_grad` of the inputs even under
    `torch.no_grad()` (and grad mode is always off inside `forward`), so `apply` records the caller's grad mode and
    `_need(ctx)` combines the two: a forward that will never be differentiated takes the inference path (fused activation
    epilogues, no saved probabilities / pre-norm sums)."""

    @classmethod
    def apply(cls, *args):
        _TRACK[0] = torch.is_grad_enabled()
        return super().apply(*args)

    def __init_subclass__(cls, **kw):
        # every backward node reports the end of its gradient-buffer writes (engine.grad_writes_done: no-op unless a
        # data-parallel gradient exchange watches the parameters)
        super().__init_subclass__(**kw)
        bw = cls.__dict__.get("backward")
        if isinstance(bw, staticmethod):
            inner = bw.__func__

            def backward(ctx, *grads):
                out = inner(ctx, *grads)
                E.grad_writes_done()
                return out
            backward.__doc__ = inner.__doc__
            cls.backward = staticmethod(backward)


def _need(ctx):
    return _TRACK[0] and any(ctx.needs_input_grad)


def _wgrad(param, row0, row1, dy, x, alpha_dev=None):
    """param.grad[row0:row1, :Kin] += dy^T @ x   (dy [M, N] bf16, x [M, Kp] bf16)."""
    g = E.grad_buffer(param)
    g2 = g.view(g.shape[0], -1)
    Kin = g2.shape[1]
    M, N = dy.shape
    sk = E.wgrad_split_k(N, Kin, M)
    K.gemm(N, Kin, M, K.Op(dy, dy.stride(0), mn=True), K.Op(x, x.stride(0), mn=True),
           K.Out(g2, g2.stride(0), off=row0 * g2.stride(0)), accumulate=2 if sk > 1 else 1, split_k=sk,
           alpha_dev=alpha_dev)


def _bgrad(param, row0, row1, dy):
    g = E.grad_buffer(param)
    K.colsum_accum(dy, g[row0:row1])


def _dgrad(dy, w16, Kp, out=None, residual=None):
    """dx[M, Kp] = dy[M, N] @ w16[N, Kp] (+ residual)."""
    M, N = dy.shape
    if out is None:
        out = torch.empty((M, Kp), dtype=torch.bfloat16, device=dy.device)
    R = K.Out(residual, residual.stride(0)) if residual is not None else None
    K.gemm(M, Kp, N, K.Op(dy, dy.stride(0)), K.Op(w16, w16.stride(0), mn=True), K.Out(out, out.stride(0)), R=R)
    return out


def pad_cols(x, mult=8):
    """Zero-pad the feature dim of a 2-D bf16/fp32 tensor to a multiple of 8 and cast to bf16."""
    rows, cols = x.shape
    if x.dtype == torch.bfloat16 and cols % mult == 0 and x.stride(1) == 1 and x.stride(0) % 8 == 0:
        return x
    return K.cast_pad(x.contiguous(), torch.bfloat16, E.pad8(cols))


class _CastPadFn(_Fn):
    """fp32/bf16 [rows, cols] -> bf16 [rows, pad8(cols)] with gradient back to the source dtype."""

    @staticmethod
    def forward(ctx, x):
        ctx.cols = x.shape[1]
        ctx.dtype = x.dtype
        return K.cast_pad(x.contiguous(), torch.bfloat16, E.pad8(x.shape[1]))

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        return K.cast_pad(dy, ctx.dtype, ctx.cols, cols_src=ctx.cols)


def to_bf16_padded(x):
    if x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0 and x.is_contiguous():
        return x
    return _CastPadFn.apply(x)


class _ToF32Fn(_Fn):
    @staticmethod
    def forward(ctx, x):
        return K.to_f32(x)

    @staticmethod
    def backward(ctx, dy):
        return K.to_bf16(dy.contiguous())


def to_f32(x):
    return x if x.dtype == torch.float32 else _ToF32Fn.apply(x)


# ------------------------------------------------------------------------------------------------
class LinearFn(_Fn):
    """y = act(x @ W[row0:row1]^T + b[row0:row1] + residual) * post_gain, bf16 out."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, row0, row1, act, post_gain):
        w16 = E.w_bf16(weight)[row0:row1]
        b = bias.detach()[row0:row1] if bias is not None else None
        M = x.shape[0]
        N = row1 - row0
        need_grad = _need(ctx)
        out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
        aux = None
        if act == K.ACT_GELU and need_grad and (M * N) % 8 == 0:
            # keep the pre-activation (GELU' needs it): GEMM writes it through the fast epilogue, GELU is one extra pass
            aux = out
            K.linear(x, w16, b, act=K.ACT_NONE, residual=residual, out=aux)
            out = K.act_fwd(aux, act, post_gain)
        else:
            if act == K.ACT_GELU and need_grad:
                aux = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
            K.linear(x, w16, b, act=act, residual=residual, out=out, aux=aux, post_gain=post_gain)
        ctx.act, ctx.post_gain, ctx.row0, ctx.row1 = act, post_gain, row0, row1
        ctx.weight, ctx.bias = weight, bias
        ctx.has_res = residual is not None
        if need_grad:
            ctx.save_for_backward(x, aux if aux is not None else (out if act != K.ACT_NONE else None))
        return out

    @staticmethod
    def backward(ctx, dy):
        x, ref = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.act != K.ACT_NONE:
            dpre = K.act_bwd(dy, ref, ctx.act, ctx.post_gain)
        elif ctx.post_gain != 1.0:
            dpre = K.axpby_bcast(dy, dy, torch.bfloat16, alpha=ctx.post_gain, beta=0.0)
        else:
            dpre = dy
        dx = None
        if ctx.needs_input_grad[0]:
            w16 = E.w_bf16(ctx.weight)[ctx.row0:ctx.row1]
            dx = _dgrad(dpre, w16, x.shape[1])
        if _needs(ctx.weight):
            _wgrad(ctx.weight, ctx.row0, ctx.row1, dpre, x)
        if _needs(ctx.bias):
            _bgrad(ctx.bias, ctx.row0, ctx.row1, dpre)
        dres = dpre if (ctx.has_res and ctx.needs_input_grad[1]) else None
        return dx, dres, None, None, None, None, None, None


def linear(x, weight, bias=None, act=K.ACT_NONE, residual=None, rows=None, post_gain=1.0):
    row0, row1 = (0, weight.shape[0]) if rows is None else rows
    return LinearFn.apply(x, residual, weight, bias, row0, row1, act, post_gain)


class FusedLinearFn(_Fn):
    """y = x @ cat(W_0..W_n)^T + cat(b_0..b_n): several Linear layers sharing one input run as ONE GEMM
    (BERT query/key/value, training/med.py:109-116,157-180)."""

    @staticmethod
    def forward(ctx, x, *wb):
        ws, bs = wb[0::2], wb[1::2]
        w16 = E.w_cat_bf16(ws)
        b = E.b_cat_f32(bs)
        out = torch.empty((x.shape[0], w16.shape[0]), dtype=torch.bfloat16, device=x.device)
        K.linear(x, w16, b, out=out)
        ctx.ws, ctx.bs = ws, bs
        if _need(ctx):
            ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = _dgrad(dy, E.w_cat_bf16(ctx.ws), x.shape[1]) if ctx.needs_input_grad[0] else None
        col = 0
        for w, b in zip(ctx.ws, ctx.bs):
            n = w.shape[0]
            sl = dy[:, col:col + n]
            if _needs(w):
                _wgrad(w, 0, n, sl, x)
            if _needs(b):
                _bgrad(b, 0, n, sl)
            col += n
        return (dx,) + (None,) * (2 * len(ctx.ws))


def fused_linear(x, layers):
    args = []
    for l in layers:
        args += [l.weight, l.bias]
    return FusedLinearFn.apply(x, *args)


import os as _os
LN_BF16_DENSE_NOGRAD = _os.environ.get("LD_LN_BF16_DENSE", "1") != "0"   # see LinearLNFn.forward


class LinearLNFn(_Fn):
    """y = LayerNorm(dropout(x @ W^T + b) + residual) — the post-norm residual block tail (BERT SelfOutput /
    Output: training/med.py:237-242,321-325; DETR: training/detr_transformer.py:210-214).  The pre-norm
    sum is formed in fp32.  When the block needs no gradient (frozen text encoder, inference) and the residual is bf16,
    the dense output travels to the LayerNorm kernel as bf16 and the residual is added there (half the GEMM's store
    traffic, and its plain bf16 epilogue instead of the fp32 + residual one); with gradients the fp32 sum is kept
    for the backward pass.  With dropout (training mode) the dense output always takes the bf16 route: the mask is drawn
    where the LayerNorm kernel reads it (ld_layernorm_res_dropout_fwd) and re-applied to the gradient in backward."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, ln_w, ln_b, eps, dropout_p):
        w16 = E.w_bf16(weight)
        M, N = x.shape[0], weight.shape[0]
        need_grad = _need(ctx)
        ctx.drop = None
        res_ok = residual is not None and residual.dtype == torch.bfloat16 and residual.stride(1) == 1
        if dropout_p > 0.0:
            if not (res_ok and N % 256 == 0):
                raise RuntimeError("linear_ln: dropout needs a bf16 residual and N %% 256 == 0 (got N=%d)" % N)
            dense = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
            K.linear(x, w16, bias.detach() if bias is not None else None, out=dense)
            site = RNG.next_site()
            y, pre, mean, rstd = K.layernorm_res_dropout_fwd(dense, residual, ln_w.detach(), ln_b.detach(), eps, dropout_p, site,
                                                             save=need_grad)
            ctx.weight, ctx.bias, ctx.ln_w, ctx.ln_b = weight, bias, ln_w, ln_b
            ctx.drop = (dropout_p, site)
            if need_grad:
                ctx.save_for_backward(x, pre, mean, rstd)
            return y
        if LN_BF16_DENSE_NOGRAD and not need_grad and res_ok and M * N >= (1 << 20):
            dense = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
            K.linear(x, w16, bias.detach() if bias is not None else None, out=dense)
            y, _, _, _ = K.layernorm_fwd(dense, ln_w.detach(), ln_b.detach(), eps, residual=residual)
            return y
        pre = torch.empty((M, N), dtype=torch.float32, device=x.device)
        K.linear(x, w16, bias.detach() if bias is not None else None, residual=residual, out=pre)
        y, _, mean, rstd = K.layernorm_fwd(pre, ln_w.detach(), ln_b.detach(), eps, save_stats=need_grad)
        ctx.weight, ctx.bias, ctx.ln_w, ctx.ln_b = weight, bias, ln_w, ln_b
        if need_grad:
            ctx.save_for_backward(x, pre, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pre, mean, rstd = ctx.saved_tensors
        dgamma = E.grad_buffer(ctx.ln_w) if _needs(ctx.ln_w) else None
        dbeta = E.grad_buffer(ctx.ln_b) if _needs(ctx.ln_b) else None
        dpre = K.layernorm_bwd(dy.contiguous(), pre, mean, rstd, ctx.ln_w.detach(), dgamma, dbeta)
        ddense = K.dropout(dpre, ctx.drop[0], ctx.drop[1]) if ctx.drop is not None else dpre     # same Philox mask as the forward
        dx = _dgrad(ddense, E.w_bf16(ctx.weight), x.shape[1]) if ctx.needs_input_grad[0] else None
        if _needs(ctx.weight):
            _wgrad(ctx.weight, 0, ctx.weight.shape[0], ddense, x)
        if _needs(ctx.bias):
            _bgrad(ctx.bias, 0, ctx.bias.shape[0], ddense)
        dres = dpre if ctx.needs_input_grad[1] else None
        return dx, dres, None, None, None, None, None, None


def linear_ln(x, residual, weight, bias, ln_w, ln_b, eps, dropout_p=0.0):
    return LinearLNFn.apply(x, residual, weight, bias, ln_w, ln_b, eps, float(dropout_p))


class DropoutFn(_Fn):
    """y = dropout(x) on a contiguous bf16 / fp32 tensor (embedding dropout training/med.py:96, the FFN-inner dropout of
    training/detr_transformer.py:212 and of nn.TransformerEncoderLayer); the backward pass re-applies the same mask."""

    @staticmethod
    def forward(ctx, x, dropout_p):
        site = RNG.next_site()
        ctx.drop = (dropout_p, site)
        return K.dropout(x.contiguous(), dropout_p, site)

    @staticmethod
    def backward(ctx, dy):
        return K.dropout(dy.contiguous(), ctx.drop[0], ctx.drop[1]), None


def dropout(x, dropout_p):
    return DropoutFn.apply(x, float(dropout_p)) if dropout_p > 0.0 else x


class LayerNormFn(_Fn):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, eps):
        need_grad = _need(ctx)
        y, _, mean, rstd = K.layernorm_fwd(x, ln_w.detach(), ln_b.detach(), eps, save_stats=need_grad)
        ctx.ln_w, ctx.ln_b = ln_w, ln_b
        if need_grad:
            ctx.save_for_backward(x, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        dgamma = E.grad_buffer(ctx.ln_w) if _needs(ctx.ln_w) else None
        dbeta = E.grad_buffer(ctx.ln_b) if _needs(ctx.ln_b) else None
        dx = K.layernorm_bwd(dy.contiguous(), x, mean, rstd, ctx.ln_w.detach(), dgamma, dbeta, out_dtype=x.dtype)
        return dx, None, None, None


def layernorm(x, ln_w, ln_b, eps):
    return LayerNormFn.apply(x, ln_w, ln_b, eps)


class AddBcastFn(_Fn):
    """out = a + b with b broadcast over the leading dim of a (src + pos)."""

    @staticmethod
    def forward(ctx, a, b):
        return K.axpby_bcast(a, b, torch.bfloat16)

    @staticmethod
    def backward(ctx, dy):
        return dy, None


def add_bcast(a, b):
    return AddBcastFn.apply(a, b)


# ------------------------------------------------------------------------------------------------
class EmbedLNFn(_Fn):
    """BERT embeddings: LN(word[ids] + pos[t]) -> bf16 (training/med.py:74-97)."""

    @staticmethod
    def forward(ctx, ids, word, pos, ln_w, ln_b, T, eps, pad_id):
        need_grad = _need(ctx)
        y, pre, mean, rstd = K.embed_ln_fwd(ids, word.detach(), pos.detach(), ln_w.detach(), ln_b.detach(), T, eps, save=need_grad)
        ctx.params = (word, pos, ln_w, ln_b)
        ctx.T, ctx.pad_id = T, pad_id
        if need_grad:
            ctx.save_for_backward(ids, pre, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        ids, pre, mean, rstd = ctx.saved_tensors
        word, pos, ln_w, ln_b = ctx.params
        dgamma = E.grad_buffer(ln_w) if _needs(ln_w) else None
        dbeta = E.grad_buffer(ln_b) if _needs(ln_b) else None
        dpre = K.layernorm_bwd(dy.contiguous(), pre, mean, rstd, ln_w.detach(), dgamma, dbeta, out_dtype=torch.float32)
        K.embed_bwd(ids, dpre, E.grad_buffer(word) if _needs(word) else None,
                    E.grad_buffer(pos) if _needs(pos) else None, ctx.T, ctx.pad_id)
        return None, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
FUSED_ATTENTION = True       # False -> batched-GEMM attention (QK^T with fused softmax epilogue, then PV); used by tests


class AttentionFn(_Fn):
    """Multi-head attention core on projected buffers.

      S = (Q K^T) * scale (+ mask);  P = softmax(S);  O = dropout(P) V        per (batch b, head h)

    q_t / k_t / v_t are 2-D bf16 buffers `[B*L, ld]`; head h of q lives in columns
    [q_off + h*d, q_off + (h+1)*d).  They may alias (fused QKV / QK buffers); gradients are returned
    once per distinct buffer.  Reference: BertSelfAttention.forward training/med.py:146-228 and
    F.multi_head_attention_forward as used by training/detr_transformer.py:208,273,277.

    <= 256 keys, head_dim <= 192: one fused kernel forward (ld_attention_fwd; only a per-row log-sum-exp is saved) and one
    fused kernel backward (ld_attention_bwd: P recomputed, dropout mask regenerated, dQ on chip) + the two transposed GEMMs.
    Larger shapes: batched tcgen05 GEMMs with the softmax in the first GEMM's epilogue / a masked-softmax kernel.
    """

    @staticmethod
    def forward(ctx, q_t, k_t, v_t, q_off, k_off, v_off, B, H, Lq, Lk, d, scale, key_mask, mask_inf, causal, dropout_p):
        dev = q_t.device
        Lkp = E.pad8(Lk)
        ldq, ldk, ldv = q_t.stride(0), k_t.stride(0), v_t.stride(0)
        need_grad = _need(ctx)
        ctx.dims = (q_off, k_off, v_off, B, H, Lq, Lk, d, scale)
        ctx.mask = (key_mask, mask_inf, causal)
        site = RNG.next_site() if dropout_p > 0.0 else 0
        ctx.drop = (dropout_p, site)
        ctx.fused = bool(Lk <= 256 and d <= 192 and d % 8 == 0 and FUSED_ATTENTION)
        if ctx.fused:
            lse = torch.empty((B * H, Lq), dtype=torch.float32, device=dev) if need_grad else None
            O = K.attention_fwd(q_t, q_off, k_t, k_off, v_t, v_off, B, H, Lq, Lk, d, scale, key_mask=key_mask,
                                mask_inf=mask_inf, causal=causal, lse_out=lse, dropout_p=dropout_p, rng_site=site)
            if need_grad:
                ctx.save_for_backward(q_t, k_t, v_t, O, lse)
            return O
        P = (torch.zeros if Lkp != Lk else torch.empty)((B * H, Lq, Lkp), dtype=torch.bfloat16, device=dev)
        if Lk <= 256:
            # scores never leave the SM: scale + mask + softmax + bf16 cast run in the GEMM's TMEM drain
            K.gemm(Lq, Lk, d, K.Op(q_t, ldq, off=q_off, sb1=Lq * ldq, sb2=d), K.Op(k_t, ldk, off=k_off, sb1=Lk * ldk, sb2=d),
                   K.Out(P, Lkp, sb1=H * Lq * Lkp, sb2=Lq * Lkp), nb1=B, nb2=H, alpha=scale,
                   softmax=dict(key_mask=key_mask, mask_inf=mask_inf, causal=causal))
        else:
            S = torch.empty((B * H, Lq, Lk), dtype=torch.float32, device=dev)
            K.gemm(Lq, Lk, d, K.Op(q_t, ldq, off=q_off, sb1=Lq * ldq, sb2=d), K.Op(k_t, ldk, off=k_off, sb1=Lk * ldk, sb2=d),
                   K.Out(S, Lk, sb1=H * Lq * Lk, sb2=Lq * Lk), nb1=B, nb2=H)
            K.softmax_fwd(S, P, B, H, Lq, Lk, scale, key_mask=key_mask, mask_inf=mask_inf, causal=causal)
            del S
        Pd = K.dropout(P, dropout_p, site) if dropout_p > 0.0 else P
        O = torch.empty((B * Lq, H * d), dtype=torch.bfloat16, device=dev)
        K.gemm(Lq, d, Lk, K.Op(Pd, Lkp, sb1=H * Lq * Lkp, sb2=Lq * Lkp), K.Op(v_t, ldv, off=v_off, sb1=Lk * ldv, sb2=d, mn=True),
               K.Out(O, H * d, sb1=Lq * H * d, sb2=d), nb1=B, nb2=H)
        if need_grad:
            ctx.save_for_backward(q_t, k_t, v_t, P, Pd)
        return O

    @staticmethod
    def backward(ctx, dO):
        q_off, k_off, v_off, B, H, Lq, Lk, d, scale = ctx.dims
        key_mask, mask_inf, causal = ctx.mask
        dropout_p, site = ctx.drop
        dev = dO.device
        dO = dO.contiguous()
        Lkp = E.pad8(Lk)
        q_t, k_t, v_t = ctx.saved_tensors[:3]
        ldq, ldk, ldv, ldo = q_t.stride(0), k_t.stride(0), v_t.stride(0), H * d
        # gradient buffers, one per distinct input buffer
        bufs = {}
        ptrs = [q_t.data_ptr(), k_t.data_ptr(), v_t.data_ptr()]

        def gbuf(t):
            key = t.data_ptr()
            if key not in bufs:
                cover = ptrs.count(key) * H * d
                mk = torch.empty_like if cover == t.shape[1] else torch.zeros_like
                bufs[key] = mk(t)
            return bufs[key]

        dq_t, dk_t, dv_t = gbuf(q_t), gbuf(k_t), gbuf(v_t)
        if ctx.fused:
            O, lse = ctx.saved_tensors[3:]
            Pd = torch.empty((B * H, Lq, Lkp), dtype=torch.bfloat16, device=dev)
            dS = torch.empty((B * H, Lq, Lkp), dtype=torch.bfloat16, device=dev)
            # P recomputed from lse, dropout mask regenerated, dQ = dS K — all in one kernel
            K.attention_bwd(q_t, q_off, k_t, k_off, v_t, v_off, O, dO, lse, dq_t, Pd, dS, B, H, Lq, Lk, d, scale,
                            key_mask=key_mask, mask_inf=mask_inf, causal=causal, dropout_p=dropout_p, rng_site=site)
        else:
            P, Pd = ctx.saved_tensors[3:]
            # dP = dO V^T  (gradient w.r.t. the dropped probabilities)
            dP = (torch.zeros if Lkp != Lk else torch.empty)((B * H, Lq, Lkp), dtype=torch.float32, device=dev)
            K.gemm(Lq, Lk, d, K.Op(dO, ldo, sb1=Lq * ldo, sb2=d), K.Op(v_t, ldv, off=v_off, sb1=Lk * ldv, sb2=d),
                   K.Out(dP, Lkp, sb1=H * Lq * Lkp, sb2=Lq * Lkp), nb1=B, nb2=H)
            if dropout_p > 0.0:
                K.dropout(dP, dropout_p, site, out=dP)
            dS = torch.empty_like(P) if Lkp == Lk else torch.zeros_like(P)
            K.softmax_bwd(P, dP, dS, B * H, Lq, Lk, scale)
            del dP
            # dQ = dS K
            K.gemm(Lq, d, Lk, K.Op(dS, Lkp, sb1=H * Lq * Lkp, sb2=Lq * Lkp), K.Op(k_t, ldk, off=k_off, sb1=Lk * ldk, sb2=d, mn=True),
                   K.Out(dq_t, ldq, off=q_off, sb1=Lq * ldq, sb2=d), nb1=B, nb2=H)
        # dK = dS^T Q
        K.gemm(Lk, d, Lq, K.Op(dS, Lkp, sb1=H * Lq * Lkp, sb2=Lq * Lkp, mn=True), K.Op(q_t, ldq, off=q_off, sb1=Lq * ldq, sb2=d, mn=True),
               K.Out(dk_t, ldk, off=k_off, sb1=Lk * ldk, sb2=d), nb1=B, nb2=H)
        # dV = dropout(P)^T dO
        K.gemm(Lk, d, Lq, K.Op(Pd, Lkp, sb1=H * Lq * Lkp, sb2=Lq * Lkp, mn=True), K.Op(dO, ldo, sb1=Lq * ldo, sb2=d, mn=True),
               K.Out(dv_t, ldv, off=v_off, sb1=Lk * ldv, sb2=d), nb1=B, nb2=H)
        gq = dq_t
        gk = dk_t if k_t.data_ptr() != q_t.data_ptr() else None
        gv = dv_t if v_t.data_ptr() not in (q_t.data_ptr(), k_t.data_ptr()) else None
        return (gq, gk, gv) + (None,) * 13


def attention(q_t, k_t, v_t, q_off, k_off, v_off, B, H, Lq, Lk, d, scale=None, key_mask=None, mask_inf=False, causal=False,
              dropout_p=0.0):
    scale = (1.0 / math.sqrt(d)) if scale is None else scale
    return AttentionFn.apply(q_t, k_t, v_t, q_off, k_off, v_off, B, H, Lq, Lk, d, scale, key_mask, mask_inf, causal, float(dropout_p))


# ------------------------------------------------------------------------------------------------
class LMHeadCEFn(_Fn):
    """loss = mean over non-ignored rows of CE(h @ W_emb^T + b, labels; label_smoothing) — the tied
    30524-way LM head + loss (training/med.py:504-538, 910-920).  Logits live only as one bf16
    buffer that is overwritten in place by d(loss)/d(logits)."""

    @staticmethod
    def forward(ctx, h, emb_weight, bias, labels, inv_n_valid, smoothing):
        M, C = h.shape
        V = emb_weight.shape[0]
        Vp = E.pad8(V)
        w16 = E.w_bf16(emb_weight)
        buf = torch.empty((M, Vp), dtype=torch.bfloat16, device=h.device)
        logits = buf[:, :V]
        K.gemm(M, V, C, K.Op(h, h.stride(0)), K.Op(w16, w16.stride(0)), K.Out(buf, Vp), col_bias=bias.detach())
        need_grad = _need(ctx)
        # inv_n_valid: fp32 device scalar = 1 / #targets (device-side so a captured CUDA graph stays valid for any batch)
        loss_rows = K.cross_entropy(logits, labels, smoothing, -100, True, logits if need_grad else None, grad_scale=1.0,
                                    grad_scale_dev=inv_n_valid)
        ctx.emb_weight, ctx.bias = emb_weight, bias
        if need_grad:
            ctx.save_for_backward(h, buf)
        ctx.V = V
        return loss_rows.sum() * inv_n_valid.reshape(())

    @staticmethod
    def backward(ctx, g):
        h, buf = ctx.saved_tensors
        V = ctx.V
        M, C = h.shape
        g = g.contiguous().float()
        dl = buf[:, :V]
        dh = None
        if ctx.needs_input_grad[0]:
            w16 = E.w_bf16(ctx.emb_weight)
            dh = torch.empty((M, C), dtype=torch.bfloat16, device=h.device)
            K.gemm(M, C, V, K.Op(buf, buf.stride(0)), K.Op(w16, w16.stride(0), mn=True), K.Out(dh, C), alpha_dev=g)
        if _needs(ctx.emb_weight):
            _wgrad(ctx.emb_weight, 0, V, dl, h, alpha_dev=g)
        if _needs(ctx.bias):
            # sum over the padded width (128-bit path); the pad columns of `buf` are never written and are dropped here
            tmp = torch.zeros(buf.shape[1], dtype=torch.float32, device=h.device)
            K.colsum_accum(buf, tmp)
            E.grad_buffer(ctx.bias).add_(tmp[:V] * g)
        return dh, None, None, None, None, None


def lm_head_ce(h, emb_weight, bias, labels, inv_n_valid, smoothing=0.1):
    return LMHeadCEFn.apply(h, emb_weight, bias, labels, inv_n_valid, smoothing)


class CrossEntropyFn(_Fn):
    """Mean cross-entropy over rows (small heads: class logits, text-length logits)."""

    @staticmethod
    def forward(ctx, logits, labels):
        logits = logits.contiguous()
        rows = logits.shape[0]
        dl = torch.empty_like(logits) if ctx.needs_input_grad[0] else None
        loss_rows = K.cross_entropy(logits, labels, 0.0, -100, True, dl, grad_scale=1.0 / rows)
        if dl is not None:
            ctx.save_for_backward(dl)
        return loss_rows.sum() / rows

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None


def cross_entropy(logits, labels):
    return CrossEntropyFn.apply(logits, labels)


# ------------------------------------------------------------------------------------------------
# convolution = (im2col) + tcgen05 GEMM with fused per-channel scale/bias (+residual) (+ReLU)
def _make_ohwi(weight):
    Cout = weight.shape[0]
    w = weight.detach().permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    return K.cast_pad(w, torch.bfloat16, E.pad8(w.shape[1]))


def conv_weight_bf16(weight):
    """OIHW fp32 parameter -> bf16 [Cout, pad8(KH*KW*Cin)] with K ordered (kh, kw, ci) to match im2col."""
    return E.derived((weight,), "ohwi", _make_ohwi)


def _make_flipT(weight):
    """OIHW -> bf16 [Cin, KH*KW*Cout] with the taps reversed: the weights of the convolution that maps dy to dx (stride 1)."""
    Cin = weight.shape[1]
    w = weight.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(Cin, -1).contiguous()
    return K.cast_pad(w, torch.bfloat16, E.pad8(w.shape[1]))


def conv_weight_flipT_bf16(weight):
    return E.derived((weight,), "ihwo_flip", _make_flipT)


IMPLICIT_CONV = os.environ.get("LD_CONV_IMPLICIT", "1") != "0"     # 0: every convolution goes through an explicit patch matrix


def _conv_geom(img, mode, B, H, W, C, Ho, Wo, KH, KW, stride, pad):
    return dict(img=img, mode=mode, B=B, H=H, W=W, C=C, Ho=Ho, Wo=Wo, KH=KH, KW=KW, stride=stride, pad=pad)


class Conv2dFn(_Fn):
    """y = act(conv(x, W) * scale + shift + residual) on channels-last bf16 activations.

    Implicit GEMM (ld_conv_gemm_bf16: the TMA producer reads boxes of pixels x 64 channels straight from the NHWC image, no
    patch matrix in HBM) whenever the channel count is a multiple of 64 (or exactly 32: 64-byte-swizzle K blocks, forward and data
    gradient) and a box of output pixels is a rectangle of whole rows —
    forward, the stride-1 data gradient (a convolution of dy with the flipped, transposed weights) and the weight gradient;
    the 3-channel stem and odd geometries keep the explicit im2col / col2im path.

    x: [B*H*W, Cin] bf16 (NHWC rows).  scale/shift: fp32 [Cout] (folded FrozenBatchNorm2d, reference
    training/detr_backbone.py:55-65, or a conv bias).  Reference convs: torchvision resnet50 via
    training/detr_backbone.py:105, nn.Conv2d input_proj training/networks_detr.py:82."""

    @staticmethod
    def forward(ctx, x, weight, scale, shift, residual, geom, act):
        B, H, W, stride, pad = geom
        Cout, Cin, KH, KW = weight.shape
        w16 = conv_weight_bf16(weight)
        direct = (KH == 1 and KW == 1 and stride == 1 and pad == 0 and Cin % 8 == 0)
        Ho, Wo = K.conv_out_size(H, KH, stride, pad), K.conv_out_size(W, KW, stride, pad)
        M = B * Ho * Wo
        out = torch.empty((M, Cout), dtype=torch.bfloat16, device=x.device)
        epi = dict(act=act, R=K.Out(residual, residual.stride(0)) if residual is not None else None,
                   col_scale=scale, col_bias=shift.detach() if shift is not None else None)
        implicit = (not direct and IMPLICIT_CONV and (Cin % 64 == 0 or Cin == 32) and x.is_contiguous() and K.conv_box_ok(Ho, Wo, stride, 128))
        if implicit:
            K.gemm(M, Cout, KH * KW * Cin, K.Op(x, Cin), K.Op(w16, w16.stride(0)), K.Out(out, Cout),
                   conv=_conv_geom(x, 1, B, H, W, Cin, Ho, Wo, KH, KW, stride, pad), **epi)
        else:
            cols = x if direct else K.im2col(x, B, H, W, Cin, KH, KW, stride, pad)[0]
            K.gemm(M, Cout, cols.shape[1], K.Op(cols, cols.stride(0)), K.Op(w16, w16.stride(0)), K.Out(out, Cout), **epi)
        # the 3-channel stem keeps its (small-K) patch matrix for the weight gradient instead of gathering it twice
        ctx.cols = cols if (not implicit and not direct and _need(ctx) and _needs(weight) and Cin < 8) else None
        ctx.geom = (B, H, W, stride, pad, Ho, Wo, direct)
        ctx.weight, ctx.act, ctx.scale = weight, act, scale
        ctx.shift_param = shift if isinstance(shift, torch.nn.Parameter) else None
        ctx.has_res = residual is not None
        if _need(ctx):
            ctx.save_for_backward(x, out if act != K.ACT_NONE else None)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, out = ctx.saved_tensors
        B, H, W, stride, pad, Ho, Wo, direct = ctx.geom
        weight = ctx.weight
        Cout, Cin, KH, KW = weight.shape
        dy = dy.contiguous()
        want_res = ctx.has_res and ctx.needs_input_grad[4]
        if ctx.scale is not None and Cout % 8 == 0 and ctx.act in (K.ACT_NONE, K.ACT_RELU, K.ACT_LRELU):
            # one pass: activation backward (for the residual branch) + FrozenBN scale (for the convolution)
            dpre, dconv = K.act_bwd_colscale(dy, out, ctx.scale, ctx.act, want_dx=want_res and ctx.act != K.ACT_NONE)
            if ctx.act == K.ACT_NONE:
                dpre = dy
        else:
            dpre = K.act_bwd(dy, out, ctx.act) if ctx.act != K.ACT_NONE else dy
            dconv = dpre
            if ctx.scale is not None:
                dconv = K.scale_channels(dpre, ctx.scale, torch.bfloat16, dpre.numel(), Cout)
        dres = dpre if want_res else None
        if ctx.shift_param is not None and ctx.needs_input_grad[3]:
            _bgrad(ctx.shift_param, 0, Cout, dpre if dpre is not None else dy)
        w16 = conv_weight_bf16(weight)
        dx = None
        dconv = dconv.contiguous()
        if ctx.needs_input_grad[0]:
            same = stride == 1 and Ho == H and Wo == W and KH == KW
            if (not direct and IMPLICIT_CONV and same and (Cout % 64 == 0 or Cout == 32) and K.conv_box_ok(H, W, 1, 128)):
                # dx = conv(dy, flipped weights), pad' = k - 1 - pad: no [M, 9 Cin] temporary, no col2im pass
                wf = conv_weight_flipT_bf16(weight)
                dx = torch.empty((B * H * W, Cin), dtype=torch.bfloat16, device=x.device)
                K.gemm(B * H * W, Cin, KH * KW * Cout, K.Op(dconv, Cout), K.Op(wf, wf.stride(0)), K.Out(dx, Cin),
                       conv=_conv_geom(dconv, 1, B, Ho, Wo, Cout, H, W, KH, KW, 1, KH - 1 - pad))
            else:
                dcols = _dgrad(dconv, w16, w16.shape[1])
                dx = dcols if direct else K.col2im(dcols, B, H, W, Cin, Ho, Wo, KH, KW, stride, pad)
        if _needs(weight):
            g = E.grad_buffer(weight)
            M = dconv.shape[0]
            Kreal = KH * KW * Cin
            sk = E.wgrad_split_k(Cout, Kreal, M)
            implicit = (not direct and IMPLICIT_CONV and Cin % 64 == 0 and x.is_contiguous() and K.conv_box_ok(Ho, Wo, stride, 64))
            cv = _conv_geom(x, 2, B, H, W, Cin, Ho, Wo, KH, KW, stride, pad) if implicit else None
            cols = x if (direct or implicit) else (ctx.cols if ctx.cols is not None else K.im2col(x, B, H, W, Cin, KH, KW, stride, pad)[0])
            if KH == 1 and KW == 1:
                g2 = g.view(Cout, Cin)
                K.gemm(Cout, Cin, M, K.Op(dconv, Cout, mn=True), K.Op(cols, cols.stride(0), mn=True), K.Out(g2, Cin),
                       accumulate=2 if sk > 1 else 1, split_k=sk, conv=cv)
            else:
                tmp = torch.zeros((Cout, Kreal), dtype=torch.float32, device=x.device) if sk > 1 else \
                    torch.empty((Cout, Kreal), dtype=torch.float32, device=x.device)
                K.gemm(Cout, Kreal, M, K.Op(dconv, Cout, mn=True), K.Op(cols, cols.stride(0), mn=True), K.Out(tmp, Kreal),
                       accumulate=2 if sk > 1 else 0, split_k=sk, conv=cv)
                g.add_(tmp.view(Cout, KH, KW, Cin).permute(0, 3, 1, 2))
        return dx, None, None, None, dres, None, None


def conv2d(x, weight, scale, shift, residual, B, H, W, stride, pad, act):
    return Conv2dFn.apply(x, weight, scale, shift, residual, (B, H, W, stride, pad), act)


class MaxPool3s2Fn(_Fn):
    @staticmethod
    def forward(ctx, x, B, H, W):
        C = x.shape[1]
        y, arg, Ho, Wo = K.maxpool3s2_fwd(x, B, H, W, C, save_argmax=ctx.needs_input_grad[0])
        ctx.geom = (B, H, W, C)
        if arg is not None:
            ctx.save_for_backward(arg)
        return y

    @staticmethod
    def backward(ctx, dy):
        (arg,) = ctx.saved_tensors
        B, H, W, C = ctx.geom
        return K.maxpool3s2_bwd(dy.contiguous(), arg, B, H, W, C), None, None, None


def maxpool3s2(x, B, H, W):
    return MaxPool3s2Fn.apply(x, B, H, W)


# ------------------------------------------------------------------------------------------------
class LinearF32Fn(_Fn):
    """Head projection with fp32 output (boxes, logits): y = x @ W^T + b."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        w16 = E.w_bf16(weight)
        out = torch.empty((x.shape[0], weight.shape[0]), dtype=torch.float32, device=x.device)
        K.linear(x, w16, bias.detach() if bias is not None else None, out=out)
        ctx.weight, ctx.bias = weight, bias
        if _need(ctx):
            ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy16 = K.cast_pad(dy.contiguous(), torch.bfloat16, E.pad8(dy.shape[1]))      # pad N for the TMA stride rule
        N = dy.shape[1]
        dy_v = dy16[:, :N]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _dgrad(dy_v, E.w_bf16(ctx.weight), x.shape[1])
        if _needs(ctx.weight):
            _wgrad(ctx.weight, 0, N, dy_v, x)
        if _needs(ctx.bias):
            _bgrad(ctx.bias, 0, N, dy_v)
        return dx, None, None


def linear_f32(x, weight, bias=None):
    return LinearF32Fn.apply(x, weight, bias)


class ScaledLinearFn(_Fn):
    """StyleGAN2 FullyConnectedLayer: y = act(x @ (W * wg)^T + b * bg) * gain, bf16 or fp32 out
    (reference training/networks_stylegan2.py:92-126; lrelu gain sqrt(2) from bias_act.py:26)."""

    @staticmethod
    def forward(ctx, x, weight, bias, wg, bg, act, gain, out_f32):
        w16 = E.w_bf16(weight)
        b = None
        if bias is not None:
            b = E.derived((bias,), "bg%g" % bg, lambda t, g=float(bg): (t.detach().float() * g).contiguous())
        out = torch.empty((x.shape[0], weight.shape[0]), dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
        K.linear(x, w16, b, act=act, out=out, alpha=wg, post_gain=gain)
        ctx.cfg = (wg, bg, act, gain, out_f32)
        ctx.weight, ctx.bias = weight, bias
        if _need(ctx):
            ctx.save_for_backward(x, out if act != K.ACT_NONE else None)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, out = ctx.saved_tensors
        wg, bg, act, gain, out_f32 = ctx.cfg
        dy = dy.contiguous()
        if out_f32:
            assert act == K.ACT_NONE
            dy = K.to_bf16(dy)
        if act != K.ACT_NONE:
            dpre = K.act_bwd(dy, out, act, gain)
        elif gain != 1.0:
            dpre = K.axpby_bcast(dy, dy, torch.bfloat16, alpha=gain, beta=0.0)
        else:
            dpre = dy
        M, N = dpre.shape
        dx = None
        if ctx.needs_input_grad[0]:
            w16 = E.w_bf16(ctx.weight)
            dx = torch.empty((M, x.shape[1]), dtype=torch.bfloat16, device=x.device)
            K.gemm(M, x.shape[1], N, K.Op(dpre, dpre.stride(0)), K.Op(w16, w16.stride(0), mn=True), K.Out(dx, dx.stride(0)), alpha=wg)
        if _needs(ctx.weight):
            g = E.grad_buffer(ctx.weight)
            K.gemm(N, g.shape[1], M, K.Op(dpre, dpre.stride(0), mn=True), K.Op(x, x.stride(0), mn=True), K.Out(g, g.stride(0)),
                   accumulate=1, alpha=wg)
        if _needs(ctx.bias):
            if bg == 1.0:
                _bgrad(ctx.bias, 0, N, dpre)
            else:
                tmp = torch.zeros(N, dtype=torch.float32, device=x.device)
                K.colsum_accum(dpre, tmp)
                E.grad_buffer(ctx.bias).add_(tmp, alpha=bg)
        return dx, None, None, None, None, None, None, None


def scaled_linear(x, weight, bias, wg, bg, act=K.ACT_NONE, gain=1.0, out_f32=False):
    return ScaledLinearFn.apply(x, weight, bias, wg, bg, act, gain, out_f32)


class ScaleChannelsFn(_Fn):
    """y[b,p,c] = x[b,p,c] * s[b,c]  (style modulation of the activations, networks_stylegan2.py:67)."""

    @staticmethod
    def forward(ctx, x, s, B, pixels, C):
        s = s.contiguous()
        y = K.scale_channels(x, s, torch.bfloat16, pixels * C, C)
        ctx.geom = (B, pixels, C)
        if _need(ctx):
            ctx.save_for_backward(x, s)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, s = ctx.saved_tensors
        B, pixels, C = ctx.geom
        dy = dy.contiguous()
        dx = K.scale_channels(dy, s, torch.bfloat16, pixels * C, C) if ctx.needs_input_grad[0] else None
        ds = K.channel_dot(x, dy, B, pixels, C) if ctx.needs_input_grad[1] else None
        return dx, ds, None, None, None


def scale_channels(x, s, B, pixels, C):
    return ScaleChannelsFn.apply(x, s, B, pixels, C)


class DemodBiasActFn(_Fn):
    """y = act(x * d[b,c] + bias[c]) * gain on channels-last activations (demodulation + bias_act)."""

    @staticmethod
    def forward(ctx, x, d, bias, B, pixels, C, act, gain):
        dd = d.contiguous() if d is not None else None
        y = K.demod_bias_act_fwd(x, dd, bias.detach() if bias is not None else None, B, pixels, C, act, gain)
        ctx.geom = (B, pixels, C, act, gain)
        ctx.bias = bias
        if _need(ctx):
            ctx.save_for_backward(x, dd, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, d, y = ctx.saved_tensors
        B, pixels, C, act, gain = ctx.geom
        dd = torch.zeros((B, C), dtype=torch.float32, device=x.device) if (d is not None and ctx.needs_input_grad[1]) else None
        db = E.grad_buffer(ctx.bias) if _needs(ctx.bias) else None
        dx = K.demod_bias_act_bwd(dy.contiguous(), y, x, d, dd, db, B, pixels, C, act, gain)
        return dx, dd, None, None, None, None, None, None


def demod_bias_act(x, d, bias, B, pixels, C, act, gain):
    return DemodBiasActFn.apply(x, d, bias, B, pixels, C, act, gain)


class ConvTransposeUp2Fn(_Fn):
    """Stride-2 transposed 3x3 convolution on channels-last bf16 (the `up=2` branch of conv2d_resample,
    torch_utils/ops/conv2d_resample.py:113-130): out[b, 2h+kh, 2w+kw, co] += x[b,h,w,ci] * W[co,ci,kh,kw].
    GEMM  cols = x @ Wt^T  ([B*H*W, 9*Cout], tcgen05)  followed by the gather-form col2im."""

    @staticmethod
    def _make_khkwoi(weight):
        Cout, Cin, KH, KW = weight.shape
        w = weight.detach().permute(2, 3, 0, 1).reshape(KH * KW * Cout, Cin).contiguous()
        return K.cast_pad(w, torch.bfloat16, E.pad8(Cin))

    @staticmethod
    def wt(weight):
        return E.derived((weight,), "khkwoi", ConvTransposeUp2Fn._make_khkwoi)

    @staticmethod
    def forward(ctx, x, weight, B, H, W):
        Cout, Cin, KH, KW = weight.shape
        wt = ConvTransposeUp2Fn.wt(weight)
        M = B * H * W
        cols = torch.empty((M, KH * KW * Cout), dtype=torch.bfloat16, device=x.device)
        K.gemm(M, KH * KW * Cout, x.shape[1], K.Op(x, x.stride(0)), K.Op(wt, wt.stride(0)), K.Out(cols, cols.stride(0)))
        Ho, Wo = 2 * H + 1, 2 * W + 1
        out = K.col2im(cols, B, Ho, Wo, Cout, H, W, KH, KW, 2, 0)
        ctx.geom = (B, H, W)
        ctx.weight = weight
        if _need(ctx):
            ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        B, H, W = ctx.geom
        weight = ctx.weight
        Cout, Cin, KH, KW = weight.shape
        Ho, Wo = 2 * H + 1, 2 * W + 1
        dcols, _, _ = K.im2col(dy.contiguous(), B, Ho, Wo, Cout, KH, KW, 2, 0)        # [B*H*W, 9*Cout]
        wt = ConvTransposeUp2Fn.wt(weight)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _dgrad(dcols, wt, wt.shape[1])
        if _needs(weight):
            M = dcols.shape[0]
            sk = E.wgrad_split_k(KH * KW * Cout, Cin, M)       # few output tiles, reduction over every pixel: split K
            tmp = (torch.zeros if sk > 1 else torch.empty)((KH * KW * Cout, Cin), dtype=torch.float32, device=x.device)
            K.gemm(KH * KW * Cout, Cin, M, K.Op(dcols, dcols.stride(0), mn=True), K.Op(x, x.stride(0), mn=True), K.Out(tmp, Cin),
                   accumulate=2 if sk > 1 else 0, split_k=sk)
            E.grad_buffer(weight).add_(tmp.view(KH, KW, Cout, Cin).permute(2, 3, 0, 1))
        return dx, None, None, None, None


def conv_transpose_up2(x, weight, B, H, W):
    return ConvTransposeUp2Fn.apply(x, weight, B, H, W)


class UpfirdnNHWCFn(_Fn):
    """upfirdn2d on channels-last bf16 `[B*H*W, C]` activations; backward is another upfirdn2d with the
    flipped filter and up/down swapped (torch_utils/ops/upfirdn2d.py:248-270)."""

    @staticmethod
    def forward(ctx, x, f, B, H, W, up, down, pad, flip, gain):
        C = x.shape[1]
        x4 = x.view(B, H, W, C).permute(0, 3, 1, 2)
        y4 = K.upfirdn2d_raw(x4, f, up, up, down, down, pad[0], pad[1], pad[2], pad[3], flip, gain, channels_last_out=True)
        oh, ow = y4.shape[2], y4.shape[3]
        ctx.cfg = (f, B, H, W, C, oh, ow, up, down, pad, flip, gain)
        return y4.permute(0, 2, 3, 1).reshape(B * oh * ow, C)

    @staticmethod
    def backward(ctx, dy):
        f, B, H, W, C, oh, ow, up, down, pad, flip, gain = ctx.cfg
        fh, fw = f.shape
        p = [fw - pad[0] - 1, W * up - ow * down + pad[0] - up + 1, fh - pad[2] - 1, H * up - oh * down + pad[2] - up + 1]
        d4 = dy.contiguous().view(B, oh, ow, C).permute(0, 3, 1, 2)
        x4 = K.upfirdn2d_raw(d4, f, down, down, up, up, p[0], p[1], p[2], p[3], not flip, gain, channels_last_out=True)
        return (x4.permute(0, 2, 3, 1).reshape(B * H * W, C),) + (None,) * 9


def upfirdn_nhwc(x, f, B, H, W, up=1, down=1, pad=(0, 0, 0, 0), flip=False, gain=1.0):
    return UpfirdnNHWCFn.apply(x, f, B, H, W, up, down, tuple(pad), flip, gain)
