"""DETR encoder-decoder (post-norm) on the sm_100a kernels.

Mirror of the reference's training/detr_transformer.py (Transformer :73, TransformerWithToken :22,
TransformerEncoderLayer.forward_post :202, TransformerDecoderLayer.forward_post :265) and of the
torch `nn.TransformerEncoderLayer` stacks the discriminator uses (training/networks_detr.py:242-243,
274-275; training/util.py:13-43).  `nn.MultiheadAttention` / `nn.Linear` / `nn.LayerNorm` objects are
kept as PARAMETER HOLDERS so state_dict keys match (`self_attn.in_proj_weight`, `linear1.weight`, ...);
their forward is never called — all math runs through layoutdetr_b200.functional.

Token layout here is batch-major `[B*L, d_model]` (row = b*L + l) instead of the reference's
`[L, B, d_model]`; results are layout-independent.  Dropout (p = 0.1: attention probabilities inside
nn.MultiheadAttention, dropout1/2/3 on the sub-layer outputs, `dropout` after the FFN's ReLU; reference
training/detr_transformer.py:185-194,210-214,270-285) is live in `.train()` mode, drawn in-kernel (layoutdetr_b200.rng).
"""
import copy

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import kernels as K
from .. import rng as RNG

LN_EPS = 1e-5


def _get_clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


def _key_mask(mask):
    """bool [B, L] (True = ignore) -> uint8 tensor for the softmax kernel, or None."""
    if mask is None:
        return None
    return mask.to(torch.uint8).contiguous()


def layer_dropout_p(layer):
    """Sub-layer dropout probability of a (this file's or torch's) transformer layer holder, 0 in eval mode."""
    p = getattr(layer, "dropout_p", None)
    if p is None:
        d = getattr(layer, "dropout1", None)            # torch nn.TransformerEncoderLayer keeps nn.Dropout modules
        p = d.p if isinstance(d, nn.Dropout) else 0.1
    return RNG.p_of(layer.training, p)


def self_attention_block(mha, norm, x, B, L, pos, key_mask, p_drop=0.0, eps=LN_EPS):
    """LN(x + dropout(out_proj(MHA(q = k = x + pos, v = x)))); MHA drops attention probabilities with its own p."""
    E_ = mha.embed_dim
    H = mha.num_heads
    d = E_ // H
    p_attn = RNG.p_of(mha.training, mha.dropout)
    if pos is not None:
        xp = Fn.add_bcast(x, pos)
        qk = Fn.linear(xp, mha.in_proj_weight, mha.in_proj_bias, rows=(0, 2 * E_))
        v = Fn.linear(x, mha.in_proj_weight, mha.in_proj_bias, rows=(2 * E_, 3 * E_))
        ctx = Fn.attention(qk, qk, v, 0, E_, 0, B, H, L, L, d, key_mask=key_mask, mask_inf=True, dropout_p=p_attn)
    else:
        qkv = Fn.linear(x, mha.in_proj_weight, mha.in_proj_bias)
        ctx = Fn.attention(qkv, qkv, qkv, 0, E_, 2 * E_, B, H, L, L, d, key_mask=key_mask, mask_inf=True, dropout_p=p_attn)
    return Fn.linear_ln(ctx, x, mha.out_proj.weight, mha.out_proj.bias, norm.weight, norm.bias, eps, dropout_p=p_drop)


def layer_activation(layer):
    """ReLU (DETR, the discriminator's torch stacks) or exact GELU (the ViT stack, networks_vit.py:178) of a layer holder."""
    fn = getattr(layer, "activation", None)
    name = getattr(fn, "__name__", str(fn)).lower()
    return K.ACT_GELU if "gelu" in name else K.ACT_RELU


def ffn_block(layer, norm, x, p_drop=0.0, act=K.ACT_RELU, eps=LN_EPS):
    h = Fn.linear(x, layer.linear1.weight, layer.linear1.bias, act=act)
    h = Fn.dropout(h, p_drop)
    return Fn.linear_ln(h, x, layer.linear2.weight, layer.linear2.bias, norm.weight, norm.bias, eps, dropout_p=p_drop)


def encoder_layer_forward(layer, x, B, L, pos, key_mask):
    """Works for both this file's TransformerEncoderLayer and torch's nn.TransformerEncoderLayer holders."""
    p = layer_dropout_p(layer)
    x = self_attention_block(layer.self_attn, layer.norm1, x, B, L, pos, key_mask, p, eps=layer.norm1.eps)
    return ffn_block(layer, layer.norm2, x, p, act=layer_activation(layer), eps=layer.norm2.eps)


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        assert activation == "relu" and not normalize_before
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout_p = dropout


class TransformerDecoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        assert activation == "relu" and not normalize_before
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout_p = dropout

    def forward(self, tgt, memory, memory_pos, B, L, S, tgt_key_mask):
        """tgt [B*L, d], memory [B*S, d], memory_pos = memory + pos [B*S, d] (all bf16)."""
        p = layer_dropout_p(self)
        tgt = self_attention_block(self.self_attn, self.norm1, tgt, B, L, None, tgt_key_mask, p)
        mha = self.multihead_attn
        E_ = mha.embed_dim
        H = mha.num_heads
        d = E_ // H
        q = Fn.linear(tgt, mha.in_proj_weight, mha.in_proj_bias, rows=(0, E_))
        k = Fn.linear(memory_pos, mha.in_proj_weight, mha.in_proj_bias, rows=(E_, 2 * E_))
        v = Fn.linear(memory, mha.in_proj_weight, mha.in_proj_bias, rows=(2 * E_, 3 * E_))
        ctx = Fn.attention(q, k, v, 0, 0, 0, B, H, L, S, d, key_mask=None, mask_inf=True, dropout_p=RNG.p_of(mha.training, mha.dropout))
        tgt = Fn.linear_ln(ctx, tgt, mha.out_proj.weight, mha.out_proj.bias, self.norm2.weight, self.norm2.bias, LN_EPS, dropout_p=p)
        return ffn_block(self, self.norm3, tgt, p)


class TransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers, norm=None):
        super().__init__()
        self.layers = _get_clones(encoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm

    def forward(self, src, B, S, pos):
        for layer in self.layers:
            src = encoder_layer_forward(layer, src, B, S, pos, None)
        if self.norm is not None:
            src = Fn.layernorm(src, self.norm.weight, self.norm.bias, LN_EPS)
        return src


class TransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, norm=None, return_intermediate=False):
        super().__init__()
        assert not return_intermediate
        self.layers = _get_clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm

    def forward(self, tgt, memory, pos, B, L, S, tgt_key_mask):
        memory_pos = Fn.add_bcast(memory, pos)
        for layer in self.layers:
            tgt = layer(tgt, memory, memory_pos, B, L, S, tgt_key_mask)
        if self.norm is not None:
            tgt = Fn.layernorm(tgt, self.norm.weight, self.norm.bias, LN_EPS)
        return tgt


class Transformer(nn.Module):
    def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048,
                 dropout=0.1, activation="relu", normalize_before=False, return_intermediate_dec=False):
        super().__init__()
        enc = TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.encoder = TransformerEncoder(enc, num_encoder_layers, None)
        dec = TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.decoder = TransformerDecoder(dec, num_decoder_layers, nn.LayerNorm(d_model), return_intermediate_dec)
        self._reset_parameters()
        self.d_model = d_model
        self.nhead = nhead

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def _prepend(self, tgt, tgt_key_padding_mask, B, L):
        return tgt, tgt_key_padding_mask, L

    def forward(self, src, pos, tgt, tgt_key_padding_mask, B, S, L):
        """src bf16 [B*S, d] (projected image tokens), pos fp32 [S, d], tgt bf16 [B*L, d],
        tgt_key_padding_mask bool [B, L] -> hs bf16 [B*L', d] (L' = L, or L+1 with the learned token)."""
        memory = self.encoder(src, B, S, pos)
        tgt, mask, L2 = self._prepend(tgt, tgt_key_padding_mask, B, L)
        hs = self.decoder(tgt, memory, pos, B, L2, S, _key_mask(mask))
        return hs, memory, L2


class TransformerWithToken(Transformer):
    """Prepends a learned token to the decoder input (reference training/detr_transformer.py:22-70)."""

    def __init__(self, *args, **kw):
        nn.Module.__init__(self)
        d_model = kw.get("d_model", args[0] if args else 512)
        self.token = nn.Parameter(torch.randn(1, 1, d_model))
        self.register_buffer("token_mask", torch.zeros(1, 1, dtype=torch.bool))
        tmp = Transformer(*args, **kw)
        self.encoder, self.decoder = tmp.encoder, tmp.decoder
        self._reset_parameters()
        self.d_model, self.nhead = tmp.d_model, tmp.nhead

    def _prepend(self, tgt, tgt_key_padding_mask, B, L):
        d = self.d_model
        tok = Fn.to_bf16_padded(self.token.view(1, d)).view(1, 1, d).expand(B, 1, d)
        tgt = torch.cat([tok, tgt.view(B, L, d)], dim=1).reshape(B * (L + 1), d)
        mask = torch.cat([self.token_mask.expand(B, -1), tgt_key_padding_mask], dim=1)
        return tgt, mask, L + 1


class TransformerEncoderStack(nn.Module):
    """Runs a torch `nn.TransformerEncoder` holder (post-norm, ReLU) through the sm_100a kernels.
    x bf16 [B*L, d] batch-major; key_padding_mask bool [B, L]."""

    @staticmethod
    def run(encoder, x, B, L, key_padding_mask):
        km = _key_mask(key_padding_mask)
        for layer in encoder.layers:
            x = encoder_layer_forward(layer, x, B, L, None, km)
        if getattr(encoder, "norm", None) is not None:
            x = Fn.layernorm(x, encoder.norm.weight, encoder.norm.bias, encoder.norm.eps)
        return x
