"""BLIP "MED" BERT text encoder / causal LM decoder on the sm_100a kernels.

Mirror of the reference's training/med.py for the code path LayoutDETR actually takes
(`mode='text'`: self-attention only, cross-attention parameters exist but are never used,
training/med.py:361).  Module / parameter names reproduce the reference state_dict tree
(`embeddings.word_embeddings.weight`, `encoder.layer.{i}.attention.self.query.weight`, ...,
`cls.predictions.decoder.weight` tied to the word embeddings) so checkpoints load by name.

Dropout (hidden 0.1 / attention-probability 0.1, configs/med_config.json:5,7; reference training/med.py:96,213,240,318)
is live whenever the module is in `.train()` mode — as in the reference's training loop, also inside the frozen text
encoder (training/training_loop.py:133-134) — and is drawn in-kernel from the Philox stream of layoutdetr_b200.rng.
`.eval()` (G_ema, metrics, parity tests) gives the deterministic path.
"""
import json
import math
import os
import types
import warnings

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import kernels as K
from .. import rng as RNG


class BertConfig:
    """Plain-attribute stand-in for transformers.BertConfig (reference training/med.py:27)."""

    def __init__(self, **kw):
        self.vocab_size = 30522
        self.hidden_size = 768
        self.num_hidden_layers = 12
        self.num_attention_heads = 12
        self.intermediate_size = 3072
        self.hidden_act = "gelu"
        self.hidden_dropout_prob = 0.1
        self.attention_probs_dropout_prob = 0.1
        self.max_position_embeddings = 512
        self.layer_norm_eps = 1e-12
        self.pad_token_id = 0
        self.initializer_range = 0.02
        self.encoder_width = 768
        self.add_cross_attention = True
        self.__dict__.update(kw)

    @classmethod
    def from_json_file(cls, path):
        with open(path) as f:
            return cls(**json.load(f))

    @classmethod
    def default(cls):
        return cls()


# ------------------------------------------------------------------------------------------------
# Pretrained weights.  The reference builds both BERT stacks with `from_pretrained('bert-base-uncased', config=...)`
# (training/networks_detr.py:92,124; training/med.py via transformers.PreTrainedModel) and then freezes the text encoder
# (training/training_loop.py:283): without the checkpoint G and D would train against a frozen RANDOM encoder.  So:
# a local Hugging Face checkpoint (a directory, or the hub cache, never the network) is loaded by parameter name — the
# module tree mirrors the reference's — and its absence is an error unless synthetic weights were asked for explicitly
# (LAYOUTDETR_SYNTHETIC_WEIGHTS=1; LAYOUTDETR_SYNTHETIC_TOKENIZER=1 implies it: tests, bench.py, smoke()).
def synthetic_weights_allowed():
    return os.environ.get("LAYOUTDETR_SYNTHETIC_WEIGHTS", "0") == "1" or os.environ.get("LAYOUTDETR_SYNTHETIC_TOKENIZER", "0") == "1"


def _find_checkpoint_file(name):
    names = ("model.safetensors", "pytorch_model.bin")
    if os.path.isdir(name):
        for fn in names:
            if os.path.exists(os.path.join(name, fn)):
                return os.path.join(name, fn)
        return None
    try:
        from transformers.utils import cached_file
    except Exception:
        return None
    for fn in names:
        try:
            path = cached_file(name, fn, local_files_only=True, _raise_exceptions_for_missing_entries=False,
                               _raise_exceptions_for_connection_errors=False)
        except Exception:
            path = None
        if path and os.path.exists(path):
            return path
    return None


def _read_checkpoint(path):
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


def load_pretrained_bert(model, name, strip_prefix):
    """Copy every tensor of the local checkpoint `name` whose (normalised) key and shape match a parameter / buffer of
    `model`; -> number of tensors loaded.  `strip_prefix`: 'bert.' for the bare encoder (BertModel), '' for the LM-head model."""
    path = _find_checkpoint_file(name)
    if path is None:
        if synthetic_weights_allowed():
            return 0
        raise FileNotFoundError(
            "pretrained checkpoint %r not found locally (a directory with model.safetensors / pytorch_model.bin, or the Hugging Face "
            "cache). The reference initialises the BERT text encoder / decoder from it and freezes the encoder; refusing to fall "
            "back to random weights silently. Provide the checkpoint, pass --resume with a trained pickle after constructing with "
            "LAYOUTDETR_SYNTHETIC_WEIGHTS=1, or set LAYOUTDETR_SYNTHETIC_WEIGHTS=1 for synthetic-weight runs." % (name,))
    src = _read_checkpoint(path)
    own = dict(model.state_dict())
    loaded = 0
    with torch.no_grad():
        for k, v in src.items():
            k = k.replace(".gamma", ".weight").replace(".beta", ".bias")             # TF-converted BERT checkpoints
            if strip_prefix and k.startswith(strip_prefix):
                k = k[len(strip_prefix):]
            elif strip_prefix == "" and not (k.startswith("bert.") or k.startswith("cls.")):
                k = "bert." + k                                                          # bare-encoder checkpoint into the LM model
            t = own.get(k)
            if t is not None and tuple(t.shape) == tuple(v.shape):
                t.copy_(v.to(t.dtype))
                loaded += 1
    if loaded == 0:
        raise RuntimeError("checkpoint %s has no tensor matching %s" % (path, type(model).__name__))
    missing = [k for k in own if "crossattention" not in k and "position_ids" not in k]
    if loaded < len(missing) // 2:
        warnings.warn("only %d of %d tensors of %s were found in %s" % (loaded, len(missing), type(model).__name__, path))
    return loaded


class BertEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))
        self.eps = config.layer_norm_eps
        self.pad_token_id = config.pad_token_id
        self.hidden_dropout_prob = config.hidden_dropout_prob

    def forward(self, input_ids):
        B, T = input_ids.shape
        y = Fn.EmbedLNFn.apply(input_ids.contiguous(), self.word_embeddings.weight, self.position_embeddings.weight,
                               self.LayerNorm.weight, self.LayerNorm.bias, T, self.eps, self.pad_token_id)
        return Fn.dropout(y, RNG.p_of(self.training, getattr(self, "hidden_dropout_prob", 0.1)))     # training/med.py:96


class BertSelfAttention(nn.Module):
    def __init__(self, config, is_cross_attention):
        super().__init__()
        kv_in = config.encoder_width if is_cross_attention else config.hidden_size
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        self.query = nn.Linear(config.hidden_size, config.hidden_size)
        self.key = nn.Linear(kv_in, config.hidden_size)
        self.value = nn.Linear(kv_in, config.hidden_size)


class BertSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertAttention(nn.Module):
    def __init__(self, config, is_cross_attention=False):
        super().__init__()
        self.self = BertSelfAttention(config, is_cross_attention)
        self.output = BertSelfOutput(config)


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class BertOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertLayer(nn.Module):
    def __init__(self, config, layer_num):
        super().__init__()
        self.attention = BertAttention(config)
        self.layer_num = layer_num
        if config.add_cross_attention:
            self.crossattention = BertAttention(config, is_cross_attention=True)   # parameters only (mode='text')
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)
        self.eps = config.layer_norm_eps
        self.hidden_dropout_prob = config.hidden_dropout_prob
        self.attention_probs_dropout_prob = config.attention_probs_dropout_prob

    def forward(self, x, B, T, key_mask, causal):
        """x: bf16 [B*T, hidden].  Post-norm BERT layer (reference training/med.py:336-386)."""
        sa = self.attention.self
        H, d = sa.num_attention_heads, sa.attention_head_size
        p_h = RNG.p_of(self.training, getattr(self, "hidden_dropout_prob", 0.1))                     # :240, :318
        p_a = RNG.p_of(self.training, getattr(self, "attention_probs_dropout_prob", 0.1))            # :213
        qkv = Fn.fused_linear(x, (sa.query, sa.key, sa.value))
        ctx = Fn.attention(qkv, qkv, qkv, 0, H * d, 2 * H * d, B, H, T, T, d, 1.0 / math.sqrt(d),
                           key_mask=key_mask, mask_inf=False, causal=causal, dropout_p=p_a)
        so = self.attention.output
        h = Fn.linear_ln(ctx, x, so.dense.weight, so.dense.bias, so.LayerNorm.weight, so.LayerNorm.bias, self.eps, dropout_p=p_h)
        inter = Fn.linear(h, self.intermediate.dense.weight, self.intermediate.dense.bias, act=K.ACT_GELU)
        o = self.output
        return Fn.linear_ln(inter, h, o.dense.weight, o.dense.bias, o.LayerNorm.weight, o.LayerNorm.bias, self.eps, dropout_p=p_h)


class BertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(config, i) for i in range(config.num_hidden_layers)])


def _init_bert_weights(module, std):
    if isinstance(module, (nn.Linear, nn.Embedding)):
        module.weight.data.normal_(mean=0.0, std=std)
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)
    if isinstance(module, nn.Linear) and module.bias is not None:
        module.bias.data.zero_()


class BertModel(nn.Module):
    """Text encoder (`add_pooling_layer=False`).  forward(...) keeps the reference call form
    `text_encoder(input_ids, attention_mask=..., return_dict=True, mode='text')`."""

    def __init__(self, config, add_pooling_layer=False):
        super().__init__()
        assert not add_pooling_layer
        self.config = config
        self.embeddings = BertEmbeddings(config)
        self.encoder = BertEncoder(config)
        self.apply(lambda m: _init_bert_weights(m, config.initializer_range))

    @classmethod
    def from_pretrained(cls, name, config=None, **kw):
        """Local checkpoint by parameter name (see load_pretrained_bert); random init only under LAYOUTDETR_SYNTHETIC_WEIGHTS=1."""
        model = cls(config, **kw)
        load_pretrained_bert(model, name, "bert.")
        return model

    def resize_token_embeddings(self, n):
        old = self.embeddings.word_embeddings
        if n == old.num_embeddings:
            return old
        new = nn.Embedding(n, old.embedding_dim, padding_idx=old.padding_idx)
        new.weight.data.normal_(0.0, self.config.initializer_range)
        k = min(n, old.num_embeddings)
        new.weight.data[:k] = old.weight.data[:k]
        self.embeddings.word_embeddings = new
        self.config.vocab_size = n
        return new

    def hidden_states(self, input_ids, attention_mask, causal=False):
        """-> bf16 [B*T, hidden] last-layer hidden states."""
        B, T = input_ids.shape
        key_mask = (attention_mask == 0).to(torch.uint8).contiguous()
        x = self.embeddings(input_ids)
        for layer in self.encoder.layer:
            x = layer(x, B, T, key_mask, causal)
        return x

    def forward(self, input_ids, attention_mask=None, return_dict=True, mode="text", is_decoder=False, **_):
        if mode != "text":
            raise NotImplementedError("only mode='text' is on the LayoutDETR path (training/networks_detr.py:146)")
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        B, T = input_ids.shape
        x = self.hidden_states(input_ids, attention_mask, causal=is_decoder)
        return types.SimpleNamespace(last_hidden_state=Fn.to_f32(x).view(B, T, -1))

    def cls_features(self, input_ids, attention_mask):
        """-> bf16 [B, hidden]: last_hidden_state[:, 0, :] (the only part of the encoder output LayoutDETR reads)."""
        B, T = input_ids.shape
        x = self.hidden_states(input_ids, attention_mask)
        return x.view(B, T, -1)[:, 0, :]


class BertPredictionHeadTransform(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertLMPredictionHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))
        self.decoder.bias = self.bias


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.predictions = BertLMPredictionHead(config)


class BertLMHeadModel(nn.Module):
    """Causal LM text decoder; forward(...) returns an object with `.loss` (label-smoothed next-token CE,
    reference training/med.py:833-933)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.bert = BertModel(config, add_pooling_layer=False)
        self.cls = BertOnlyMLMHead(config)
        self.apply(lambda m: _init_bert_weights(m, config.initializer_range))
        self._tie()

    def _tie(self):
        self.cls.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight

    @classmethod
    def from_pretrained(cls, name, config=None, **kw):
        model = cls(config, **kw)
        load_pretrained_bert(model, name, "")
        model._tie()
        return model

    def resize_token_embeddings(self, n):
        old_n = self.bert.embeddings.word_embeddings.num_embeddings
        new = self.bert.resize_token_embeddings(n)
        self.config.vocab_size = n
        pred = self.cls.predictions
        if n != old_n:
            dec = nn.Linear(new.embedding_dim, n, bias=False)
            bias = nn.Parameter(torch.zeros(n))
            k = min(n, old_n)
            bias.data[:k] = pred.bias.data[:k]
            pred.bias = bias
            dec.bias = bias
            pred.decoder = dec
        self._tie()
        return new

    def forward(self, input_ids, attention_mask=None, labels=None, return_dict=True, mode="text", n_valid=None, **_):
        if mode != "text":
            raise NotImplementedError("only mode='text' is on the LayoutDETR path (training/networks_detr.py:180,339)")
        B, T = input_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        x = self.bert.hidden_states(input_ids, attention_mask, causal=True)
        tr = self.cls.predictions.transform
        h = Fn.linear(x, tr.dense.weight, tr.dense.bias, act=K.ACT_GELU)
        h = Fn.layernorm(h, tr.LayerNorm.weight, tr.LayerNorm.bias, self.config.layer_norm_eps)
        if labels is None:
            raise NotImplementedError("logits-only decoding is not on the training path")
        # next-token prediction: position t predicts labels[t+1]; the last position has no target
        shifted = torch.full_like(labels, -100)
        shifted[:, :-1] = labels[:, 1:]
        if n_valid is None:
            inv = 1.0 / (shifted != -100).sum().clamp(min=1).float().reshape(1)          # device-side, no sync
        elif torch.is_tensor(n_valid):
            inv = n_valid                                                                  # already 1 / count (device scalar)
        else:
            inv = torch.full((1,), 1.0 / max(1, int(n_valid)), dtype=torch.float32, device=h.device)
        loss = Fn.lm_head_ce(h, self.bert.embeddings.word_embeddings.weight, self.cls.predictions.bias,
                             shifted.reshape(-1).contiguous(), inv, 0.1)
        return types.SimpleNamespace(loss=loss)
