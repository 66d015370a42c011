"""Deterministic synthetic weights, inputs and tokenizer (no network: there are no pretrained
checkpoints, no bert-base-uncased vocab and no dataset on the build / GPU boxes).

Everything is a pure function of (name, shape, seed) computed with the CPU generator, so the
reference run in the build container (tests/golden/gen_golden.py), the CPU oracle and the CUDA path on
the GPU box all see bit-identical parameters and inputs.
"""
import zlib

import numpy as np
import torch


# ---------------------------------------------------------------------------------------------
# tokenizer stand-in for BertTokenizer('bert-base-uncased') + [DEC] / [ENC]
# (reference: training/blip.py:190-195; call site training/networks_detr.py:145,289)
# ---------------------------------------------------------------------------------------------
class _Encoding(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def to(self, device):
        return _Encoding({k: v.to(device) for k, v in self.items()})


class SyntheticTokenizer:
    """Character-level tokenizer with BERT's special-token layout.

    ids = [CLS=101] + [1000 + (ord(c) * 7919) % 28000 for c in text] + [SEP=102], padded with 0.
    len() == 30524 (30522 + [DEC], [ENC]); bos_token_id = 30522 ([DEC]), enc_token_id = 30523.
    """
    pad_token_id = 0
    cls_token_id = 101
    sep_token_id = 102
    bos_token_id = 30522
    enc_token_id = 30523

    def __len__(self):
        return 30524

    def encode_one(self, text, max_length):
        body = [1000 + (ord(c) * 7919) % 28000 for c in text][: max(0, max_length - 2)]
        return [self.cls_token_id] + body + [self.sep_token_id]

    def __call__(self, texts, padding="max_length", truncation=True, max_length=256, return_tensors="pt"):
        if isinstance(texts, str):
            texts = [texts]
        rows = [self.encode_one(t, max_length) for t in texts]
        width = max_length if padding == "max_length" else max(len(r) for r in rows)
        ids = np.zeros((len(rows), width), dtype=np.int64)
        mask = np.zeros((len(rows), width), dtype=np.int64)
        for i, r in enumerate(rows):
            ids[i, : len(r)] = r
            mask[i, : len(r)] = 1
        return _Encoding(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask))


# ---------------------------------------------------------------------------------------------
# weights by name
# ---------------------------------------------------------------------------------------------
def _gen(name, seed):
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


def synth_tensor(name, shape, seed=0):
    """Deterministic fp32 value for a parameter / buffer called `name` (state_dict key)."""
    g = _gen(name, seed)
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    r = lambda s=1.0: torch.randn(shape, generator=g) * s
    if name.endswith("position_ids"):
        return torch.arange(shape[-1]).expand(shape).clone()
    if name.endswith("token_mask"):
        return torch.zeros(shape, dtype=torch.bool)
    if leaf == "running_var":
        return 1.0 + 0.2 * torch.rand(shape, generator=g)
    if leaf == "running_mean":
        return r(0.1)
    if "bg_decoder" in name:
        if leaf == "resample_filter":
            f = torch.tensor([1.0, 3.0, 3.0, 1.0])
            f = torch.outer(f, f)
            return f / f.sum()
        if leaf == "w_avg":
            return torch.zeros(shape)
        if ".mapping." in name and leaf == "weight":
            return r(100.0)                       # randn / lr_multiplier(0.01), networks_stylegan2.py:108
        if ".affine." in name and leaf == "bias":
            return 1.0 + r(0.05)                  # bias_init=1
        if leaf == "bias":
            return r(0.05)
        return r(1.0)                             # conv weights / const: randn
    is_norm = ("norm" in name.lower() or ".bn" in name or "downsample.1" in name) and len(shape) == 1
    if is_norm and leaf == "weight":
        if ".bn3." in name:
            # last FrozenBN of a bottleneck: a small gain keeps the 16 residual additions of ResNet-50 from doubling the activation
            # variance per block (with gain 1 the image tokens reach |x| ~ 1e3, DETR encoder layer 0 sees attention scores of 1e6 and
            # its softmax is exactly one-hot: every gradient upstream of it is then rounding noise in ANY bf16 implementation)
            return 0.1 + r(0.01)
        return 1.0 + r(0.1)
    if leaf == "bias" or (is_norm and leaf == "bias"):
        return r(0.02)
    if "embeddings" in name or name.startswith(("emb_", "enc_text_len")) or ".emb_" in name:
        return r(0.02) if "text_" in name.split(".")[0] else r(0.5)
    if len(shape) == 4:                           # conv OIHW
        fan_in = shape[1] * shape[2] * shape[3]
        return r(float(np.sqrt(2.0 / fan_in)))
    if len(shape) == 2:                           # linear [out, in]
        return r(float(1.0 / np.sqrt(shape[1])))
    if len(shape) == 3:                           # tokens [.., 1, d]
        return r(0.5)
    return r(0.02)


def synth_state_dict(module, seed=0):
    """In-place: overwrite every parameter and buffer of `module` with synth_tensor(name, shape)."""
    sd = module.state_dict()
    out = {}
    for name, t in sd.items():
        v = synth_tensor(name, t.shape, seed)
        out[name] = v.to(t.dtype)
    # tied LM-head decoder weight <-> word embeddings (BertLMHeadModel, transformers 4.19 tie_weights)
    for name in list(out):
        if name.endswith("cls.predictions.decoder.weight"):
            emb = name.replace("cls.predictions.decoder.weight", "bert.embeddings.word_embeddings.weight")
            if emb in out:
                out[name] = out[emb]
        if name.endswith("cls.predictions.decoder.bias"):
            b = name.replace("cls.predictions.decoder.bias", "cls.predictions.bias")
            if b in out:
                out[name] = out[b]
    module.load_state_dict(out, strict=True)
    return out


# ---------------------------------------------------------------------------------------------
# inputs
# ---------------------------------------------------------------------------------------------
_ALPHABET = "abcdefghijklmnopqrstuvwxyz ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789%$!.,-"


def make_inputs(batch, n_valid=8, n_slots=9, background_size=256, z_dim=4, num_bbox_labels=8, seed=1, device="cpu"):
    """Synthetic batch with the reference's tensor contract (training/networks_detr.py:133,279)."""
    g = torch.Generator().manual_seed(seed)
    B, N = batch, n_slots
    background = torch.randn((B, 3, background_size, background_size), generator=g)
    u = lambda lo, hi: torch.rand((B, N), generator=g) * (hi - lo) + lo
    bbox_real = torch.stack([u(0.2, 0.8), u(0.2, 0.8), u(0.1, 0.6), u(0.03, 0.2)], dim=-1)
    padding_mask = torch.zeros((B, N), dtype=torch.bool)
    padding_mask[:, n_valid:] = True
    bbox_real = bbox_real * (~padding_mask).unsqueeze(-1)
    bbox_class = torch.randint(0, num_bbox_labels, (B, N), generator=g)
    bbox_class = bbox_class * (~padding_mask)
    lens = torch.randint(4, 41, (B, N), generator=g)
    chars = torch.randint(0, len(_ALPHABET), (B, N, 40), generator=g)
    bbox_text = [["".join(_ALPHABET[int(c)] for c in chars[b, n, : int(lens[b, n])]) if n < n_valid else ""
                  for n in range(N)] for b in range(B)]
    z = torch.randn((B, N, z_dim), generator=g)
    bbox_patch = torch.zeros((B, N, 3, 1, 1))
    c = torch.zeros((B, 0))
    d = dict(z=z, bbox_class=bbox_class, bbox_real=bbox_real, bbox_text=bbox_text, bbox_patch=bbox_patch,
             padding_mask=padding_mask, background=background, c=c)
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in d.items()}


def make_ragged_inputs(n_valid_per_sample, seed=1, **kw):
    """Like make_inputs, but with a different number of real elements in every layout (what a real dataset batch looks like):
    padded slots get zero boxes / class 0 / the empty string and `padding_mask` True."""
    B = len(n_valid_per_sample)
    d = make_inputs(B, n_valid=max(n_valid_per_sample), seed=seed, **kw)
    N = d["padding_mask"].shape[1]
    pad = torch.arange(N)[None, :] >= torch.tensor(list(n_valid_per_sample))[:, None]
    d["padding_mask"] = pad
    d["bbox_real"] = d["bbox_real"] * (~pad).unsqueeze(-1)
    d["bbox_class"] = d["bbox_class"] * (~pad)
    d["bbox_text"] = [[t if not bool(pad[b, n]) else "" for n, t in enumerate(row)] for b, row in enumerate(d["bbox_text"])]
    return d
