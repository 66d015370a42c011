"""LayoutDETR Generator / Discriminator on the sm_100a kernels — drop-in mirror of the reference's
training/networks_detr.py (Generator :65-187, Discriminator :190-361): same constructor signatures,
forward signatures / return tuples, attribute names and state_dict keys, so `training_loop`, the
metrics and `generate.py` can use these classes unchanged and reference checkpoints load by name.

All heavy math (ResNet-50 convs, BERT encoder/decoder, DETR attention + FFN, LM head + loss,
StyleGAN2 background decoder) runs in hand-written CUDA through layoutdetr_b200.functional; PyTorch
ops appear only as glue on tiny tensors (concatenation / indexing of `[B*9, d]` rows, scalar losses).
There is no CPU path: inputs must live on a CUDA device.
"""
import collections
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as Fn
from .. import kernels as K
from ..lanes import LANES
from .detr_backbone import build_backbone as _build_backbone
from .detr_transformer import Transformer, TransformerWithToken, TransformerEncoderStack
from .med import BertConfig, BertModel, BertLMHeadModel
from .util import TransformerWithToken_layoutganpp


def merge_lists(lists):
    ret = []
    for l in lists:
        ret += l
    return ret


def split_list(list_a, chunk_size):
    return [list_a[i:i + chunk_size] for i in range(0, len(list_a), chunk_size)]


def normalize_2nd_moment(x, eps=1e-8):
    return x * (x.square().mean(dim=1, keepdim=True) + eps).rsqrt()


def build_backbone(kind="resnet50", background_size=256, hidden_dim=256):
    """'resnet50': ResNet-50 + FrozenBN + sine position embedding (reference training/detr_backbone.py:98-115, what the reference
    wires).  'vit_b16': the ViT-B/16 encoder of training/networks_vit.py behind the same interface."""
    if kind == "resnet50":
        return _build_backbone()
    if kind == "vit_b16":
        from .networks_vit import build_vit_backbone
        return build_vit_backbone(background_size, background_size, hidden_dim)
    raise ValueError("unknown backbone %r (resnet50 | vit_b16)" % (kind,))


def init_tokenizer():
    """BertTokenizer('bert-base-uncased') + [DEC] / [ENC] from a LOCAL vocabulary (reference training/blip.py:190-195).
    LAYOUTDETR_SYNTHETIC_TOKENIZER=1 selects the deterministic synthetic tokenizer (same special-token layout, ids unrelated
    to any real checkpoint) — an explicit opt-in: a missing vocabulary is an error, never a silent change of token ids."""
    if os.environ.get("LAYOUTDETR_SYNTHETIC_TOKENIZER", "0") == "1":
        from ..synthetic import SyntheticTokenizer
        return SyntheticTokenizer()
    try:
        from transformers import BertTokenizer
        tok = BertTokenizer.from_pretrained("bert-base-uncased", local_files_only=True)
        if tok.vocab_size != 30522:          # transformers >= 5 hands back an EMPTY tokenizer when no vocabulary file is cached
            raise FileNotFoundError("vocab.txt of bert-base-uncased not found (tokenizer has %d entries)" % tok.vocab_size)
    except Exception as e:
        raise RuntimeError("the bert-base-uncased vocabulary is not available locally (%s: %s). Token ids must match the checkpoint "
                           "the text encoder was trained with; set LAYOUTDETR_SYNTHETIC_TOKENIZER=1 only for synthetic-weight runs "
                           "(tests, bench.py)." % (type(e).__name__, e)) from e
    tok.add_special_tokens({"bos_token": "[DEC]"})
    tok.add_special_tokens({"additional_special_tokens": ["[ENC]"]})
    tok.enc_token_id = tok.convert_tokens_to_ids("[ENC]")
    return tok


def _med_config(path):
    if path is not None and os.path.exists(path):
        return BertConfig.from_json_file(path)
    return BertConfig.default()       # identical to the reference's configs/med_config.json


class MLP(nn.Module):
    """Linear/ReLU stack (reference training/networks_detr.py:50-62); `final_act` fuses the caller's
    trailing `torch.relu(...)` into the last GEMM epilogue."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x, final_act=K.ACT_NONE, out_f32=False):
        for i, layer in enumerate(self.layers):
            last = i == self.num_layers - 1
            if last and out_f32:
                x = Fn.linear_f32(x, layer.weight, layer.bias)
            else:
                x = Fn.linear(x, layer.weight, layer.bias, act=final_act if last else K.ACT_RELU)
        return x


class _TextFrontEnd:
    """Host-side text handling shared by G and D: tokenise once per distinct batch of strings, keep the
    device copies, and remember the host-side facts (valid slots, token counts) so the forward pass never
    synchronises on the device to learn a shape."""

    def __init__(self, tokenizer, max_text_length):
        self.tokenizer = tokenizer
        self.max_text_length = max_text_length
        self._key = None
        self._val = None

    def __call__(self, bbox_text, device):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        flat = merge_lists(bbox_text)
        key = (tuple(flat), str(device))
        if key != self._key:
            enc = self.tokenizer(flat, padding="max_length", truncation=True, max_length=self.max_text_length,
                                 return_tensors="pt")
            ids, mask = enc.input_ids, enc.attention_mask
            text_len = torch.from_numpy(np.array([len(t) for t in flat], dtype=np.int64))
            old = self._val
            if old is not None and old["ids"].shape == ids.shape and old["ids"].device == device:
                # refresh in place: device addresses stay stable (a captured CUDA graph keeps reading these buffers).  The sources
                # are pinned and the copies asynchronous: a copy from pageable memory blocks the host until it has run, i.e. until
                # the previous iteration's graph replay has finished, and the next replay could only be queued after that
                # (measured as 3-4 % between the bare replay and the training_loop entry point).
                if device.type == "cuda":
                    src = [t.pin_memory() for t in (ids, mask, text_len)]
                    old["ids"].copy_(src[0], non_blocking=True); old["mask"].copy_(src[1], non_blocking=True)
                    old["text_len"].copy_(src[2], non_blocking=True)
                else:
                    old["ids"].copy_(ids); old["mask"].copy_(mask); old["text_len"].copy_(text_len)
                old["ids_cpu"], old["mask_cpu"] = ids, mask
            else:
                self._val = dict(ids=ids.to(device), mask=mask.to(device), ids_cpu=ids, mask_cpu=mask,
                                 text_len=text_len.to(device))
            self._key = key
        return self._val


_host_mask_cache = {}


def _host_mask(padding_mask):
    """CPU copy of the (tiny) padding mask, cached per tensor version: one D2H sync per new batch."""
    key = (padding_mask.data_ptr(), padding_mask._version, tuple(padding_mask.shape))
    v = _host_mask_cache.get(key)
    if v is None:
        if len(_host_mask_cache) > 16:
            _host_mask_cache.clear()
        v = padding_mask.detach().cpu().numpy().astype(bool)
        _host_mask_cache[key] = v
    return v


_valid_idx_cache = {}
_ITERATION = [0]


def new_iteration():
    """Marks the start of a training iteration: per-iteration caches (frozen text-encoder features under `text_dedup`)
    are only valid within one iteration."""
    _ITERATION[0] += 1


def valid_index(padding_mask):
    """(device int64 indices, numpy indices) of the valid (non-padding) slots of a `[B, N]` mask, row-major —
    the sync-free equivalent of boolean indexing with `~padding_mask`."""
    key = (padding_mask.data_ptr(), padding_mask._version, tuple(padding_mask.shape))
    v = _valid_idx_cache.get(key)
    if v is None:
        if len(_valid_idx_cache) > 16:
            _valid_idx_cache.clear()
        cpu = np.flatnonzero(~_host_mask(padding_mask).reshape(-1))
        v = (torch.from_numpy(cpu).to(padding_mask.device), cpu)
        _valid_idx_cache[key] = v
    return v


def _encode_text_now(module, text):
    """Runs the text encoder on the current stream -> bf16 [B*N, bert_f_dim] CLS features."""
    enc = module.text_encoder
    ids, mask = text["ids"], text["mask"]
    if module.text_trim:
        T = int(text["mask_cpu"].sum(1).max())
        T = min(ids.shape[1], (T + 7) // 8 * 8)
        ids, mask = ids[:, :T].contiguous(), mask[:, :T].contiguous()
    frozen = not any(p.requires_grad for p in enc.parameters())
    if module.text_dedup and frozen:
        key = (_ITERATION[0], id(text["ids"]), ids.shape[1], tuple(p._version for p in enc.parameters()))
        if module._cls_cache is not None and module._cls_cache[0] == key:
            return module._cls_cache[1]
        with torch.no_grad():
            feat = enc.cls_features(ids, mask).contiguous()
        module._cls_cache = (key, feat)
        return feat
    return enc.cls_features(ids, mask).contiguous()


def trimmed_widths(module, bbox_text, padding_mask_cpu, device):
    """(encoder width, decoder width) `text_trim` would use for this batch — host-side, from the tokenizer's attention mask
    (the front end's cached copy); the key of anything that bakes these widths in (trainer.GraphedStep)."""
    if not module.text_trim:
        return (module.max_text_length, module.max_text_length)
    mask = module._front()(bbox_text, device)["mask_cpu"]
    T = mask.shape[1]
    t_enc = min(T, (int(mask.sum(1).max()) + 7) // 8 * 8)
    valid = np.flatnonzero(~padding_mask_cpu.numpy().astype(bool).reshape(-1))
    t_dec = min(T, (int(mask[valid].sum(1).max()) + 7) // 8 * 8) if len(valid) else T
    return (t_enc, t_dec)


def prefetch_text(modules, bbox_text, device):
    """Lane scheduling (lanes.py, level 1): issue the frozen text-encoder call of every module in `modules` (one entry
    per forward pass that will follow, in the order the passes are issued on the host) on the T lane.  Each module's
    `_encode_text` then pops its result and makes the consuming stream wait for that call only.  The encoder's only
    inputs are the token ids (reference training/networks_detr.py:145-147), so hoisting the calls changes nothing."""
    texts = {id(m): m._front()(bbox_text, device) for m in modules}          # token ids reach the device on this stream
    sT = LANES.fork("T", cta_limit=LANES.text_ctas, bulk=True, detached=True)
    with torch.cuda.stream(sT), torch.no_grad():
        for m in modules:
            if any(p.requires_grad for p in m.text_encoder.parameters()):
                raise RuntimeError("prefetch_text needs a frozen text encoder (training/training_loop.py:283)")
            feat = _encode_text_now(m, texts[id(m)])
            ev = torch.cuda.Event()
            ev.record(sT)
            m.__dict__.setdefault("_te_queue", collections.deque()).append([feat, ev])
    return sT


def pending_text(modules):
    return sum(len(m.__dict__.get("_te_queue", ())) for m in modules)


def _encode_text(module, text, B, N):
    """CLS feature of the frozen/unfrozen text encoder for every slot -> bf16 [B*N, bert_f_dim]."""
    q = module.__dict__.get("_te_queue")
    if q:
        feat, ev = q.popleft()
        cur = torch.cuda.current_stream()
        if ev is not None:
            cur.wait_event(ev)
        feat.record_stream(cur)
        return feat
    return _encode_text_now(module, text)


def lm_target_count(text, valid_idx_cpu, pad_id, trim=False):
    """Number of next-token targets of the valid slots (host-side; what the LM loss is averaged over)."""
    ids_cpu = text["ids_cpu"][valid_idx_cpu]
    return int(((ids_cpu != pad_id)[:, 1:]).sum())


def _decode_text_loss(module, text, valid_idx_dev, valid_idx_cpu, bos_id, pad_id):
    """Causal-LM reconstruction loss of the strings of the valid slots (reference :169-181 / :328-340).
    The decoder runs mode='text': `encoder_hidden_states` is never read (training/med.py:361)."""
    ids = text["ids"].index_select(0, valid_idx_dev).clone()
    mask = text["mask"].index_select(0, valid_idx_dev)
    mask_cpu = text["mask_cpu"][valid_idx_cpu]
    if module.text_trim:
        T = int(mask_cpu.sum(1).max())
        T = min(ids.shape[1], (T + 7) // 8 * 8)
        ids, mask, mask_cpu = ids[:, :T].contiguous(), mask[:, :T].contiguous(), mask_cpu[:, :T]
    ids[:, 0] = bos_id
    labels = ids.masked_fill(ids == pad_id, -100)
    # next-token targets: positions 1.. of every non-pad token
    ids_cpu = text["ids_cpu"][valid_idx_cpu][:, :ids.shape[1]]
    n_valid = int(((ids_cpu != pad_id)[:, 1:]).sum())
    # 1 / #targets lives in a persistent device scalar: under CUDA-graph capture it is NOT rewritten here (the owner of the
    # graph refreshes it before every replay, see trainer.GraphedStep), so the captured kernels stay valid for any batch.
    buf = module.__dict__.get("_inv_n_valid")
    if buf is None or buf.device != ids.device:
        buf = torch.zeros(1, dtype=torch.float32, device=ids.device)
        module.__dict__["_inv_n_valid"] = buf
    if not torch.cuda.is_current_stream_capturing():
        buf.fill_(1.0 / max(1, n_valid))
    out = module.text_decoder(ids, attention_mask=mask, labels=labels, return_dict=True, mode="text", n_valid=buf)
    return out.loss


class Generator(nn.Module):
    def __init__(self, z_dim, num_bbox_labels, img_channels, img_height, img_width, c_dim,
                 f_dim=256, num_heads=4, num_layers=8,
                 hidden_dim=256,
                 med_config='configs/med_config.json', bert_f_dim=768, bert_num_encoder_layers=12, bert_num_decoder_layers=12, bert_num_heads=12,
                 background_size=1024, im_f_dim=512,
                 max_text_length=256, backbone='resnet50'):
        super().__init__()
        self.z_dim = z_dim
        self.num_bbox_labels = num_bbox_labels
        self.c_dim = c_dim
        self.max_text_length = max_text_length
        self.hidden_dim = hidden_dim
        # execution options (exact, parity-preserving): see DESIGN.md "text path"
        self.text_trim = False       # drop all-padding token columns before the BERT stacks
        self.text_dedup = False      # reuse the frozen encoder's CLS features across calls on the same strings
        self._cls_cache = None

        self.backbone = build_backbone(backbone, background_size, hidden_dim)
        self.input_proj = nn.Conv2d(self.backbone.num_channels, hidden_dim, kernel_size=1)

        self.fc_z = nn.Linear(z_dim * 9, bert_f_dim)
        self.emb_label = nn.Embedding(num_bbox_labels, bert_f_dim)

        self.tokenizer = init_tokenizer()
        encoder_config = _med_config(med_config)
        encoder_config.encoder_width = bert_f_dim
        encoder_config.num_hidden_layers = bert_num_encoder_layers
        encoder_config.num_attention_heads = bert_num_heads
        self.text_encoder = BertModel.from_pretrained('bert-base-uncased', config=encoder_config, add_pooling_layer=False)
        self.text_encoder.resize_token_embeddings(len(self.tokenizer))

        self.enc_text_len = nn.Embedding(max_text_length, bert_f_dim)
        self.fc_in = MLP(input_dim=bert_f_dim * 4, hidden_dim=bert_f_dim, output_dim=hidden_dim, num_layers=3)

        self.transformer = Transformer(d_model=hidden_dim, dropout=0.1, nhead=8, dim_feedforward=2048,
                                       num_encoder_layers=6, num_decoder_layers=6, normalize_before=False,
                                       return_intermediate_dec=False)
        self.bbox_embed = MLP(input_dim=hidden_dim, hidden_dim=hidden_dim, output_dim=4, num_layers=3)

        self.fc_z_rec = nn.Linear(hidden_dim, z_dim * 9)
        self.fc_out_cls = nn.Linear(hidden_dim, num_bbox_labels)

        decoder_config = _med_config(med_config)
        decoder_config.encoder_width = im_f_dim
        decoder_config.num_hidden_layers = bert_num_decoder_layers
        decoder_config.num_attention_heads = bert_num_heads
        self.text_decoder = BertLMHeadModel.from_pretrained('bert-base-uncased', config=decoder_config)
        self.text_decoder.resize_token_embeddings(len(self.tokenizer))

        self.fc_text_len_rec = nn.Linear(hidden_dim, max_text_length)
        self._text = None

    def _front(self):
        if self._text is None or self._text.tokenizer is not self.tokenizer:
            self._text = _TextFrontEnd(self.tokenizer, self.max_text_length)
        return self._text

    def forward(self, z, bbox_class, bbox_real, bbox_text, bbox_patch, padding_mask, background, c, reconst=False):
        if isinstance(background, (list, tuple)):
            background = torch.stack(list(background))
        dev = background.device
        if dev.type != "cuda":
            raise RuntimeError("layoutdetr_b200.Generator runs on CUDA (sm_100a) only; there is no CPU fallback")
        B, N = bbox_patch.shape[0], bbox_patch.shape[1]
        H = self.hidden_dim

        # lane scheduling (lanes.py, level 2): the text-decoder LM loss depends on the token ids only
        branches = reconst and LANES.active(2)
        text = self._front()(bbox_text, dev)
        s_td = loss_lm = None
        if reconst:
            valid, valid_cpu = valid_index(padding_mask)
        if branches:
            s_td = LANES.fork("td", cta_limit=LANES.lm_ctas, bulk=True)
            with torch.cuda.stream(s_td):
                loss_lm = _decode_text_loss(self, text, valid, valid_cpu, self.tokenizer.bos_token_id, self.tokenizer.pad_token_id)

        feat, pos, h, w = self.backbone(background.float())
        S = h * w
        src = Fn.conv2d(feat, self.input_proj.weight, None, self.input_proj.bias, None, B, h, w, 1, 0, K.ACT_NONE)

        z0 = normalize_2nd_moment(z.reshape(B, -1).float())
        zf = Fn.linear(Fn.to_bf16_padded(z0), self.fc_z.weight, self.fc_z.bias)                     # [B, 768]
        l = F.embedding(bbox_class, self.emb_label.weight)                                          # [B, N, 768]
        text_feat = _encode_text(self, text, B, N)                                                  # [B*N, 768] bf16
        text_len = text["text_len"].view(B, N)
        text_len_feat = F.embedding(text_len, self.enc_text_len.weight)
        x = torch.cat([zf.unsqueeze(1).expand(-1, N, -1), l.to(torch.bfloat16), text_feat.view(B, N, -1),
                       text_len_feat.to(torch.bfloat16)], dim=-1).reshape(B * N, -1)
        x = self.fc_in(x, final_act=K.ACT_RELU)                                                     # [B*N, 256]

        hs, _, _ = self.transformer(src, pos, x, padding_mask, B, S, N)                             # [B*N, 256]
        bbox_fake = self.bbox_embed(hs, out_f32=True).sigmoid().view(B, N, 4)
        if not reconst:
            return bbox_fake

        xv = hs.index_select(0, valid)                                                              # [M, 256]
        z_rec = Fn.linear_f32(xv, self.fc_z_rec.weight, self.fc_z_rec.bias)
        z_tgt = z0.unsqueeze(1).expand(-1, N, -1).reshape(B * N, -1).index_select(0, valid)
        loss_z = F.mse_loss(z_rec, z_tgt)
        logit_cls = Fn.linear_f32(xv, self.fc_out_cls.weight, self.fc_out_cls.bias)
        if s_td is None:
            loss_lm = _decode_text_loss(self, text, valid, valid_cpu, self.tokenizer.bos_token_id, self.tokenizer.pad_token_id)
        text_len_rec = Fn.linear_f32(xv, self.fc_text_len_rec.weight, self.fc_text_len_rec.bias)
        loss_text_len = Fn.cross_entropy(text_len_rec, text_len.reshape(-1).index_select(0, valid))
        if s_td is not None:
            LANES.join(s_td, loss_lm)
        return bbox_fake, loss_z, logit_cls, loss_lm, loss_text_len


class Discriminator(nn.Module):
    def __init__(self, num_bbox_labels, img_channels, img_height, img_width, c_dim,
                 f_dim=256, num_heads=4, num_layers=8, max_bbox=50,
                 hidden_dim=256,
                 med_config='configs/med_config.json', bert_f_dim=768, bert_num_encoder_layers=12, bert_num_decoder_layers=12, bert_num_heads=12,
                 background_size=1024, im_f_dim=512,
                 max_text_length=256, backbone='resnet50'):
        super().__init__()
        from .networks_stylegan2 import Decoder
        self.num_bbox_labels = num_bbox_labels
        self.c_dim = c_dim
        self.max_text_length = max_text_length
        self.hidden_dim = hidden_dim
        self.text_trim = False
        self.text_dedup = False
        self._cls_cache = None

        self.backbone = build_backbone(backbone, background_size, hidden_dim)
        self.input_proj = nn.Conv2d(self.backbone.num_channels, hidden_dim, kernel_size=1)

        self.fc_bbox = nn.Linear(4, bert_f_dim)
        self.emb_label = nn.Embedding(num_bbox_labels, bert_f_dim)

        self.tokenizer = init_tokenizer()
        encoder_config = _med_config(med_config)
        encoder_config.encoder_width = bert_f_dim
        encoder_config.num_hidden_layers = bert_num_encoder_layers
        encoder_config.num_attention_heads = bert_num_heads
        self.text_encoder = BertModel.from_pretrained('bert-base-uncased', config=encoder_config, add_pooling_layer=False)
        self.text_encoder.resize_token_embeddings(len(self.tokenizer))

        self.enc_text_len = nn.Embedding(max_text_length, bert_f_dim)
        self.enc_fc_in = MLP(input_dim=bert_f_dim * 4, hidden_dim=bert_f_dim, output_dim=hidden_dim, num_layers=3)
        self.enc_transformer = TransformerWithToken(d_model=hidden_dim, dropout=0.1, nhead=8, dim_feedforward=2048,
                                                    num_encoder_layers=6, num_decoder_layers=6, normalize_before=False,
                                                    return_intermediate_dec=False)
        self.fc_out_disc = nn.Linear(hidden_dim, 1)

        self.pos_token = nn.Parameter(torch.rand(max_bbox, 1, hidden_dim))
        self.dec_fc_in = nn.Linear(hidden_dim + hidden_dim, hidden_dim)
        te = nn.TransformerEncoderLayer(d_model=hidden_dim, nhead=8, dim_feedforward=2048)
        self.dec_transformer = nn.TransformerEncoder(te, num_layers=6, enable_nested_tensor=False)
        self.bbox_embed = nn.Linear(hidden_dim, 4)
        self.fc_out_cls = nn.Linear(hidden_dim, num_bbox_labels)

        decoder_config = _med_config(med_config)
        decoder_config.encoder_width = im_f_dim
        decoder_config.num_hidden_layers = bert_num_decoder_layers
        decoder_config.num_attention_heads = bert_num_heads
        self.text_decoder = BertLMHeadModel.from_pretrained('bert-base-uncased', config=decoder_config)
        self.text_decoder.resize_token_embeddings(len(self.tokenizer))

        self.fc_text_len_rec = nn.Linear(hidden_dim, max_text_length)
        self.bg_decoder = Decoder(z_dim=hidden_dim, w_dim=im_f_dim, channel_max=im_f_dim, channel_base=8192,
                                  img_channels=img_channels, img_resolution=background_size, use_noise=False,
                                  num_fp16_res=0, conv_clamp=None, fused_modconv_default=False)

        self.fc_bbox_uncond = nn.Linear(4, bert_f_dim)
        self.emb_label_uncond = nn.Embedding(num_bbox_labels, bert_f_dim)
        self.enc_fc_in_uncond = MLP(input_dim=bert_f_dim + bert_f_dim, hidden_dim=bert_f_dim, output_dim=hidden_dim, num_layers=3)
        self.enc_transformer_uncond = TransformerWithToken_layoutganpp(d_model=hidden_dim, dim_feedforward=2048, nhead=8, num_layers=6)
        self.fc_out_disc_uncond = nn.Linear(hidden_dim, 1)

        self.pos_token_uncond = nn.Parameter(torch.rand(max_bbox, 1, hidden_dim))
        self.dec_fc_in_uncond = nn.Linear(hidden_dim + hidden_dim, hidden_dim)
        te_uncond = nn.TransformerEncoderLayer(d_model=hidden_dim, nhead=8, dim_feedforward=2048)
        self.dec_transformer_uncond = nn.TransformerEncoder(te_uncond, num_layers=6, enable_nested_tensor=False)
        self.bbox_embed_uncond = nn.Linear(hidden_dim, 4)
        self.fc_out_cls_uncond = nn.Linear(hidden_dim, num_bbox_labels)
        self._text = None

    def _front(self):
        if self._text is None or self._text.tokenizer is not self.tokenizer:
            self._text = _TextFrontEnd(self.tokenizer, self.max_text_length)
        return self._text

    def _decode_branch(self, x0, pos_token, dec_fc_in, dec_transformer, B, N, padding_mask, valid):
        """x0 [B, H] bf16 -> per-slot features of the valid slots [M, H] (reference :315-321, :352-357)."""
        H = self.hidden_dim
        t = Fn.to_bf16_padded(pos_token[:N].reshape(N, H))                                          # [N, H]
        x = torch.cat([x0.unsqueeze(1).expand(-1, N, -1), t.unsqueeze(0).expand(B, -1, -1)], dim=-1).reshape(B * N, 2 * H)
        x = Fn.linear(x, dec_fc_in.weight, dec_fc_in.bias, act=K.ACT_RELU)
        x = TransformerEncoderStack.run(dec_transformer, x, B, N, padding_mask)
        return x.index_select(0, valid)

    def forward(self, bbox, bbox_class, bbox_text, bbox_patch, padding_mask, background, c, reconst=False):
        if isinstance(background, (list, tuple)):
            background = torch.stack(list(background))
        dev = background.device
        if dev.type != "cuda":
            raise RuntimeError("layoutdetr_b200.Discriminator runs on CUDA (sm_100a) only; there is no CPU fallback")
        B, N = bbox_patch.shape[0], bbox_patch.shape[1]
        H = self.hidden_dim

        # lane scheduling (lanes.py, level 2): branches that do not depend on the conditional DETR chain run beside it —
        # the LM text decoder (token ids only), the unconditional branch (boxes / classes only), the background decoder
        # (token output only).  Single-stream when lanes are off: same calls, same order as the reference (:279-361).
        branches = LANES.active(2)
        text = self._front()(bbox_text, dev)
        s_td = s_un = s_bg = loss_lm = bg_rec = None
        valid = valid_cpu = None
        if reconst:
            valid, valid_cpu = valid_index(padding_mask)
        if branches and reconst:
            s_td = LANES.fork("td", cta_limit=LANES.lm_ctas, bulk=True)
            with torch.cuda.stream(s_td):
                loss_lm = _decode_text_loss(self, text, valid, valid_cpu, self.tokenizer.bos_token_id, self.tokenizer.pad_token_id)

        bbox2d = Fn.to_bf16_padded(bbox.reshape(B * N, 4).float())

        def uncond():
            b_u = Fn.linear(bbox2d, self.fc_bbox_uncond.weight, self.fc_bbox_uncond.bias)
            l_u = F.embedding(bbox_class, self.emb_label_uncond.weight)
            x_u = torch.cat([b_u.view(B, N, -1), l_u.to(torch.bfloat16)], dim=-1).reshape(B * N, -1)
            x_u = self.enc_fc_in_uncond(x_u, final_act=K.ACT_RELU)
            x_u = self.enc_transformer_uncond(x_u, B, N, padding_mask)                              # [B*(N+1), 256]
            x0_u = x_u.view(B, N + 1, H)[:, 0, :]
            logit_u = Fn.linear_f32(x0_u, self.fc_out_disc_uncond.weight, self.fc_out_disc_uncond.bias).squeeze(-1)
            if not reconst:
                return logit_u, None, None
            xv_u = self._decode_branch(x0_u, self.pos_token_uncond, self.dec_fc_in_uncond, self.dec_transformer_uncond,
                                       B, N, padding_mask, valid)
            bbox_u = Fn.linear_f32(xv_u, self.bbox_embed_uncond.weight, self.bbox_embed_uncond.bias).sigmoid()
            cls_u = Fn.linear_f32(xv_u, self.fc_out_cls_uncond.weight, self.fc_out_cls_uncond.bias)
            return logit_u, bbox_u, cls_u

        un_out = None
        if branches:
            s_un = LANES.fork("un", bbox2d, bbox_class, padding_mask)
            with torch.cuda.stream(s_un):
                un_out = uncond()

        feat, pos, h, w = self.backbone(background.float())
        S = h * w
        src = Fn.conv2d(feat, self.input_proj.weight, None, self.input_proj.bias, None, B, h, w, 1, 0, K.ACT_NONE)

        b = Fn.linear(bbox2d, self.fc_bbox.weight, self.fc_bbox.bias)                               # [B*N, 768]
        l = F.embedding(bbox_class, self.emb_label.weight)
        text_feat = _encode_text(self, text, B, N)
        text_len = text["text_len"].view(B, N)
        text_len_feat = F.embedding(text_len, self.enc_text_len.weight)
        x = torch.cat([b.view(B, N, -1), l.to(torch.bfloat16), text_feat.view(B, N, -1),
                       text_len_feat.to(torch.bfloat16)], dim=-1).reshape(B * N, -1)
        x = self.enc_fc_in(x, final_act=K.ACT_RELU)

        hs, _, L = self.enc_transformer(src, pos, x, padding_mask, B, S, N)                         # [B*(N+1), 256]
        x0 = hs.view(B, L, H)[:, 0, :]                                                              # token output
        if branches and reconst:
            s_bg = LANES.fork("bg", x0)
            with torch.cuda.stream(s_bg):
                bg_rec = self.bg_decoder(x0)
        logit_disc = Fn.linear_f32(x0, self.fc_out_disc.weight, self.fc_out_disc.bias).squeeze(-1)

        if not branches:
            un_out = uncond()
        if not reconst:
            if s_un is not None:
                LANES.join(s_un, un_out[0])
            return logit_disc, un_out[0]

        xv = self._decode_branch(x0, self.pos_token, self.dec_fc_in, self.dec_transformer, B, N, padding_mask, valid)
        bbox_pred = Fn.linear_f32(xv, self.bbox_embed.weight, self.bbox_embed.bias).sigmoid()
        logit_cls = Fn.linear_f32(xv, self.fc_out_cls.weight, self.fc_out_cls.bias)
        if s_td is None:
            loss_lm = _decode_text_loss(self, text, valid, valid_cpu, self.tokenizer.bos_token_id, self.tokenizer.pad_token_id)
        text_len_rec = Fn.linear_f32(xv, self.fc_text_len_rec.weight, self.fc_text_len_rec.bias)
        loss_text_len = Fn.cross_entropy(text_len_rec, text_len.reshape(-1).index_select(0, valid))
        if s_bg is None:
            bg_rec = self.bg_decoder(x0)
        logit_disc_uncond, bbox_pred_uncond, logit_cls_uncond = un_out
        if branches:
            LANES.join(s_un, logit_disc_uncond, bbox_pred_uncond, logit_cls_uncond)
            LANES.join(s_td, loss_lm)
            LANES.join(s_bg, bg_rec)
        return (logit_disc, logit_disc_uncond, bbox_pred, logit_cls, loss_lm, loss_text_len, bg_rec,
                bbox_pred_uncond, logit_cls_uncond)
