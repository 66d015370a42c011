"""TEST INFRASTRUCTURE — import the UNMODIFIED reference (/root/reference) in the build container.

Only tests/golden/gen_golden.py and oracle-pinning tests use this, and only where /root/reference exists
(it does not exist on the GPU box).  Nothing is copied from the reference: the shims below only
patch import-time drift between the reference's pinned deps (transformers 4.19, timm, ...) and this
image, and replace network downloads with random init.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("LAYOUTDETR_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "training"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """Returns the reference's `training.networks_detr` module (plus side modules via sys.modules)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REF_ROOT)
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, REF_ROOT)
    os.chdir(REF_ROOT)  # configs/med_config.json is opened by relative path

    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    mu.prune_linear_layer = pu.prune_linear_layer
    mu.find_pruneable_heads_and_indices = None

    class PatchEmbed(nn.Module):
        def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
            super().__init__()
            hw = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
            self.proj = nn.Conv2d(in_chans, embed_dim, patch_size, patch_size)
            self.num_patches = (hw[0] // patch_size) * (hw[1] // patch_size)

        def forward(self, x):
            return self.proj(x).flatten(2).transpose(1, 2)

    _stub("timm"); _stub("timm.models")
    _stub("timm.models.vision_transformer", _cfg=lambda **k: {}, PatchEmbed=PatchEmbed)
    _stub("timm.models.registry", register_model=lambda f: f)
    _stub("timm.models.hub", download_cached_file=None)
    _stub("timm.models.layers", trunc_normal_=nn.init.trunc_normal_, DropPath=nn.Identity)
    _stub("timm.models.helpers", named_apply=None, adapt_input_conv=None)
    _stub("fairscale"); _stub("fairscale.nn"); _stub("fairscale.nn.checkpoint")
    _stub("fairscale.nn.checkpoint.checkpoint_activations", checkpoint_wrapper=lambda m: m)
    _stub("pytorch_fid"); _stub("pytorch_fid.fid_score", calculate_frechet_distance=None)
    _stub("seaborn"); _stub("skimage"); _stub("skimage.transform")
    _stub("selenium"); _stub("selenium.webdriver", Chrome=None)
    torch.hub.load_state_dict_from_url = lambda *a, **k: {}

    import training.med as med

    def _iw(self):
        self.apply(self._init_weights)
        if hasattr(self, "cls"):
            self.cls.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight

    med.BertPreTrainedModel.init_weights = _iw
    med.BertPreTrainedModel.get_head_mask = lambda self, hm, n, *a, **k: [None] * n
    med.BertPreTrainedModel.invert_attention_mask = lambda self, m: (1.0 - m[:, None, None, :].float()) * -10000.0
    med.BertModel.get_input_embeddings = lambda self: self.embeddings.word_embeddings
    med.BertLMHeadModel.get_input_embeddings = lambda self: self.bert.embeddings.word_embeddings

    def _resize(self, n):
        is_lm = hasattr(self, "cls")
        emb_owner = self.bert.embeddings if is_lm else self.embeddings
        old = emb_owner.word_embeddings
        new = nn.Embedding(n, old.embedding_dim, padding_idx=old.padding_idx)
        new.weight.data.normal_(0.0, 0.02)
        k = min(n, old.num_embeddings)
        new.weight.data[:k] = old.weight.data[:k]
        emb_owner.word_embeddings = new
        self.config.vocab_size = n
        if is_lm:
            pred = self.cls.predictions
            dec = nn.Linear(old.embedding_dim, n, bias=False)
            dec.weight = new.weight
            bias = nn.Parameter(torch.zeros(n))
            bias.data[:k] = pred.bias.data[:k]
            pred.bias = bias
            dec.bias = bias
            pred.decoder = dec
        return new

    med.BertPreTrainedModel.resize_token_embeddings = _resize
    for c in (med.BertModel, med.BertLMHeadModel):
        c.from_pretrained = classmethod(lambda k, name, config=None, **kw: k(config, **kw))

    if repo_root not in sys.path:
        sys.path.append(repo_root)
    from layoutdetr_b200.synthetic import SyntheticTokenizer
    import training.blip as blip
    blip.init_tokenizer = lambda: SyntheticTokenizer()
    import training.networks_detr as nd
    nd.init_tokenizer = blip.init_tokenizer
    _loaded = nd
    return nd


# constructor kwargs that train.py's defaults produce
G_KWARGS = dict(z_dim=4, num_bbox_labels=8, img_channels=3, img_height=1024, img_width=1024, c_dim=0,
                background_size=256, bert_f_dim=768, bert_num_heads=4, bert_num_encoder_layers=12,
                bert_num_decoder_layers=2, im_f_dim=512)
D_KWARGS = dict(num_bbox_labels=8, img_channels=3, img_height=1024, img_width=1024, c_dim=0,
                background_size=256, bert_f_dim=768, bert_num_heads=4, bert_num_encoder_layers=12,
                bert_num_decoder_layers=2, im_f_dim=512)
