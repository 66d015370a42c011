"""Shared test helpers: build the product modules with the deterministic synthetic weights."""
import functools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

G_KWARGS = dict(z_dim=4, num_bbox_labels=8, img_channels=3, img_height=1024, img_width=1024, c_dim=0,
                background_size=256, bert_f_dim=768, bert_num_heads=4, bert_num_encoder_layers=12,
                bert_num_decoder_layers=2, im_f_dim=512)
D_KWARGS = dict(num_bbox_labels=8, img_channels=3, img_height=1024, img_width=1024, c_dim=0,
                background_size=256, bert_f_dim=768, bert_num_heads=4, bert_num_encoder_layers=12,
                bert_num_decoder_layers=2, im_f_dim=512)


def golden(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=False)


@functools.lru_cache(maxsize=None)
def build(which):
    """Product Generator / Discriminator (CPU) with synth weights.  Cached per process."""
    os.environ["LAYOUTDETR_SYNTHETIC_TOKENIZER"] = "1"
    from layoutdetr_b200.synthetic import synth_state_dict
    from layoutdetr_b200.training import networks_detr as nd
    m = nd.Generator(**G_KWARGS) if which == "G" else nd.Discriminator(**D_KWARGS)
    m.eval()
    synth_state_dict(m)
    return m


def state_dict_f32(m):
    return {k: v.detach().cpu() for k, v in m.state_dict().items()}


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))
