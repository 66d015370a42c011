"""ViT-B/16 image encoder (reference training/networks_vit.py:139-225): the oracle restatement is pinned to
outputs of the reference class (tests/golden/vit_ref.pt, gen_golden.py --only-vit) on CPU; the product module is checked against
both on the GPU."""
import pytest
import torch

from helpers import golden


def _inputs(entry):
    h, w = entry["size"]
    g = torch.Generator().manual_seed(entry["seed"])
    x = torch.randn((2, 3, h, w), generator=g)
    mask = torch.ones((2, 1, h, w))
    mask[1, :, :16, 16:48] = 0
    return x, mask


def _module(h, w):
    from layoutdetr_b200.synthetic import synth_state_dict
    from layoutdetr_b200.training.networks_vit import VisionTransformer
    torch.manual_seed(0)
    m = VisionTransformer(img_height=h, img_width=w).eval()
    synth_state_dict(m)
    return m


def test_vit_oracle_matches_reference_golden():
    from oracle import layoutdetr_oracle as O
    g = golden("vit_ref.pt")["small"]
    x, mask = _inputs(g)
    m = _module(*g["size"])
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    with torch.no_grad():
        torch.testing.assert_close(O.vit_forward(sd, x, mask), g["y"], atol=2e-4, rtol=1e-4)
        torch.testing.assert_close(O.vit_forward(sd, x, torch.ones_like(mask)), g["y_nomask"], atol=2e-4, rtol=1e-4)
    assert not torch.allclose(g["y"][1], g["y_nomask"][1], atol=1e-3)          # the mask matters for the masked sample ...
    torch.testing.assert_close(g["y"][0], g["y_nomask"][0], atol=1e-5, rtol=0)  # ... and only for it


def test_vit_state_dict_keys_match_reference_layout():
    m = _module(64, 96)
    keys = set(m.state_dict())
    for k in ("cls_token", "pos_embed", "token_mask", "patch_embed.proj.weight", "transformer.layers.11.self_attn.in_proj_weight",
              "transformer.layers.0.linear1.weight", "transformer.norm.weight", "norm.bias"):
        assert k in keys, k
    assert m.pos_embed.shape == (1, 4 * 6 + 1, 768)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small", "bg256"])
def test_vit_forward_matches_reference_golden(name):
    g = golden("vit_ref.pt")[name]
    x, mask = _inputs(g)
    m = _module(*g["size"]).cuda()
    with torch.no_grad():
        y = m(x.cuda(), mask.cuda()).float().cpu()
        y0 = m(x.cuda(), None).float().cpu()
    sc = float(g["y"].abs().max())
    assert float((y - g["y"]).abs().max()) < 3e-2 * sc, float((y - g["y"]).abs().max()) / sc          # bf16 through 12 layers
    assert float((y0 - g["y_nomask"]).abs().max()) < 3e-2 * sc


@pytest.mark.gpu
def test_generator_with_vit_backbone_1024_and_12_slots():
    """BASELINE configs[3]: 1024^2 background through the ViT-B/16 backbone (4096 image tokens), 12 element slots with the
    latent kept [B, 9, 4]: forward + backward run and give finite boxes / gradients."""
    import os
    os.environ["LAYOUTDETR_SYNTHETIC_TOKENIZER"] = "1"
    from helpers import G_KWARGS
    from layoutdetr_b200.synthetic import make_inputs, synth_state_dict
    from layoutdetr_b200.training import networks_detr as nd
    kw = dict(G_KWARGS, background_size=1024, backbone="vit_b16")
    torch.manual_seed(0)
    G = nd.Generator(**kw).eval()
    synth_state_dict(G)
    G = G.cuda()
    inp = make_inputs(1, n_valid=10, n_slots=12, background_size=1024, seed=3)
    inp["z"] = torch.randn((1, 9, 4), generator=torch.Generator().manual_seed(1))
    inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()}
    with torch.no_grad():
        box = G(inp["z"], inp["bbox_class"], inp["bbox_real"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"], inp["background"], inp["c"])
    assert box.shape == (1, 12, 4) and torch.isfinite(box).all() and float(box.min()) >= 0 and float(box.max()) <= 1
    G.requires_grad_(True)
    G.text_encoder.requires_grad_(False)
    out = G(inp["z"], inp["bbox_class"], inp["bbox_real"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"], inp["background"], inp["c"], True)
    (out[0].square().mean() + out[1] + out[3] + out[4]).backward()
    gw = G.backbone.body.transformer.layers[0].linear1.weight.grad
    assert gw is not None and torch.isfinite(gw).all() and float(gw.abs().max()) > 0
