"""Ragged batches — a different number of real elements per layout, as every real dataset batch has — through the product
Generator / Discriminator against the CPU oracle (the oracle itself is pinned to the reference on a ragged golden in
tests/test_oracle_pinned.py::test_oracle_matches_reference_on_ragged_batch)."""
import pytest
import torch

from helpers import build, state_dict_f32

pytestmark = pytest.mark.gpu
BOX_TOL = 1e-2


def _dev(inp):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()}


def test_generator_and_discriminator_on_ragged_batch_match_oracle():
    from layoutdetr_b200.synthetic import make_ragged_inputs, SyntheticTokenizer
    from oracle import layoutdetr_oracle as O
    inp = make_ragged_inputs([1, 5, 9], seed=13)
    tok = SyntheticTokenizer()
    G, D = build("G").cuda(), build("D").cuda()
    d = _dev(inp)
    with torch.no_grad():
        rg = O.generator_forward(state_dict_f32(G), tok, inp["z"], inp["bbox_class"], inp["bbox_text"], inp["padding_mask"], inp["background"], reconst=True)
        rd = O.discriminator_forward(state_dict_f32(D), tok, inp["bbox_real"], inp["bbox_class"], inp["bbox_text"], inp["padding_mask"], inp["background"], reconst=True)
        og = G(d["z"], d["bbox_class"], d["bbox_real"], d["bbox_text"], d["bbox_patch"], d["padding_mask"], d["background"], d["c"], reconst=True)
        od = D(d["bbox_real"], d["bbox_class"], d["bbox_text"], d["bbox_patch"], d["padding_mask"], d["background"], d["c"], reconst=True)
    keep = ~inp["padding_mask"]
    bbox_fake, loss_z, logit_cls, loss_lm, loss_text_len = [t.float().cpu() for t in og]
    assert float((bbox_fake - rg[0])[keep].abs().max()) < BOX_TOL                        # padded slots carry no contract
    assert logit_cls.shape == rg[2].shape == (15, 8) and float((logit_cls - rg[2]).abs().max()) < 5e-2 * max(1.0, float(rg[2].abs().max()))
    assert abs(float(loss_lm) - float(rg[3])) < 2e-2 * float(rg[3])
    assert abs(float(loss_z) - float(rg[1])) < 3e-2 * max(1e-3, float(rg[1]))
    assert abs(float(loss_text_len) - float(rg[4])) < 3e-2 * float(rg[4])
    names = ["logit_disc", "logit_disc_uncond", "bbox_pred", "logit_cls", "loss_lm", "loss_text_len", "bg_rec", "bbox_pred_uncond", "logit_cls_uncond"]
    o = dict(zip(names, [t.float().cpu() for t in od]))
    r = dict(zip(names, rd))
    assert o["bbox_pred"].shape == (15, 4) and float((o["bbox_pred"] - r["bbox_pred"]).abs().max()) < BOX_TOL
    assert float((o["bbox_pred_uncond"] - r["bbox_pred_uncond"]).abs().max()) < BOX_TOL
    for k in ("logit_disc", "logit_disc_uncond", "logit_cls", "logit_cls_uncond"):
        assert float((o[k] - r[k]).abs().max()) < 5e-2 * max(1.0, float(r[k].abs().max())), k
    assert abs(float(o["loss_lm"]) - float(r["loss_lm"])) < 2e-2 * float(r["loss_lm"])


@pytest.mark.xfail(strict=False, reason="1024-token path (keys > 256: GEMM + softmax kernel instead of the fused attention) was added to the "
                                        "suite after this round's GPU budget was spent; not yet run on a B200")
def test_generator_at_1024_background_matches_reference_golden():
    """BASELINE configs[3] geometry: 1024 x 1024 background, 32 x 32 image tokens (golden from the reference, oracle pinned on CPU)."""
    from helpers import golden
    from layoutdetr_b200.synthetic import make_inputs
    g = golden("model_b1_bg1024.pt")
    inp = make_inputs(g["batch"], n_valid=g["n_valid"], seed=g["inputs_seed"], background_size=g["background_size"])
    G = build("G").cuda()
    d = _dev(inp)
    with torch.no_grad():
        out = G(d["z"], d["bbox_class"], d["bbox_real"], d["bbox_text"], d["bbox_patch"], d["padding_mask"], d["background"], d["c"]).float().cpu()
    keep = ~inp["padding_mask"]
    assert float((out - g["bbox_fake"])[keep].abs().max()) < BOX_TOL
