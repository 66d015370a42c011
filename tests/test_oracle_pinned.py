"""CPU: pin the oracle restatement (oracle/layoutdetr_oracle.py) against outputs of the REAL reference
(tests/golden/*.pt, produced by tests/golden/gen_golden.py where /root/reference exists)."""
import os

import pytest
import torch

from helpers import build, golden, state_dict_f32, GOLD


def _tok():
    from layoutdetr_b200.synthetic import SyntheticTokenizer
    return SyntheticTokenizer()


def test_state_dict_tree_matches_reference_manifest():
    import json
    with open(os.path.join(GOLD, "state_dict_manifest.json")) as f:
        man = json.load(f)
    for which in ("G", "D"):
        ours = {k: list(v.shape) for k, v in build(which).state_dict().items()}
        assert ours == man[which], "state_dict of %s differs from the reference tree" % which


def test_oracle_ops_match_reference_ref_impls():
    from oracle import layoutdetr_oracle as O
    ops = golden("ops_ref.pt")
    for c in ops["bias_act"]:
        y = O.bias_act(c["x"], c["b"], dim=c["dim"], act=c["act"], gain=c["gain"], clamp=c["clamp"])
        torch.testing.assert_close(y, c["y"], atol=1e-6, rtol=1e-6)
    for c in ops["upfirdn2d"]:
        y = O.upfirdn2d(c["x"], c["f"], up=c["up"], down=c["down"], padding=tuple(c["padding"]), flip_filter=c["flip_filter"], gain=c["gain"])
        torch.testing.assert_close(y, c["y"], atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("name", ["model_b1_v4"])
def test_oracle_generator_matches_reference(name):
    from layoutdetr_b200.synthetic import make_inputs
    from oracle import layoutdetr_oracle as O
    g = golden(name + ".pt")
    inp = make_inputs(g["batch"], n_valid=g["n_valid"], seed=g["inputs_seed"])
    sd = state_dict_f32(build("G"))
    with torch.no_grad():
        out = O.generator_forward(sd, _tok(), inp["z"], inp["bbox_class"], inp["bbox_text"], inp["padding_mask"],
                                  inp["background"], reconst=True)
    for key, val in zip(["bbox_fake", "loss_z", "logit_cls", "loss_lm", "loss_text_len"], out[:5]):
        torch.testing.assert_close(val, g["G"][key], atol=2e-4, rtol=2e-4, msg=lambda m: "%s: %s" % (key, m))
    torch.testing.assert_close(out[5]["text_cls"], g["G_inter"]["text_cls"], atol=2e-4, rtol=2e-4)
    torch.testing.assert_close(out[5]["hs"], g["G_inter"]["hs"], atol=5e-4, rtol=5e-4)


@pytest.mark.parametrize("name", ["model_b1_v4"])
def test_oracle_discriminator_matches_reference(name):
    from layoutdetr_b200.synthetic import make_inputs
    from oracle import layoutdetr_oracle as O
    g = golden(name + ".pt")
    inp = make_inputs(g["batch"], n_valid=g["n_valid"], seed=g["inputs_seed"])
    sd = state_dict_f32(build("D"))
    names = ["logit_disc", "logit_disc_uncond", "bbox_pred", "logit_cls", "loss_lm", "loss_text_len", "bg_rec",
             "bbox_pred_uncond", "logit_cls_uncond"]
    with torch.no_grad():
        out = O.discriminator_forward(sd, _tok(), inp["bbox_real"], inp["bbox_class"], inp["bbox_text"], inp["padding_mask"],
                                      inp["background"], reconst=True)
    for key, val in zip(names, out):
        if key == "bg_rec":
            torch.testing.assert_close(val[:, :, ::8, ::8], g["D"]["bg_rec_sub"], atol=1e-3, rtol=1e-3)
            assert abs(float(val.mean()) - g["D"]["bg_rec_mean"]) < 1e-3 * max(1.0, abs(g["D"]["bg_rec_std"]))
        else:
            torch.testing.assert_close(val, g["D"][key], atol=3e-4, rtol=3e-4, msg=lambda m: "%s: %s" % (key, m))


def test_oracle_matches_reference_on_ragged_batch():
    """Ragged batch (1, 5 and 9 real elements of 9 slots — the empty-ish, typical and maximum cases in one batch): G and D of
    the unmodified reference (model_b3_ragged.pt) vs the oracle; the product is checked against the oracle on the same inputs
    on the GPU (tests/test_zz_ragged_gpu.py)."""
    from layoutdetr_b200.synthetic import make_ragged_inputs
    from oracle import layoutdetr_oracle as O
    g = golden("model_b3_ragged.pt")
    inp = make_ragged_inputs(g["ragged"], seed=g["inputs_seed"])
    keep = ~inp["padding_mask"]
    with torch.no_grad():
        og = O.generator_forward(state_dict_f32(build("G")), _tok(), inp["z"], inp["bbox_class"], inp["bbox_text"], inp["padding_mask"],
                                 inp["background"], reconst=True)
        od = O.discriminator_forward(state_dict_f32(build("D")), _tok(), inp["bbox_real"], inp["bbox_class"], inp["bbox_text"],
                                     inp["padding_mask"], inp["background"], reconst=True)
    for key, val in zip(["bbox_fake", "loss_z", "logit_cls", "loss_lm", "loss_text_len"], og[:5]):
        ref = g["G"][key]
        if key == "bbox_fake":
            val, ref = val[keep], ref[keep]
        torch.testing.assert_close(val, ref, atol=3e-4, rtol=3e-4, msg=lambda m: "%s: %s" % (key, m))
    names = ["logit_disc", "logit_disc_uncond", "bbox_pred", "logit_cls", "loss_lm", "loss_text_len", "bg_rec", "bbox_pred_uncond", "logit_cls_uncond"]
    for key, val in zip(names, od):
        if key == "bg_rec":
            torch.testing.assert_close(val[:, :, ::8, ::8], g["D"]["bg_rec_sub"], atol=1e-3, rtol=1e-3)
        else:
            torch.testing.assert_close(val, g["D"][key], atol=3e-4, rtol=3e-4, msg=lambda m: "%s: %s" % (key, m))


def test_oracle_generator_matches_reference_at_1024_background():
    """BASELINE configs[3] geometry (1024 x 1024 background -> 1024 image tokens): G.forward of the reference vs the oracle."""
    from layoutdetr_b200.synthetic import make_inputs
    from oracle import layoutdetr_oracle as O
    g = golden("model_b1_bg1024.pt")
    inp = make_inputs(g["batch"], n_valid=g["n_valid"], seed=g["inputs_seed"], background_size=g["background_size"])
    with torch.no_grad():
        out = O.generator_forward(state_dict_f32(build("G")), _tok(), inp["z"], inp["bbox_class"], inp["bbox_text"], inp["padding_mask"],
                                  inp["background"])
    keep = ~inp["padding_mask"]
    torch.testing.assert_close(out[keep], g["bbox_fake"][keep], atol=3e-4, rtol=3e-4)


# ------------------------------------------------------------------------------------------------
# evaluation sweep: goldens in eval_ref.pt come from the reference's LayoutNet, metric functions,
# FeatureStats and layout-FID formula (tests/golden/gen_golden.py gen_eval)
# ------------------------------------------------------------------------------------------------
def _layoutnet_sd():
    from layoutdetr_b200.synthetic import synth_state_dict
    from layoutdetr_b200.training.networks_layoutnet import LayoutNet
    net = LayoutNet(13)
    return net, synth_state_dict(net)


def test_layoutnet_state_dict_tree_matches_reference():
    g = golden("eval_ref.pt")
    net, _ = _layoutnet_sd()
    assert {k: list(v.shape) for k, v in net.state_dict().items()} == g["layoutnet_keys"]


def test_oracle_eval_sweep_matches_reference():
    from oracle import layoutdetr_oracle as O
    g = golden("eval_ref.pt")
    _, sd = _layoutnet_sd()
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}
    real, fake, label, mask = g["bbox_real"], g["bbox_fake"], g["label"], g["mask"]
    with torch.no_grad():
        f_real = O.layoutnet_extract_features(sd, real, label, ~mask)
        f_fake = O.layoutnet_extract_features(sd, fake, label, ~mask)
    torch.testing.assert_close(f_real[:32], g["f_real"], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(f_fake[:32], g["f_fake"], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(O.compute_overlap(fake, mask), g["overlap"], atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(O.compute_alignment(fake, mask), g["alignment"], atol=1e-6, rtol=1e-5)
    iou, doc = O.layoutwise_iou_docsim(real, fake, mask)
    torch.testing.assert_close(iou.double(), g["iou"], atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(doc.double(), g["docsim"], atol=1e-6, rtol=1e-5)
    mu_r, s_r = O.feature_mean_cov(f_real.numpy())
    mu_f, s_f = O.feature_mean_cov(f_fake.numpy())
    fid = O.layout_fid(mu_f, s_f, mu_r, s_r)
    assert abs(fid - g["fid"]) <= 1e-4 * abs(g["fid"]) + 1e-7, (fid, g["fid"])


def test_pair_metric_arithmetic_matches_reference_goldens():
    """The arithmetic ld_layout_pair_metrics compiles (host build of csrc/box_loss_math.h) on the reference's own outputs."""
    import ctypes
    from test_host_logic import _host_box_lib
    g = golden("eval_ref.pt")
    lib = _host_box_lib()
    real, fake = g["bbox_real"].contiguous(), g["bbox_fake"].contiguous()
    v8 = g["mask"].to(torch.uint8).contiguous()
    B, N = v8.shape
    iou, doc = torch.empty(B), torch.empty(B)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    lib.host_layout_pair_metrics(P(real), P(fake), P(v8), ctypes.c_long(B), N, P(iou), P(doc))
    torch.testing.assert_close(iou.double(), g["iou"], atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(doc.double(), g["docsim"], atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("phase", ["Gmain", "Dmain"])
def test_oracle_train_step_gradients_match_reference(phase):
    """oracle/train_step.py (the CPU baseline bench.py times, and the checker of the loss tests) against per-parameter gradients of
    the reference's own accumulate_gradients (tests/golden/loss_b2_v8.pt, fp32 CPU, dropout off)."""
    from helpers import build, state_dict_f32
    from layoutdetr_b200.synthetic import SyntheticTokenizer, make_inputs
    from oracle import train_step
    g = golden("loss_b2_v8.pt")
    sdG, sdD = state_dict_f32(build("G")), state_dict_f32(build("D"))
    inp = make_inputs(g["batch"], n_valid=g["n_valid"], seed=g["inputs_seed"])
    grads = train_step.phase_gradients(sdG, sdD, SyntheticTokenizer(), inp, phase)
    ref = g["grads"][phase]
    worst = (0.0, "")
    n = 0
    for k, n_ref in ref["norms"].items():
        if n_ref < 1e-7:
            continue
        assert k in grads, "oracle produced no gradient for %s" % k
        e = abs(float(grads[k].norm()) - n_ref) / n_ref
        worst = max(worst, (e, k))
        n += 1
    for k, t in ref["small"].items():
        if float(t.norm()) < 1e-7:
            continue
        e = float((grads[k] - t).norm() / t.norm())
        worst = max(worst, (e, k + " (full tensor)"))
    print(phase, "parameters compared:", n, "worst relative error: %.3e (%s)" % worst)
    assert n > 250 and worst[0] < 2e-3, worst
