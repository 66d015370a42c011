"""GPU parity of the product Generator / Discriminator (hand-written sm_100a kernels, bf16 tensor cores)
against (a) golden outputs of the real reference and (b) the CPU oracle, on identical weights/inputs.
Tolerance: north_star's 1e-2 for bf16 on boxes / logits (absolute, boxes are in [0,1])."""
import pytest
import torch

from helpers import build, golden, state_dict_f32, rel_err

pytestmark = pytest.mark.gpu

BOX_TOL = 1e-2


def _to_dev(inp):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()}


@pytest.fixture(scope="module")
def G_cuda():
    G = build("G")
    return G.cuda()


@pytest.fixture(scope="module")
def D_cuda():
    D = build("D")
    return D.cuda()


@pytest.mark.parametrize("name", ["model_b1_v4", "model_b2_v8", "model_b16_v8"])
def test_generator_forward_matches_reference_golden(G_cuda, name):
    from layoutdetr_b200.synthetic import make_inputs
    g = golden(name + ".pt")
    inp = _to_dev(make_inputs(g["batch"], n_valid=g["n_valid"], seed=g["inputs_seed"]))
    with torch.no_grad():
        out = G_cuda(inp["z"], inp["bbox_class"], inp["bbox_real"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"],
                     inp["background"], inp["c"], reconst=True)
    torch.cuda.synchronize()
    bbox_fake, loss_z, logit_cls, loss_lm, loss_text_len = out
    ref = g["G"]
    report = {k: float((v.float().cpu() - ref[k]).abs().max()) for k, v in
              zip(["bbox_fake", "loss_z", "logit_cls", "loss_lm", "loss_text_len"], out)}
    print(name, report)
    assert report["bbox_fake"] < BOX_TOL, report
    assert report["logit_cls"] < 5e-2 * max(1.0, float(ref["logit_cls"].abs().max())), report
    assert abs(float(loss_lm) - float(ref["loss_lm"])) < 2e-2 * float(ref["loss_lm"]), report
    assert abs(float(loss_z) - float(ref["loss_z"])) < 3e-2 * max(1e-3, float(ref["loss_z"])), report
    assert abs(float(loss_text_len) - float(ref["loss_text_len"])) < 3e-2 * float(ref["loss_text_len"]), report


@pytest.mark.parametrize("name", ["model_b1_v4", "model_b2_v8", "model_b16_v8"])
def test_discriminator_forward_matches_reference_golden(D_cuda, name):
    from layoutdetr_b200.synthetic import make_inputs
    g = golden(name + ".pt")
    inp = _to_dev(make_inputs(g["batch"], n_valid=g["n_valid"], seed=g["inputs_seed"]))
    names = ["logit_disc", "logit_disc_uncond", "bbox_pred", "logit_cls", "loss_lm", "loss_text_len", "bg_rec",
             "bbox_pred_uncond", "logit_cls_uncond"]
    with torch.no_grad():
        out = D_cuda(inp["bbox_real"], inp["bbox_class"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"],
                     inp["background"], inp["c"], reconst=True)
    torch.cuda.synchronize()
    ref = g["D"]
    rep = {}
    for k, v in zip(names, out):
        v = v.float().cpu()
        if k == "bg_rec":
            st = v.shape[-1] // ref["bg_rec_sub"].shape[-1]          # goldens keep every 8th (bs16: every 16th) pixel
            diff = v[:, :, ::st, ::st] - ref["bg_rec_sub"].float()
            rep[k] = float(diff.abs().max()) / (ref["bg_rec_std"] + 1e-9)
            rep["bg_rec_rms"] = float(diff.square().mean().sqrt()) / (ref["bg_rec_std"] + 1e-9)
        else:
            rep[k] = float((v - ref[k]).abs().max())
    print(name, rep)
    assert rep["bbox_pred"] < BOX_TOL and rep["bbox_pred_uncond"] < BOX_TOL, rep
    scale = max(1.0, float(ref["logit_disc"].abs().max()))
    assert rep["logit_disc"] < 5e-2 * scale and rep["logit_disc_uncond"] < 5e-2 * max(1.0, float(ref["logit_disc_uncond"].abs().max())), rep
    assert rep["loss_lm"] < 2e-2 * float(ref["loss_lm"]), rep
    # background reconstruction through 13 bf16 conv layers, in units of the image std: rms over pixels and the single worst pixel
    assert rep["bg_rec_rms"] < 0.06 and rep["bg_rec"] < 0.3, rep


def test_generator_matches_cpu_oracle_other_seed(G_cuda):
    """Same check against the CPU oracle on inputs the goldens do not cover (B=1, 6 valid, seed 7)."""
    from layoutdetr_b200.synthetic import make_inputs, SyntheticTokenizer
    from oracle import layoutdetr_oracle as O
    inp = make_inputs(1, n_valid=6, seed=7)
    sd = state_dict_f32(G_cuda)
    with torch.no_grad():
        ref = O.generator_forward(sd, SyntheticTokenizer(), inp["z"], inp["bbox_class"], inp["bbox_text"], inp["padding_mask"],
                                  inp["background"], reconst=False)
        d = _to_dev(inp)
        out = G_cuda(d["z"], d["bbox_class"], d["bbox_real"], d["bbox_text"], d["bbox_patch"], d["padding_mask"], d["background"], d["c"])
    err = float((out.float().cpu() - ref).abs().max())
    print("oracle-vs-cuda bbox_fake max abs err", err)
    assert err < BOX_TOL
