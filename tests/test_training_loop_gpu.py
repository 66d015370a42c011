"""GPU: the drop-in `training_loop` (reference training/training_loop.py:63-467 signature) on the product modules, and the
contract the REFERENCE loop puts on G / D (:274-328, :398-411) exercised directly: zero_grad(set_to_none) + flat-gradient
write-back + torch.optim.Adam, deepcopy, pickle round trip, forward hooks of print_module_summary, reference-keyed state dicts."""
import copy
import io
import json
import os
import pickle

import pytest
import torch

from helpers import D_KWARGS, G_KWARGS, GOLD, build

pytestmark = pytest.mark.gpu

_COMMON = ("num_bbox_labels", "img_channels", "img_height", "img_width", "c_dim", "background_size")


def _net_kwargs(kw, cls):
    return dict({k: v for k, v in kw.items() if k not in _COMMON}, class_name="layoutdetr_b200.training.networks_detr." + cls)


def _loop_kwargs(tmp_path, **over):
    ds = dict(class_name="layoutdetr_b200.training.synthetic_dataset.SyntheticLayoutDataset", num_items=64, n_valid=8)
    kw = dict(run_dir=str(tmp_path), training_set_kwargs=ds, validation_set_kwargs=ds, data_loader_kwargs=dict(num_workers=0),
              G_kwargs=_net_kwargs(G_KWARGS, "Generator"), D_kwargs=_net_kwargs(D_KWARGS, "Discriminator"),
              G_opt_kwargs=dict(class_name="torch.optim.Adam", lr=1e-5, betas=[0, 0.99], eps=1e-8),
              D_opt_kwargs=dict(class_name="torch.optim.Adam", lr=1e-5, betas=[0, 0.99], eps=1e-8),
              loss_kwargs={}, metrics=[], random_seed=0, num_gpus=1, rank=0, batch_size=2, batch_gpu=2, G_reg_interval=4,
              D_reg_interval=16, total_kimg=1, kimg_per_tick=1, image_snapshot_ticks=None, network_snapshot_ticks=1)
    kw.update(over)
    return kw


def test_training_loop_runs_snapshots_and_logs(tmp_path):
    from layoutdetr_b200.training.training_loop import training_loop
    res = training_loop(**_loop_kwargs(tmp_path, max_iterations=4))
    assert res["iterations"] == 4 and res["graphed"] is not None and len(res["graphed"].graphs) == 1
    assert res["graphed"].eager_steps == 0
    lines = [json.loads(l) for l in open(tmp_path / "stats.jsonl")]
    assert len(lines) >= 1 and "Loss/G/loss_Ggen_bbox_rec" in lines[-1] and "Timing/iteration_ms" in lines[-1]
    pkls = sorted(p for p in os.listdir(tmp_path) if p.startswith("network-snapshot-"))
    assert pkls, os.listdir(tmp_path)
    with open(tmp_path / pkls[-1], "rb") as f:
        snap = pickle.load(f)
    assert set(snap) == {"G", "D", "G_ema", "augment_pipe", "training_set_kwargs"}
    tr = res["trainer"]
    g_now = {k: v.detach().cpu() for k, v in tr.G.state_dict().items()}
    moved = sum(1 for k, v in snap["G"].state_dict().items() if v.is_floating_point() and not torch.equal(v, g_now[k]))
    assert moved == 0                                            # the last snapshot is the final weights
    G0 = build("G")                                              # training changed the trainable weights, not the frozen text encoder
    changed = [k for k, v in g_now.items() if v.is_floating_point() and k.startswith("transformer.") and "weight" in k]
    assert changed
    assert not snap["G"].training and not snap["G_ema"].training
    # G_ema trails G: different from G after a few steps of Adam, but finite
    ema = snap["G_ema"].state_dict()
    k = "transformer.decoder.layers.0.linear1.weight"
    assert torch.isfinite(ema[k]).all() and not torch.equal(ema[k], g_now[k])


def test_training_loop_micro_batches_eager(tmp_path):
    """batch_gpu < batch_size // num_gpus: gradients of two micro-batches accumulate before one optimizer step."""
    from layoutdetr_b200.training.training_loop import training_loop
    res = training_loop(**_loop_kwargs(tmp_path, batch_size=4, batch_gpu=2, max_iterations=2, network_snapshot_ticks=None))
    assert res["iterations"] == 2 and res["graphed"] is None
    assert all(torch.isfinite(v.float()).all() for ph in res["trainer"].loss.last.values() for v in ph.values())


def test_reference_loop_contract_on_product_modules():
    """What the reference's hot loop does to a module (training_loop.py:281-313), step by step, on the product Generator."""
    from layoutdetr_b200.synthetic import make_inputs
    from layoutdetr_b200.torch_utils import misc
    from layoutdetr_b200.training.loss import StyleGAN2Loss
    dev = torch.device("cuda")
    G = copy.deepcopy(build("G")).to(dev).train().requires_grad_(False)
    D = copy.deepcopy(build("D")).to(dev).train().requires_grad_(False)
    # reference-keyed checkpoint: every key / shape of the real reference's state_dict (manifest written by gen_golden.py)
    manifest = json.load(open(os.path.join(GOLD, "state_dict_manifest.json")))
    ref_sd = {k: torch.zeros(shape) for k, shape in manifest["G"].items()}
    missing, unexpected = G.load_state_dict({**G.state_dict(), **{k: v for k, v in ref_sd.items() if k not in G.state_dict()}}, strict=False)
    assert not [k for k in unexpected] and set(manifest["G"]) == set(G.state_dict())
    inp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in make_inputs(2, n_valid=8, seed=3).items()}
    opt = torch.optim.Adam(G.parameters(), lr=1e-5, betas=(0.0, 0.99), eps=1e-8)
    loss = StyleGAN2Loss(device=dev, G=G, D=D)
    before = {k: v.detach().clone() for k, v in G.named_parameters()}
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        G.requires_grad_(True)
        G.text_encoder.requires_grad_(False)
        loss.accumulate_gradients(phase="Gmain", bbox_real=inp["bbox_real"], bbox_class=inp["bbox_class"], bbox_text=inp["bbox_text"],
                                  bbox_patch=inp["bbox_patch"], padding_mask=inp["padding_mask"], background=inp["background"],
                                  real_c=inp["c"], gen_z=inp["z"], gen_c=inp["c"], gain=1, cur_nimg=0)
        G.requires_grad_(False)
        params = [p for p in G.parameters() if p.grad is not None]
        assert len(params) > 250
        flat = torch.cat([p.grad.flatten() for p in params])
        flat /= 1
        torch.nan_to_num(flat, nan=0, posinf=1e5, neginf=-1e5, out=flat)
        for p, g in zip(params, flat.split([p.numel() for p in params])):
            p.grad = g.reshape(p.shape)
        opt.step()
    torch.cuda.synchronize()
    changed = sum(1 for k, p in G.named_parameters() if not torch.equal(p, before[k]))
    frozen_same = all(torch.equal(p, before[k]) for k, p in G.named_parameters() if k.startswith("text_encoder."))
    assert changed > 250 and frozen_same
    # the bf16 shadows followed the optimizer: a second forward sees the new weights
    with torch.no_grad():
        G.eval()
        a = G(inp["z"], inp["bbox_class"], inp["bbox_real"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"], inp["background"], inp["c"])
        G2 = copy.deepcopy(G)                                    # deepcopy (:135, :401) and pickle (:411) round trips give the same function
        b = G2(inp["z"], inp["bbox_class"], inp["bbox_real"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"], inp["background"], inp["c"])
        buf = io.BytesIO()
        pickle.dump(copy.deepcopy(G).cpu(), buf)
        G3 = pickle.loads(buf.getvalue()).to(dev)
        c = G3(inp["z"], inp["bbox_class"], inp["bbox_real"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"], inp["background"], inp["c"])
    assert torch.equal(a, b) and torch.equal(a, c)
    # print_module_summary: forward hooks on every sub-module (torch_utils/misc.py:199-217)
    out = misc.print_module_summary(G, [inp["z"], inp["bbox_class"], inp["bbox_real"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"],
                                        inp["background"], inp["c"], True], file=io.StringIO())
    assert len(out) == 5 and torch.equal(out[0], a)


def test_managed_shadow_follows_load_state_dict():
    """a torch-level in-place write to a flat-stored parameter (load_state_dict after Trainer construction) must
    reach the bf16 tensor-core shadow."""
    from layoutdetr_b200.synthetic import make_inputs
    from layoutdetr_b200.training.trainer import Trainer
    dev = torch.device("cuda")
    G = copy.deepcopy(build("G")).to(dev)
    D = copy.deepcopy(build("D")).to(dev)
    tr = Trainer(G, D, dev, batch_size=2)
    inp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in make_inputs(2, n_valid=8, seed=4).items()}
    run = lambda: G(inp["z"], inp["bbox_class"], inp["bbox_real"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"], inp["background"], inp["c"])
    with torch.no_grad():
        a = run().clone()
        sd = {k: v.clone() for k, v in G.state_dict().items()}
        sd["bbox_embed.layers.2.bias"] = sd["bbox_embed.layers.2.bias"] + 1.0
        sd["transformer.decoder.layers.5.linear2.weight"] = sd["transformer.decoder.layers.5.linear2.weight"] * 0.5
        G.load_state_dict(sd)
        b = run().clone()
        G2 = copy.deepcopy(build("G")).to(dev)
        G2.load_state_dict(sd)
        c = G2(inp["z"], inp["bbox_class"], inp["bbox_real"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"], inp["background"], inp["c"])
    assert not torch.equal(a, b)
    torch.testing.assert_close(b, c, atol=1e-6, rtol=0)
