"""GPU: the CUDA-graph replay of a training iteration must update the weights exactly like the eager iteration."""
import copy

import pytest
import torch

from helpers import G_KWARGS, D_KWARGS

pytestmark = pytest.mark.gpu


def _small_models():
    import os
    os.environ["LAYOUTDETR_SYNTHETIC_TOKENIZER"] = "1"
    from layoutdetr_b200.training import networks_detr as nd
    kw_g = dict(G_KWARGS, bert_num_encoder_layers=2, max_text_length=64)
    kw_d = dict(D_KWARGS, bert_num_encoder_layers=2, max_text_length=64)
    torch.manual_seed(0)
    return nd.Generator(**kw_g).cuda(), nd.Discriminator(**kw_d).cuda()


def test_graph_replay_matches_eager():
    from layoutdetr_b200 import engine
    from layoutdetr_b200.synthetic import make_inputs
    from layoutdetr_b200.training.trainer import Trainer, GraphedStep
    hb = [make_inputs(2, n_valid=8, seed=s) for s in (1, 2, 3)]
    zs = [torch.randn((2, 9, 4), device="cuda", generator=torch.Generator(device="cuda").manual_seed(i)) for i in range(6)]
    results = []
    for mode in ("eager", "graph"):
        engine.clear_cache()
        G, D = _small_models()
        tr = Trainer(G, D, torch.device("cuda"), batch_size=2, lr=1e-3)
        gs = GraphedStep(tr) if mode == "graph" else None
        for it in range(3):
            b = hb[it]
            if gs is not None:
                gs.run(b, zs[2 * it], zs[2 * it + 1])
            else:
                dev_b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
                tr.iteration(dev_b, zs[2 * it], zs[2 * it + 1])
        torch.cuda.synchronize()
        results.append((tr.flat["G"].p.clone(), tr.flat["D"].p.clone(), tr.flat_ema.p.clone()))
    for a, b, name in zip(results[0], results[1], ("G", "D", "G_ema")):
        diff = float((a - b).abs().max())
        scale = float(a.abs().max())
        print(name, "max abs diff eager vs graph", diff, "scale", scale)
        assert diff < 2e-3 * 1e-3 * 50 + 1e-6, (name, diff)     # a few Adam steps of lr 1e-3; atomics reorder fp32 sums
