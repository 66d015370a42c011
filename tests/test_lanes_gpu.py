"""GPU: the lane scheduler (layoutdetr_b200/lanes.py) must not change results — one training iteration issued on
parallel streams (text-encoder lane, forward / backward branches, real-sample lane) gives the loss terms and the weight
update of the single-stream schedule, eagerly and as a captured CUDA graph."""
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


def _run(level, graph, iters=3):
    from layoutdetr_b200 import engine
    from layoutdetr_b200.lanes import LANES
    from layoutdetr_b200.synthetic import make_inputs
    from layoutdetr_b200.training.trainer import Trainer, GraphedStep
    LANES.configure(level=level, text_ctas=96, lm_ctas=96)
    engine.clear_cache()
    G, D = _small_models()
    tr = Trainer(G, D, torch.device("cuda"), batch_size=2, lr=1e-5)
    init = (tr.flat["G"].p.clone(), tr.flat["D"].p.clone())
    hb = [make_inputs(2, n_valid=8, seed=s) for s in (1, 2, 3)]
    zs = [torch.randn((2, 9, 4), device="cuda", generator=torch.Generator(device="cuda").manual_seed(i)) for i in range(6)]
    gs = GraphedStep(tr) if graph else None
    losses = []
    for it in range(iters):
        b = hb[it]
        if gs is not None:
            out = gs.run(b, zs[2 * it], zs[2 * it + 1])
        else:
            dev_b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
            out = tr.iteration(dev_b, zs[2 * it], zs[2 * it + 1])
        torch.cuda.synchronize()
        losses.append({ph + "/" + k: float(v.float().mean()) for ph in ("Gmain", "Dmain") for k, v in out[ph].items()})
    upd = tuple(f - i for f, i in zip((tr.flat["G"].p, tr.flat["D"].p), init))
    return losses, upd


@pytest.fixture(scope="module")
def single_stream():
    return _run(0, graph=False)


@pytest.mark.parametrize("level,graph", [(1, False), (2, False), (3, False), (3, True)])
def test_lanes_match_single_stream(single_stream, level, graph):
    from layoutdetr_b200.lanes import LANES
    try:
        losses, upd = _run(level, graph)
    finally:
        LANES.configure(level=3, text_ctas=128, lm_ctas=128)
    ref_losses, ref_upd = single_stream
    for it, (a, b) in enumerate(zip(losses, ref_losses)):
        assert set(a) == set(b)
        worst = max((abs(a[k] - b[k]) / (abs(b[k]) + 1e-3), k) for k in a)
        print("level", level, "graph", graph, "iter", it, "worst loss-term rel diff %.4g (%s)" % worst)
        assert worst[0] < 5e-2, worst
    for ue, ur, name in zip(upd, ref_upd, ("G", "D")):
        rel = float((ue - ur).norm() / (ur.norm() + 1e-20))
        print("level", level, "graph", graph, name, "relative L2 difference of the accumulated update: %.4f" % rel)
        assert rel < 0.15, (name, rel)      # Adam turns ~0 gradients (atomics-order noise) into +-lr steps; real bugs give O(1)


def test_stream_cta_limit_caps_the_gemm_grid():
    """A capped stream still computes the same GEMM (static tile striding works for any grid size)."""
    import ctypes
    from layoutdetr_b200 import _lib, kernels as K
    s = torch.cuda.Stream()
    a = torch.randn((4096, 512), device="cuda").to(torch.bfloat16)
    w = torch.randn((768, 512), device="cuda").to(torch.bfloat16)
    ref = K.linear(a, w)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().ld_set_stream_cta_limit(ctypes.c_void_p(s.cuda_stream), 24))
    assert _lib.lib().ld_get_stream_cta_limit(ctypes.c_void_p(s.cuda_stream)) == 24
    with torch.cuda.stream(s):
        out = K.linear(a, w)
    torch.cuda.synchronize()
    _lib.lib().ld_set_stream_cta_limit(ctypes.c_void_p(s.cuda_stream), 0)
    assert torch.equal(out, ref)
