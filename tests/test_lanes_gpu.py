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
    # .eval(): these tests compare SCHEDULES of the same arithmetic; dropout (live under .train(), masks addressed by a host-side site
    # counter that follows the issue order) is covered by tests/test_dropout_gpu.py
    return nd.Generator(**kw_g).cuda().eval(), nd.Discriminator(**kw_d).cuda().eval()


def _run(level, graph, iters=3, dry=0):
    from layoutdetr_b200 import engine
    from layoutdetr_b200.lanes import LANES
    from layoutdetr_b200.synthetic import make_inputs
    from layoutdetr_b200.training.trainer import Trainer, GraphedStep
    LANES.configure(level=level, text_ctas=96, lm_ctas=96, dry=dry)
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
        LANES.configure(level=3, text_ctas=0, lm_ctas=0, dry=0)
    ref_losses, ref_upd = single_stream
    for it, (a, b) in enumerate(zip(losses, ref_losses)):
        assert set(a) == set(b)
        worst = max((abs(a[k] - b[k]) / (abs(b[k]) + 1e-3), k) for k in a)
        print("level", level, "graph", graph, "iter", it, "worst loss-term rel diff %.4g (%s)" % worst)
        assert worst[0] < (1e-3 if it == 0 else 0.1), worst      # iteration 0: same weights -> same forward
    # The accumulated weight update is only reported: with the reference's Adam (beta1 = 0, eps = 1e-8) every parameter
    # moves by +-lr whatever the size of its gradient, so bf16 summation-order differences on small gradients flip steps;
    # the schedules are pinned at gradient level by test_lane_gradients_match_single_stream below.
    for ue, ur, name in zip(upd, ref_upd, ("G", "D")):
        rel = float((ue - ur).norm() / (ur.norm() + 1e-20))
        print("level", level, "graph", graph, name, "relative L2 difference of the accumulated update: %.4f" % rel)
        assert rel < 0.5, (name, rel)


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


def _grads(level, graph, dry=0, loss_kwargs=None):
    """Flat gradient buffers after the last iteration of a Trainer with lr = 0 (weights never move, so every schedule sees
    the same forward; the first iteration of a Trainer is single-stream by design)."""
    from layoutdetr_b200 import engine
    from layoutdetr_b200.lanes import LANES
    from layoutdetr_b200.synthetic import make_inputs
    from layoutdetr_b200.training.trainer import Trainer, GraphedStep
    LANES.configure(level=level, text_ctas=96, lm_ctas=96, dry=dry)
    engine.clear_cache()
    G, D = _small_models()
    tr = Trainer(G, D, torch.device("cuda"), batch_size=2, lr=0.0, loss_kwargs=loss_kwargs)
    hb = make_inputs(2, n_valid=8, seed=5)
    zs = [torch.randn((2, 9, 4), device="cuda", generator=torch.Generator(device="cuda").manual_seed(i)) for i in range(2)]
    if graph:
        gs = GraphedStep(tr)
        gs.run(hb, zs[0], zs[1])
        gs.run(hb, zs[0], zs[1])
    else:
        dev_b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()}
        for _ in range(3):
            tr.iteration(dev_b, zs[0], zs[1])
    torch.cuda.synchronize()
    return tr.flat["G"].g.clone(), tr.flat["D"].g.clone()


def _rel(a, b):
    return [float((x - y).norm() / (y.norm() + 1e-20)) for x, y in zip(a, b)]


def test_lane_gradients_match_single_stream():
    """Race detector + parity of the schedules, on the flat gradient buffers of one iteration.

    * Lanes on real streams vs the SAME issue order on one stream (dry run) must agree to fp32 atomics noise (1e-5) —
      anything larger is a missing dependency between lanes.  This part runs with the background-reconstruction weight
      at 0: with it on, D's gradients are only reproducible to ~1e-2 even run-to-run on ONE stream, because the
      fp32 atomics of the style-gradient reduction (channel_dot) flip single bf16 roundings that the 8-layer mapping
      network amplifies (measured: tools/debug_nondet.py); G's gradients stay at 1e-8 either way.
    * Any schedule vs the single-stream one, full loss: G tight, D within that amplification bound (adds the order in
      which autograd sums bf16 gradients meeting at the token output and the order of the two Dmain backward passes)."""
    from layoutdetr_b200.lanes import LANES
    nobg = dict(Dreal_im_rec_weight=0.0)
    try:
        ref = _grads(0, False)
        ref_nobg = _grads(0, False, loss_kwargs=nobg)
        print("single-stream run-to-run gradient noise (rel L2), no bg term: G %.3e  D %.3e"
              % tuple(_rel(_grads(0, False, loss_kwargs=nobg), ref_nobg)))
        for level in (1, 2, 3):
            dry = _grads(level, False, dry=1, loss_kwargs=nobg)
            for graph in (False, True) if level == 3 else (False,):
                r_dry = _rel(_grads(level, graph, loss_kwargs=nobg), dry)
                r_ref = _rel(_grads(level, graph), ref)
                print("level %d graph %s: rel L2 vs dry run (no bg term) G %.3e D %.3e | vs single stream (full loss) G %.3e D %.3e"
                      % (level, graph, r_dry[0], r_dry[1], r_ref[0], r_ref[1]))
                assert max(r_dry) < 1e-5, (level, graph, r_dry)
                assert r_ref[0] < 1e-5 and r_ref[1] < 5e-2, (level, graph, r_ref)
    finally:
        LANES.configure(level=3, text_ctas=0, lm_ctas=0, dry=0)
