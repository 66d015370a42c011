"""CPU: host-side logic that needs no GPU — tokenizer stand-in, synthetic data contract, box losses vs the oracle,
Adam / EMA reference math used by the fused kernel tests, and the data-parallel gradient exchange under gloo (2 ranks)."""
import os

import pytest
import torch


def test_synthetic_tokenizer_contract():
    from layoutdetr_b200.synthetic import SyntheticTokenizer
    tok = SyntheticTokenizer()
    enc = tok(["hello", "", "x" * 400], padding="max_length", truncation=True, max_length=256, return_tensors="pt")
    assert enc.input_ids.shape == (3, 256) and enc.attention_mask.shape == (3, 256)
    assert enc.input_ids[0, 0] == 101 and enc.input_ids[0, 6] == 102 and enc.attention_mask[0].sum() == 7
    assert enc.attention_mask[1].sum() == 2                      # empty string -> [CLS][SEP]
    assert enc.attention_mask[2].sum() == 256 and enc.input_ids[2, 255] == 102
    assert len(tok) == 30524 and tok.bos_token_id == 30522 and tok.pad_token_id == 0
    assert int(enc.input_ids.max()) < 30522


def test_make_inputs_contract():
    from layoutdetr_b200.synthetic import make_inputs
    a = make_inputs(3, n_valid=8, seed=5)
    b = make_inputs(3, n_valid=8, seed=5)
    assert torch.equal(a["background"], b["background"]) and a["bbox_text"] == b["bbox_text"]
    assert a["padding_mask"].shape == (3, 9) and a["padding_mask"][:, 8:].all() and not a["padding_mask"][:, :8].any()
    assert all(t == "" for row in a["bbox_text"] for t in row[8:])
    assert float(a["bbox_real"][:, 8:].abs().max()) == 0.0


def _host_box_lib(tmp_path_factory=None):
    """g++ build of the product's box-loss arithmetic (csrc/box_loss_math.h, shared with the CUDA kernels)."""
    import ctypes, subprocess, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(tempfile.mkdtemp(prefix="ld_boxloss_"), "box_loss_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(root, "tests", "native", "box_loss_host.cpp")], check=True)
    return ctypes.CDLL(out)


def test_box_loss_arithmetic_matches_oracle_values_and_autograd():
    """The arithmetic the box-loss kernels compile (value + analytic Jacobian) against autograd of the oracle restatement of
    metrics/metric_layoutnet.py:153-201,245-275 — ragged layouts, a single valid slot, junk in padded slots."""
    import ctypes
    from oracle import layoutdetr_oracle as O
    lib = _host_box_lib()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    g = torch.Generator().manual_seed(0)
    for trial in range(8):
        B, N = 5, 9
        bbox = torch.rand((B, N, 4), generator=g) * torch.tensor([0.6, 0.6, 0.5, 0.3]) + torch.tensor([0.2, 0.2, 0.05, 0.03])
        mask = torch.ones((B, N), dtype=torch.bool)
        mask[:, 8:] = False
        mask[1, 3:] = False
        mask[2, 1:] = False
        if trial % 2 == 0:
            bbox[:, 8:] = torch.rand((B, 1, 4), generator=g)          # the generator's output in padded slots is arbitrary
        b2 = bbox.clone().requires_grad_(True)
        ov, al = O.compute_overlap(b2, mask), O.compute_alignment(b2, mask)
        w1, w2 = torch.rand(B, generator=g), torch.rand(B, generator=g)
        ((ov * w1).sum() + (al * w2).sum()).backward()
        v8 = mask.to(torch.uint8).contiguous()
        o, a, jo, ja = torch.empty(B), torch.empty(B), torch.empty(B, N, 4), torch.empty(B, N, 4)
        lib.host_layout_losses(P(bbox), P(v8), ctypes.c_long(B), N, P(o), P(a), P(jo), P(ja))
        torch.testing.assert_close(o, ov.detach(), atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(a, al.detach(), atol=1e-6, rtol=1e-5)
        grad = jo * w1[:, None, None] + ja * w2[:, None, None]
        assert torch.isfinite(grad).all()
        torch.testing.assert_close(grad, b2.grad, atol=2e-5, rtol=1e-4)
        f, r = bbox[mask].contiguous(), bbox.flip(0)[mask].contiguous()
        f2 = f.clone().requires_grad_(True)
        ref = O.generalized_iou_loss(f2, r)
        ref.backward()
        lo, jf = torch.empty(1), torch.empty(f.shape[0], 4)
        lib.host_giou_loss(P(f), P(r), ctypes.c_long(f.shape[0]), P(lo), P(jf))
        torch.testing.assert_close(lo[0], ref.detach(), atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(jf, f2.grad, atol=1e-6, rtol=1e-4)


def test_box_ops_refuse_cpu_tensors():
    from layoutdetr_b200 import box_ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        box_ops.layout_losses(torch.rand(2, 9, 4), torch.ones(2, 9, dtype=torch.bool))


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the exchange step of Trainer._phase on a flat gradient buffer: SUM all-reduce, then 1/world folded into the optimizer
    torch.manual_seed(rank)
    flat_g = torch.randn(1000)
    mine = flat_g.clone()
    dist.all_reduce(flat_g)
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    expect = sum(gathered)
    ok = torch.allclose(flat_g, expect)
    # identical update on every rank => replicas stay bit-identical (reference check_ddp_consistency, torch_utils/misc.py:183)
    p = torch.ones(1000)
    p -= 1e-3 * (flat_g / world)
    ps = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(ps, p)
    same = all(torch.equal(ps[0], q) for q in ps)
    if rank == 0:
        out.put((ok, same))
    dist.destroy_process_group()


def test_data_parallel_gradient_exchange_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, same = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok and same


def test_stream_cta_limit_table():
    """ld_set_stream_cta_limit / ld_get_stream_cta_limit: per-stream cap, clamped to the SM count, cleared with <= 0."""
    import ctypes
    from layoutdetr_b200 import _lib
    lib = _lib.lib()
    s1, s2 = ctypes.c_void_p(0x1000), ctypes.c_void_p(0x2000)
    full = lib.ld_get_stream_cta_limit(s1)
    assert full >= 1
    assert lib.ld_set_stream_cta_limit(s1, 100) == 0 and lib.ld_set_stream_cta_limit(s2, 7) == 0
    assert lib.ld_get_stream_cta_limit(s1) == min(100, full) and lib.ld_get_stream_cta_limit(s2) == 7
    assert lib.ld_set_stream_cta_limit(s1, 10 ** 6) == 0 and lib.ld_get_stream_cta_limit(s1) == full
    lib.ld_set_stream_cta_limit(s1, 0)
    lib.ld_set_stream_cta_limit(s2, -1)
    assert lib.ld_get_stream_cta_limit(s1) == full and lib.ld_get_stream_cta_limit(s2) == full


def test_engine_refresh_stale_recomputes_in_place():
    """engine.refresh_stale: a derived shadow is recomputed into the SAME storage when its source changed."""
    import torch
    from layoutdetr_b200 import engine as E
    p = torch.nn.Parameter(torch.arange(6.0).reshape(2, 3))
    v = E.derived((p,), "twice", lambda t: (t.detach() * 2).clone())
    ptr = v.data_ptr()
    assert E.refresh_stale({id(p)}) == 0
    with torch.no_grad():
        p.add_(1.0)                       # bumps the version counter
    assert E.refresh_stale({id(object())}) == 0       # not ours: left alone
    assert E.refresh_stale({id(p)}) == 1
    v2 = E.derived((p,), "twice", lambda t: None)     # cache hit, no recompute
    assert v2.data_ptr() == ptr and torch.equal(v2, (p.detach() * 2))


def test_engine_tables_hold_no_strong_references():
    """shadows die with their module (a training run deep-copies G_ema per snapshot / sweep), an entry is never served to
    another object that re-uses the id(), and a managed (flat-storage) shadow follows torch-level in-place writes."""
    import gc
    import weakref
    import torch
    from layoutdetr_b200 import engine as E
    lin = torch.nn.Linear(8, 4)
    E.derived((lin.weight,), "probe", lambda t: t.detach() * 3)
    E.derived((lin.weight, lin.bias), "pair", lambda w, b: torch.cat([w.detach().reshape(-1), b.detach()]))
    n0 = len(E._shadow)
    wref = weakref.ref(lin.weight)
    del lin
    gc.collect()
    assert wref() is None and len(E._shadow) == n0 - 2
    # same key, different object: treated as a miss
    a = torch.nn.Parameter(torch.ones(2, 8))
    key = ((id(a),), "k")
    E.derived((a,), "k", lambda t: t.detach() + 1)
    ent = E._shadow[key]
    b = torch.nn.Parameter(torch.zeros(2, 8))
    E._shadow[((id(b),), "k")] = ent                  # simulate a recycled id(): the entry of a dead object under b's key
    assert torch.equal(E.derived((b,), "k", lambda t: t.detach() + 1), torch.ones(2, 8))
    # managed shadow: refreshed when the parameter is written behind the optimizer kernel's back
    p = torch.nn.Parameter(torch.arange(16.0).reshape(2, 8))
    shadow = p.detach().to(torch.bfloat16).clone()
    E.register_managed(p, shadow)
    assert E._managed_shadow(p) is shadow
    with torch.no_grad():
        p.mul_(2.0)
    assert torch.equal(E._managed_shadow(p).float(), p.detach()) and E._managed_shadow(p) is shadow


def _eval_worker(rank, world, port, out):
    import torch.distributed as dist
    from layoutdetr_b200.metrics.eval_sweep import LayoutEvalAccumulator
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    feats = torch.randn((2, 40, 16), generator=g, dtype=torch.float64)       # [real | fake], 40 layouts, split over the ranks
    vals = torch.rand((4, 40), generator=g, dtype=torch.float64)
    acc = LayoutEvalAccumulator(feature_dim=16, device="cpu")
    sl = slice(rank, 40, world)                                               # the reference's interleaved item subset
    acc.n += vals[:, sl].shape[1]
    acc.sums += vals[:, sl].sum(dim=1)
    for i in range(2):
        acc.raw_mean[i] += feats[i, sl].sum(dim=0)
        acc.raw_cov[i] += feats[i, sl].t() @ feats[i, sl]
    res = acc.all_reduce().result()
    if rank == 0:
        torch.save(dict(res=res, feats=feats, vals=vals), out)
    dist.destroy_process_group()


def test_eval_sweep_exchange_step_two_ranks_gloo(tmp_path):
    """The eval sweep's only collective: one SUM all-reduce of the running statistics; 2 ranks == 1 process on all items."""
    import torch.multiprocessing as mp
    from oracle import layoutdetr_oracle as O
    out = str(tmp_path / "eval.pt")
    mp.spawn(_eval_worker, args=(2, 29517 + os.getpid() % 1000, out), nprocs=2, join=True)
    d = torch.load(out, weights_only=False)
    res, feats, vals = d["res"], d["feats"], d["vals"]
    assert res["num_items"] == 40
    for k, i in [("overlap", 0), ("alignment", 1), ("layoutwise_iou", 2), ("layoutwise_docsim", 3)]:
        assert abs(res[k] - float(vals[i].mean())) < 1e-12
    x = feats.numpy()
    mean = [x[i].sum(0) / 40 for i in range(2)]
    cov = [x[i].T @ x[i] / 40 - __import__("numpy").outer(mean[i], mean[i]) for i in range(2)]
    ref = O.layout_fid(mean[1], cov[1], mean[0], cov[0])
    assert abs(res["layout_fid"] - ref) <= 1e-9 * max(1.0, abs(ref))


def test_box_loss_arithmetic_edge_cases_match_oracle_autograd():
    """Ties and degenerate geometry, where autograd's conventions matter (max / min ties split evenly, first minimum wins,
    nan_to_num / masked_fill stop gradients): all-zero padded slots (the dataset's convention), duplicated boxes, shared
    edges, a box nested in another, a single valid slot.  Two degenerate inputs are NOT reproduced and never occur on the path
    (boxes are sigmoid outputs, every layout has >= 1 element): a valid box of zero area (reference gradient: NaN, here 0 for
    the 0/0 terms) and a layout without any valid slot (value NaN in both; reference gradient 0, here NaN)."""
    import ctypes
    from oracle import layoutdetr_oracle as O
    lib = _host_box_lib()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    g = torch.Generator().manual_seed(1)
    base = torch.rand((2, 9, 4), generator=g) * 0.4 + 0.2
    mask6 = torch.ones(2, 9, dtype=torch.bool)
    mask6[:, 6:] = False
    single = torch.zeros(2, 9, dtype=torch.bool)
    single[:, 0] = True
    cases = []
    b = base.clone(); b[:, 6:] = 0; cases.append((b, mask6))
    b = base.clone(); b[:, 1] = b[:, 0]; cases.append((b, mask6))
    b = base.clone(); b[:, 1, 0] = b[:, 0, 0]; b[:, 1, 2] = b[:, 0, 2]; cases.append((b, mask6))
    b = base.clone(); b[:, 1] = torch.tensor([0.5, 0.5, 0.1, 0.05]); b[:, 0] = torch.tensor([0.5, 0.5, 0.4, 0.3]); cases.append((b, mask6))
    cases.append((base.clone(), single))
    for bbox, mask in cases:
        B, N, _ = bbox.shape
        ref = bbox.clone().requires_grad_(True)
        ov, al = O.compute_overlap(ref, mask), O.compute_alignment(ref, mask)
        (ov.sum() + al.sum()).backward()
        bb, v8 = bbox.contiguous(), mask.to(torch.uint8).contiguous()
        o, a, jo, ja = torch.empty(B), torch.empty(B), torch.empty(B, N, 4), torch.empty(B, N, 4)
        lib.host_layout_losses(P(bb), P(v8), ctypes.c_long(B), N, P(o), P(a), P(jo), P(ja))
        torch.testing.assert_close(o, ov.detach(), atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(a, al.detach(), atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(jo + ja, ref.grad, atol=5e-6, rtol=1e-4)
    f = base[0, :6].clone().contiguous()
    r = f.clone()
    r[:3] += 0.05                                                    # rows 3..5: fake == real exactly (every max / min is a tie)
    fr = f.clone().requires_grad_(True)
    ref = O.generalized_iou_loss(fr, r)
    ref.backward()
    lo, jf = torch.empty(1), torch.empty(6, 4)
    lib.host_giou_loss(P(f), P(r), ctypes.c_long(6), P(lo), P(jf))
    torch.testing.assert_close(lo[0], ref.detach(), atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(jf, fr.grad, atol=1e-6, rtol=1e-4)


def _exchange_worker(rank, world, port, out):
    """Two ranks (gloo, CPU): a toy network whose gradients reach the flat buffer through BOTH write paths of the product —
    torch's AccumulateGrad (post-accumulate hooks) and a hand-written `_Fn` backward adding straight into engine.grad_buffer —
    exchanged by exchange.GradExchange: one trace pass, then passes armed with the traced counts."""
    import torch.distributed as dist
    from layoutdetr_b200 import engine as E
    from layoutdetr_b200 import functional as Fn
    from layoutdetr_b200.exchange import GradExchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class DirectLinear(Fn._Fn):                      # weight gradient written in place, None returned to autograd (functional.py convention)
        @staticmethod
        def forward(ctx, x, w):
            ctx.w = w
            ctx.save_for_backward(x)
            return x @ w.detach().t()

        @staticmethod
        def backward(ctx, dy):
            x, = ctx.saved_tensors
            E.grad_buffer(ctx.w).add_(dy.t() @ x)
            return dy @ ctx.w.detach(), None

    torch.manual_seed(0)
    shapes = [(8, 8)] * 6 + [(8,)] * 3
    params = [torch.nn.Parameter(torch.randn(s) * 0.3) for s in shapes]
    offs, total = [], 0
    for p in params:
        offs.append(total)
        total += (p.numel() + 7) // 8 * 8
    g = torch.zeros(total)
    for p, o in zip(params, offs):
        p.grad = g[o:o + p.numel()].view(p.shape)
        p.requires_grad_(False)                     # as the Trainer leaves them between phases
    ex = GradExchange(g, params, offs, world=world, bucket_mb=128 * 4 / (1 << 20), overlap=True, name="toy")   # 128 floats per bucket
    nb = len(ex.bounds)

    def backward_pass(x, skip_last=False):
        for p in params:
            p.requires_grad_(True)
        h = x
        for i in range(6):
            w = params[i]
            h = DirectLinear.apply(h, w) if i % 2 == 0 else h @ w.t()          # in-place path / AccumulateGrad path
            if i < 3:
                h = h + params[6 + i]
            h = torch.tanh(h)
        if not skip_last:
            h = DirectLinear.apply(h, params[0])                               # weight 0 is used twice: two writes into its bucket
        h.sum().backward()

    def local_grads(x):
        g.zero_()
        backward_pass(x)
        return g.clone()

    results = []
    xs = [torch.randn(4, 8, generator=torch.Generator().manual_seed(10 * it + rank)) for it in range(3)]
    expected = None
    for it in range(3):
        mine = local_grads(xs[it])
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        g.zero_()
        ex.begin(expected)
        backward_pass(xs[it])
        counts = ex.finish()
        results.append(bool(torch.allclose(g, sum(gathered), atol=1e-6)))
        if it == 0:
            assert ex.stats["early"] == 0                  # the trace pass exchanges after the backward pass
            expected = counts
    early = ex.stats["early"]
    # a pass with FEWER writes than traced: nothing is lost (late exchange); a write AFTER a bucket went out must raise
    g.zero_()
    ex.begin(expected)
    backward_pass(xs[0], skip_last=True)
    ex.finish()
    g.zero_()
    ex.begin([max(0, c - 1) if i == ex.bucket_of[id(params[0])] else c for i, c in enumerate(expected)])
    raised = False
    try:
        backward_pass(xs[0])
    except RuntimeError as e:
        raised = "after its all-reduce" in str(e)
    if not raised:
        ex.finish()
    ex.close()
    if rank == 0:
        out.put((results, nb, early, raised))
    dist.barrier() if raised is False else None
    dist.destroy_process_group()


def test_bucketed_gradient_exchange_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, 29653, q)) for r in range(2)]
    for p in procs:
        p.start()
    results, nb, early, raised = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    assert nb >= 3, nb
    assert results == [True, True, True], results
    assert early == 2 * nb, (early, nb)              # both armed passes sent every bucket from inside the backward pass
    assert raised


def test_conv_box_rule_and_split_k_model():
    """Host-side rules of the convolution / weight-gradient paths: a box of output pixels must be whole rows / whole images
    (mirrors make_conv_map in csrc/gemm_sm100.cu), and the split-K factor stays within its bounds and follows the measured optima of
    profiles/r2_splitk_probe.txt."""
    from layoutdetr_b200 import engine as E, kernels as K
    ok = K.conv_box_ok
    assert ok(64, 64, 1, 128) and ok(32, 32, 2, 128) and ok(16, 16, 1, 128) and ok(8, 8, 1, 128)       # ResNet-50 stages at 256^2
    assert ok(256, 256, 1, 128) and ok(128, 128, 1, 64) and ok(4, 4, 1, 64)
    assert not ok(12, 12, 1, 128) and not ok(24, 24, 1, 128) and not ok(7, 7, 1, 64)
    assert not ok(64, 192, 1, 128)                              # a row of 192 pixels is not a whole number of 128-pixel boxes
    assert not ok(8, 128, 4, 128)                               # box of 128 pixels x stride 4 exceeds the 256-element TMA box limit
    old = E._SM_COUNT
    E._SM_COUNT = 148
    try:
        for shape in [(256, 64, 65536), (64, 576, 65536), (512, 128, 16384), (256, 2304, 4096), (512, 4608, 1024), (768, 768, 32768),
                      (30524, 768, 32768), (256, 256, 160), (32, 288, 1048576), (8, 8, 64), (3, 32, 1048576)]:
            sk = E.wgrad_split_k(*shape)
            kb = (shape[2] + 63) // 64
            assert 1 <= sk <= max(1, kb // 2), (shape, sk)
        assert E.wgrad_split_k(512, 4608, 1024) == 1            # many tiles, short reduction: never split (measured 15 vs 34 us)
        assert E.wgrad_split_k(30524, 768, 32768) == 1          # the LM-head weight gradient fills the chip by itself
        assert 32 <= E.wgrad_split_k(256, 64, 65536) <= 64      # measured optimum 32-64 (20 us; the round-1 rule chose 148: 40 us)
        assert 4 <= E.wgrad_split_k(768, 768, 32768) <= 8
    finally:
        E._SM_COUNT = old
