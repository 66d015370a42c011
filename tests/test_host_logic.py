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


def test_box_losses_match_oracle_and_have_finite_grads():
    from layoutdetr_b200 import box_ops
    from oracle import layoutdetr_oracle as O
    g = torch.Generator().manual_seed(0)
    bbox = torch.rand((4, 9, 4), generator=g) * 0.5 + 0.2
    mask = torch.ones((4, 9), dtype=torch.bool)
    mask[:, 7:] = False
    mask[1, 3:] = False
    b1 = bbox.clone().requires_grad_(True)
    b2 = bbox.clone().requires_grad_(True)
    ours = box_ops.overlap(b1, mask).sum() + box_ops.alignment(b1, mask).sum() + box_ops.giou_loss(b1[mask], bbox.flip(0)[mask])
    ref = O.compute_overlap(b2, mask).sum() + O.compute_alignment(b2, mask).sum() + O.generalized_iou_loss(b2[mask], bbox.flip(0)[mask])
    torch.testing.assert_close(ours, ref, atol=1e-5, rtol=1e-5)
    ours.backward()
    ref.backward()
    assert torch.isfinite(b1.grad).all()
    torch.testing.assert_close(b1.grad, b2.grad, atol=1e-4, rtol=1e-4)


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
    v = E.derived((p,), "twice", lambda: (p.detach() * 2).clone())
    ptr = v.data_ptr()
    assert E.refresh_stale({id(p)}) == 0
    with torch.no_grad():
        p.add_(1.0)                       # bumps the version counter
    assert E.refresh_stale({id(object())}) == 0       # not ours: left alone
    assert E.refresh_stale({id(p)}) == 1
    v2 = E.derived((p,), "twice", lambda: None)       # cache hit, no recompute
    assert v2.data_ptr() == ptr and torch.equal(v2, (p.detach() * 2))
