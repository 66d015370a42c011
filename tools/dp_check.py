"""Data-parallel check on N >= 2 GPUs (run under torchrun): the overlapped, bucketed gradient exchange captured into the iteration's
CUDA graph (exchange.py) against the plain schedule (graph segments with one NCCL all-reduce of the whole buffer between them).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/dp_check.py

Per mode: small G / D in .eval() (deterministic kernels), lr = 0 (weights never move: every replay sees the same forward), the
same per-rank batches; compared: the exchanged flat gradient buffers (must agree to fp32 summation-order noise) and, after a few
steps at lr > 0, that all ranks hold bit-identical weights (reference check_ddp_consistency, torch_utils/misc.py:183).
Rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("LD_DP_CHECK_DUMP_AFTER", "240")), exit=True)     # a hang prints where every thread is
    import torch
    import torch.distributed as dist
    os.environ["LAYOUTDETR_SYNTHETIC_TOKENIZER"] = "1"
    os.environ.setdefault("LAYOUTDETR_SYNTHETIC_WEIGHTS", "1")
    os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
    from helpers import G_KWARGS, D_KWARGS
    from layoutdetr_b200 import engine
    from layoutdetr_b200.synthetic import make_inputs
    from layoutdetr_b200.training import networks_detr as nd
    from layoutdetr_b200.training.trainer import Trainer, GraphedStep
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    kw_g = dict(G_KWARGS, bert_num_encoder_layers=2, max_text_length=64)
    kw_d = dict(D_KWARGS, bert_num_encoder_layers=2, max_text_length=64)
    nobg = dict(Dreal_im_rec_weight=0.0)        # D's style-gradient atomics are not run-to-run reproducible (tools/debug_nondet.py)

    def run(overlap, lr, steps, train_mode=False):
        os.environ["LD_DP_OVERLAP"] = "1" if overlap else "0"
        os.environ["LD_DP_BUCKET_MB"] = "8"
        engine.clear_cache()
        torch.manual_seed(0)
        G = nd.Generator(**kw_g).to(dev).train(train_mode)
        D = nd.Discriminator(**kw_d).to(dev).train(train_mode)
        tr = Trainer(G, D, dev, batch_size=2 * world, num_gpus=world, lr=lr, loss_kwargs=nobg)
        gs = GraphedStep(tr)
        hb = make_inputs(2, n_valid=8, seed=5 + rank)
        zs = [torch.randn((2, 9, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(10 * rank + i)) for i in range(2)]
        for _ in range(steps):
            gs.run(hb, zs[0], zs[1])
        torch.cuda.synchronize()
        ent = dict(next(iter(gs.graphs.values())))
        n_graphs = len(ent["graphs"])
        ent = dict(exchange=ent.get("exchange"), graphs=[None] * n_graphs)
        gs.close()                  # graphs with captured NCCL collectives must not outlive the process group
        return tr, ent

    out = {}

    def log(msg):
        print("[dp_check rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)
    log("plain schedule ...")
    tr0, e0 = run(False, 0.0, 2)
    log("overlapped schedule ...")
    g_plain = [tr0.flat[n].g.clone() for n in ("G", "D")]
    tr1, e1 = run(True, 0.0, 2)
    g_over = [tr1.flat[n].g.clone() for n in ("G", "D")]
    out["graphs_plain"], out["graphs_overlap"] = len(e0["graphs"]), len(e1["graphs"])
    out["exchange"] = e1.get("exchange")
    out["rel_l2_G"], out["rel_l2_D"] = [float((a - b).norm() / (b.norm() + 1e-30)) for a, b in zip(g_over, g_plain)]
    # gradients differ between ranks before the exchange and agree after it
    gg = [torch.zeros_like(g_over[0]) for _ in range(world)]
    dist.all_gather(gg, g_over[0])
    out["grad_identical_across_ranks"] = bool(all(torch.equal(gg[0], x) for x in gg))
    del tr0, tr1, e0, e1
    log("real steps ...")
    # a few real steps, dropout live: replicas must stay bit-identical
    tr2, e2 = run(True, 1e-4, 4, train_mode=True)
    sums = torch.stack([f.p.view(torch.int32).to(torch.int64).sum() for f in (tr2.flat["G"], tr2.flat["D"], tr2.flat_ema)])
    allsums = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(allsums, sums)
    out["replicas_identical"] = bool(all(torch.equal(allsums[0], a) for a in allsums))
    out["world"] = world
    ok = out["rel_l2_G"] < 1e-5 and out["rel_l2_D"] < 1e-4 and out["grad_identical_across_ranks"] and out["replicas_identical"] \
        and out["graphs_overlap"] == 1 and (out["exchange"] or {}).get("G", {}).get("early", 0) > 0
    out["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
