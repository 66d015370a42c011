#!/usr/bin/env python
"""Sweep the lane-scheduler settings (layoutdetr_b200/lanes.py) on ONE set of models: for every configuration capture the
training iteration into a CUDA graph, check the first step's loss terms / weight update against the single-stream
schedule from the same weights, and time graph replays (L2 flushed between steps, CUDA events).

    python tools/lane_sweep.py [--batch 16] [--steps 6] [--configs "0;1,128;2,128,128;3,128,128"]

A config is level[,text_ctas[,lm_ctas[,priority]]].  Prints one JSON line per configuration.
"""
import argparse
import gc
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LAYOUTDETR_SYNTHETIC_TOKENIZER", "1")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--enc-layers", type=int, default=12)
    ap.add_argument("--configs", default="0;1,128;1,112;2,128,128;3,128,128;3,112,112;3,136,136;3,148,148;3,128,128,0")
    args = ap.parse_args()

    import torch
    import bench
    from layoutdetr_b200.lanes import LANES
    from layoutdetr_b200.synthetic import make_inputs
    from layoutdetr_b200.training import networks_detr as nd
    from layoutdetr_b200.training.trainer import Trainer, GraphedStep

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    gk = dict(bench.G_KWARGS, bert_num_encoder_layers=args.enc_layers)
    dk = dict(bench.D_KWARGS, bert_num_encoder_layers=args.enc_layers)
    G = nd.Generator(**gk).to(dev)
    D = nd.Discriminator(**dk).to(dev)
    B = args.batch
    tr = Trainer(G, D, dev, batch_size=B)
    hb = make_inputs(B, n_valid=8, seed=1)
    gz = torch.Generator(device=dev).manual_seed(1234)
    zs = [torch.randn((B, 9, 4), device=dev, generator=gz) for _ in range(2)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    LANES.configure(level=0)
    base_snap = None
    ref = None
    for cfg in args.configs.split(";"):
        parts = [int(x) for x in cfg.split(",")]
        level = parts[0]
        text_ctas = parts[1] if len(parts) > 1 else 128
        lm_ctas = parts[2] if len(parts) > 2 else 128
        prio = parts[3] if len(parts) > 3 else 1
        real_first = parts[4] if len(parts) > 4 else 0
        LANES.configure(level=level, text_ctas=text_ctas, lm_ctas=lm_ctas, high_priority=prio, real_first=real_first)
        gs = GraphedStep(tr)
        rec = dict(level=level, text_ctas=text_ctas, lm_ctas=lm_ctas, priority=prio, real_first=real_first)
        try:
            if base_snap is None:
                # one throw-away capture brings every cache to its steady state; then remember the weights
                gs.run(hb, zs[0], zs[1])
                torch.cuda.synchronize()
                base_snap = gs._snapshot()
                st0 = next(iter(gs.graphs.values()))["static"]
                gs._restore(base_snap, st0)
                gs.graphs.clear()
            else:
                gs._restore(base_snap, {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in dict(hb, z_g=zs[0], z_d=zs[1]).items()})
            p0 = {n: f.p.clone() for n, f in tr.flat.items()}
            out = gs.run(hb, zs[0], zs[1])                      # capture + first real step from the base weights
            torch.cuda.synchronize()
            terms = {ph + "/" + k: float(v.float().mean()) for ph in ("Gmain", "Dmain") for k, v in out[ph].items()}
            upd = {n: (tr.flat[n].p - p0[n]) for n in tr.flat}
            grads = {n: tr.flat[n].g.clone() for n in tr.flat}            # gradients of the step just taken
            if ref is None:
                ref = (terms, upd, grads)
            rec["grad_rel_l2_diff_vs_first"] = {n: float((grads[n] - ref[2][n]).norm() / (ref[2][n].norm() + 1e-20)) for n in grads}
            rec["grad_max_abs_diff_vs_first"] = {n: float((grads[n] - ref[2][n]).abs().max()) for n in grads}
            rec["loss_max_rel_diff_vs_first"] = max(abs(terms[k] - ref[0][k]) / (abs(ref[0][k]) + 1e-3) for k in terms)
            rec["update_rel_l2_diff_vs_first"] = {n: float((upd[n] - ref[1][n]).norm() / (ref[1][n].norm() + 1e-20)) for n in upd}
            for _ in range(2):
                gs.run_static()
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for i in range(args.steps):
                flush.zero_()
                ev[i][0].record()
                gs.run_static()
                ev[i][1].record()
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in ev)
            rec["ms_per_step"] = sum(ts) / len(ts)
            rec["ms_min"] = ts[0]
            rec["samples_per_s"] = B / (rec["ms_per_step"] * 1e-3)
            rec["launches"] = next(iter(gs.graphs.values()))["launches"]
        except Exception as e:      # keep sweeping: one bad configuration must not lose the others
            import traceback
            rec["error"] = repr(e)
            traceback.print_exc()
        print(json.dumps(rec), flush=True)
        del gs
        gc.collect()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
