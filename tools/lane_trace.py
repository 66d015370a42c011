#!/usr/bin/env python
"""Timeline of one captured training iteration (CUDA-graph replay) with the lane scheduler: which stream ran what, when.

    python tools/lane_trace.py [--lanes 3] [--out gpurun_out/lane_trace] [--bins 120]

Uses torch.profiler (CUPTI) on two graph replays; writes <out>.json (kernels of the 2nd replay: name, stream, start us,
duration us, grid) and prints a text summary: span of the step, busy time per stream, an SM-occupancy estimate
(sum over running kernels of min(grid, 148) / 148) per time bin, and the longest stretches with low occupancy.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LAYOUTDETR_SYNTHETIC_TOKENIZER", "1")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lanes", type=int, default=3)
    ap.add_argument("--text-ctas", type=int, default=128)
    ap.add_argument("--lm-ctas", type=int, default=128)
    ap.add_argument("--priority", type=int, default=1)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--bins", type=int, default=110)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "lane_trace"))
    args = ap.parse_args()

    import torch
    from torch.profiler import profile, ProfilerActivity
    import bench
    from layoutdetr_b200.lanes import LANES
    from layoutdetr_b200.synthetic import make_inputs
    from layoutdetr_b200.training import networks_detr as nd
    from layoutdetr_b200.training.trainer import Trainer, GraphedStep

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    LANES.configure(level=args.lanes, text_ctas=args.text_ctas, lm_ctas=args.lm_ctas, high_priority=args.priority)
    G = nd.Generator(**bench.G_KWARGS).to(dev)
    D = nd.Discriminator(**bench.D_KWARGS).to(dev)
    B = args.batch
    tr = Trainer(G, D, dev, batch_size=B)
    hb = make_inputs(B, n_valid=8, seed=1)
    gz = torch.Generator(device=dev).manual_seed(1234)
    zs = [torch.randn((B, 9, 4), device=dev, generator=gz) for _ in range(2)]
    from layoutdetr_b200 import kernels as K
    gs = GraphedStep(tr)
    gs.run(hb, zs[0], zs[1])                      # brings the caches / the lane streams to their steady state
    gs.graphs.clear()
    orig_graph = torch.cuda.graph

    class _logged_graph(orig_graph):              # log GEMM shapes of the CAPTURE only (not of the eager warm-up iterations)
        def __enter__(self):
            r = super().__enter__()
            K.GEMM_LOG = []
            return r

        def __exit__(self, *a):
            self_log = K.GEMM_LOG
            K.GEMM_LOG = None
            _logged_graph.log = self_log
            return super().__exit__(*a)

    torch.cuda.graph = _logged_graph
    gs.run(hb, zs[0], zs[1])
    torch.cuda.graph = orig_graph
    gemm_log = _logged_graph.log
    for _ in range(2):
        gs.run_static()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        gs.run_static()
        torch.cuda.synchronize()
    tmp = args.out + ".chrome.json"
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    prof.export_chrome_trace(tmp)
    with open(tmp) as f:
        tr_json = json.load(f)
    os.remove(tmp)
    ks = [e for e in tr_json["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ks.sort(key=lambda e: e["ts"])
    t0 = ks[0]["ts"]
    rows = []
    for e in ks:
        a = e.get("args", {})
        grid = a.get("grid", [1, 1, 1])
        n = 1
        for g in grid:
            n *= int(g)
        rows.append(dict(name=e["name"][:60], stream=int(a.get("stream", -1)), t=round(e["ts"] - t0, 2), d=round(e["dur"], 2), grid=n))
    with open(args.out + ".json", "w") as f:
        json.dump(rows, f)
    span = max(r["t"] + r["d"] for r in rows)
    print("kernels %d, span %.2f ms, sum of kernel durations %.2f ms" % (len(rows), span / 1e3, sum(r["d"] for r in rows) / 1e3))
    streams = {}
    for r in rows:
        s = streams.setdefault(r["stream"], dict(n=0, busy=0.0, first=r["t"], last=0.0, big=0.0))
        s["n"] += 1
        s["busy"] += r["d"]
        s["last"] = max(s["last"], r["t"] + r["d"])
        if r["grid"] >= 100 and "gemm" in r["name"]:
            s["big"] += r["d"]
    print("%8s %6s %9s %9s %9s %9s" % ("stream", "n", "busy ms", "bigGEMM", "first ms", "last ms"))
    for k, s in sorted(streams.items(), key=lambda kv: kv[1]["first"]):
        print("%8d %6d %9.2f %9.2f %9.2f %9.2f" % (k, s["n"], s["busy"] / 1e3, s["big"] / 1e3, s["first"] / 1e3, s["last"] / 1e3))
    nb = args.bins
    w = span / nb
    occ = [0.0] * nb
    nrun = [0.0] * nb
    per_stream = {k: [0.0] * nb for k in streams}
    for r in rows:
        f = min(r["grid"], 148) / 148.0
        b0, b1 = int(r["t"] // w), min(nb - 1, int((r["t"] + r["d"]) // w))
        for b in range(b0, b1 + 1):
            lo, hi = max(r["t"], b * w), min(r["t"] + r["d"], (b + 1) * w)
            if hi > lo:
                occ[b] += f * (hi - lo) / w
                nrun[b] += (hi - lo) / w
                per_stream[r["stream"]][b] += (hi - lo) / w
    print("time bins of %.2f ms: SM-occupancy estimate (sum of min(grid,148)/148 over running kernels) | kernels in flight | busy streams" % (w / 1e3))
    order = [k for k, _ in sorted(streams.items(), key=lambda kv: kv[1]["first"])]
    for b in range(nb):
        marks = "".join("#" if per_stream[k][b] > 0.5 else ("+" if per_stream[k][b] > 0.1 else ".") for k in order)
        print("%7.2f ms  occ %5.2f  inflight %5.2f  %s" % (b * w / 1e3, occ[b], nrun[b], marks))
    print("mean occupancy estimate %.3f" % (sum(occ) / nb))

    # ---- time by kernel name
    agg = {}
    for r in rows:
        a = agg.setdefault(r["name"], [0, 0.0])
        a[0] += 1
        a[1] += r["d"]
    tot = sum(a[1] for a in agg.values())
    print("\nkernel time by name (this replay; durations are stretched when lanes overlap)")
    for name, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%-62s %6d %10.1f us %5.1f%%" % (name, n, d, 100 * d / tot))

    # ---- GEMM time by shape: with a single stream the graph runs its kernels in capture order, so the i-th gemm kernel of
    # the trace is the i-th ld_gemm_bf16 call of the capture
    gk = [r for r in rows if "gemm_bf16" in r["name"]]
    if args.lanes == 0 and gemm_log is not None and len(gk) == len(gemm_log):
        shp = {}
        for r, g in zip(gk, gemm_log):
            key = g[:7] + (g[7],)
            a = shp.setdefault(key, [0, 0.0])
            a[0] += 1
            a[1] += r["d"]
        print("\nGEMM time by (M, N, K, batches, a_mn, b_mn, split_k, caller): count, total us, us each, TFLOP/s")
        gtot = sum(a[1] for a in shp.values())
        for key, (n, d) in sorted(shp.items(), key=lambda kv: -kv[1][1])[:70]:
            M_, N_, K_, nb_ = key[:4]
            tf = 2.0 * M_ * N_ * K_ * nb_ * n / (d * 1e-6) / 1e12
            print("%-100s %5d %10.1f %8.1f %8.1f  %4.1f%%" % (str(key), n, d, d / n, tf, 100 * d / gtot))
        print("GEMM total %.1f us in %d launches" % (gtot, len(gk)))
    else:
        print("\n(GEMM shape table needs --lanes 0: %d gemm kernels in the trace, %s logged calls)" % (len(gk), None if gemm_log is None else len(gemm_log)))


if __name__ == "__main__":
    main()
