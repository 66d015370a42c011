"""Throughput of the evaluation sweep (BASELINE configs[4]: layout FID + overlap / alignment / layout-wise IoU / DocSim,
inference only) at 64 layouts per GPU per batch, next to the CPU oracle of the same sweep.  Prints one JSON line.

    python tools/eval_sweep_bench.py [--batch 64] [--batches 4] [--cpu-sample 4]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LAYOUTDETR_SYNTHETIC_TOKENIZER", "1")
import torch

import bench as B
from layoutdetr_b200.synthetic import make_inputs, SyntheticTokenizer
from layoutdetr_b200.training import networks_detr as nd
from layoutdetr_b200.training.networks_layoutnet import LayoutNet
from layoutdetr_b200.metrics import eval_sweep
from layoutdetr_b200 import _lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--cpu-sample", type=int, default=4)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    G = nd.Generator(**B.G_KWARGS).to(dev).eval().requires_grad_(False)
    net = LayoutNet(13).to(dev).eval().requires_grad_(False)
    host = [make_inputs(args.batch, n_valid=8, seed=50 + i) for i in range(args.batches)]
    for hb in host:
        for k, v in hb.items():
            if torch.is_tensor(v):
                hb[k] = v.pin_memory()
    to_dev = lambda hb: {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in hb.items()}
    eval_sweep.run_sweep(G, net, [to_dev(host[0])])                       # warm-up (tokeniser cache, weight shadows)
    torch.cuda.synchronize()
    _lib.launch_count_reset()
    t0 = time.perf_counter()
    res = eval_sweep.run_sweep(G, net, (to_dev(hb) for hb in host))      # H2D of every batch + D2H of the metric dict inside
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    n = args.batch * args.batches
    # CPU oracle of the same sweep on a bounded sample
    from oracle import layoutdetr_oracle as O
    from layoutdetr_b200.synthetic import synth_state_dict
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sdG = {k: v.detach().cpu() for k, v in G.state_dict().items()}
    sdL = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    inp = make_inputs(args.cpu_sample, n_valid=8, seed=50)
    tok = SyntheticTokenizer()
    t1 = time.perf_counter()
    with torch.no_grad():
        fake = O.generator_forward(sdG, tok, inp["z"], inp["bbox_class"], inp["bbox_text"], inp["padding_mask"], inp["background"])
        mask = ~inp["padding_mask"]
        O.layoutnet_extract_features(sdL, inp["bbox_real"], inp["bbox_class"], inp["padding_mask"])
        O.layoutnet_extract_features(sdL, fake, inp["bbox_class"], inp["padding_mask"])
        O.compute_overlap(fake, mask); O.compute_alignment(fake, mask); O.layoutwise_iou_docsim(inp["bbox_real"], fake, mask)
    cpu_sec = time.perf_counter() - t1
    print(json.dumps(dict(metric="eval-sweep layouts/sec (G_ema fwd + LayoutNet features + overlap/alignment/IoU/DocSim + FID), bs%d" % args.batch,
                          value=n / sec, unit="layouts/s", n_gpus=1, layouts=n, seconds=sec, gpu_launches=_lib.launch_count(),
                          gflop_per_layout=425.2, tflops=n / sec * 425.2e9 / 1e12, result=res,
                          cpu_baseline=dict(value=args.cpu_sample / cpu_sec, unit="layouts/s", cores=threads, kind="port",
                                            sample="%d layouts, oracle sweep (%.1f s)" % (args.cpu_sample, cpu_sec)))))


if __name__ == "__main__":
    main()
