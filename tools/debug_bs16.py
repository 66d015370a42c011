"""Diagnostic: Gmain / Dmain gradients at the benched batch size against the reference golden (tests/golden/loss_b16_v8.pt) under
different execution settings (lane level, fused attention on / off, CTA-pair GEMM via LD_GEMM_2SM in the environment).
    python tools/debug_bs16.py --lanes 0 --fused 1 [--phase Gmain]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("LAYOUTDETR_SYNTHETIC_WEIGHTS", "1")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lanes", type=int, default=3)
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--phase", default="Gmain")
    ap.add_argument("--golden", default="loss_b16_v8")
    args = ap.parse_args()
    import torch
    from helpers import build, golden
    from layoutdetr_b200 import functional as Fn
    from layoutdetr_b200.lanes import LANES
    from layoutdetr_b200.synthetic import make_inputs
    from test_train_gpu import _phase_grads, _to_dev
    LANES.configure(level=args.lanes)
    Fn.FUSED_ATTENTION = bool(args.fused)
    g = golden(args.golden + ".pt")
    G, D = build("G").cuda(), build("D").cuda()
    inp = _to_dev(make_inputs(g["batch"], n_valid=g["n_valid"], seed=g["inputs_seed"]))
    grads, _ = _phase_grads(args.phase, G, D, inp)
    ref = g["grads"][args.phase]
    rows = sorted(((abs(float(grads[k].float().norm()) - n) / n, k) for k, n in ref["norms"].items() if n > 1e-7 and k in grads), reverse=True)
    cos = sorted((float(torch.dot(grads[k].float().cpu().reshape(-1), t.reshape(-1)) / (grads[k].float().norm().cpu() * t.norm() + 1e-20)), k)
                 for k, t in ref["small"].items() if float(t.norm()) > 1e-7 and k in grads)
    tag = "lanes=%d fused=%d 2sm=%s %s" % (args.lanes, args.fused, os.environ.get("LD_GEMM_2SM", "1"), args.phase)
    print(tag, "| worst norm err %.3f (%s) | within 10%%: %.3f | worst cos %.4f (%s), %.4f (%s)" % (
        rows[0][0], rows[0][1], sum(1 for r in rows if r[0] < 0.1) / len(rows), cos[0][0], cos[0][1], cos[1][0], cos[1][1]), flush=True)


if __name__ == "__main__":
    main()
