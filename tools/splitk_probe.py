"""Split-K sweep for the weight-gradient shapes of one step whose output is small and whose reduction is long (conv / DETR wgrads):
time per launch inside a CUDA graph, fp32 atomics into a pre-zeroed buffer (what functional._wgrad / Conv2dFn.backward issue)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from layoutdetr_b200 import kernels as K, engine as E


def graph_time(fn, reps=50):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (M, N, Kd) in [(256, 64, 65536), (64, 256, 65536), (64, 576, 65536), (512, 128, 16384), (128, 512, 16384), (128, 1152, 16384),
                   (256, 2304, 4096), (1024, 256, 4096), (256, 1024, 4096), (512, 4608, 1024), (256, 256, 1024), (768, 768, 32768)]:
    a = torch.randn((Kd, M), device="cuda").to(torch.bfloat16)
    b = torch.randn((Kd, N), device="cuda").to(torch.bfloat16)
    out = torch.zeros((M, N), dtype=torch.float32, device="cuda")
    cur = E.wgrad_split_k(M, N, Kd)
    row = []
    for sk in (1, 2, 4, 8, 16, 32, 64, 148):
        if sk > (Kd + 63) // 64:
            continue
        t = graph_time(lambda: K.gemm(M, N, Kd, K.Op(a, M, mn=True), K.Op(b, N, mn=True), K.Out(out, N), accumulate=2 if sk > 1 else 1, split_k=sk))
        row.append("sk%d %.1f" % (sk, t))
    print("M%5d N%5d K%6d  heuristic sk=%3d | %s us" % (M, N, Kd, cur, "  ".join(row)), flush=True)
