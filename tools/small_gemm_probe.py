"""Per-launch cost of small GEMMs inside a CUDA graph (back-to-back, no CPU gaps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from layoutdetr_b200 import kernels as K

def graph_time(fn, reps=200):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); 
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for (M, N, Kd) in [(144, 256, 256), (144, 2048, 256), (1024, 512, 256), (1024, 2048, 256), (1024, 256, 2048), (16384, 64, 64), (65536, 64, 576), (4096, 512, 1152), (36864, 768, 768)]:
    a = torch.randn((M, Kd), device="cuda").to(torch.bfloat16); w = torch.randn((N, Kd), device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda"); o = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    t = graph_time(lambda: K.linear(a, w, bias, act=K.ACT_RELU, out=o))
    tc = graph_time(lambda: torch.addmm(bias.to(torch.bfloat16), a, w.t(), out=o))
    print("M%6d N%5d K%5d  ours %7.1f us   cuBLAS(context) %7.1f us   %.1f GF" % (M, N, Kd, t, tc, 2.0*M*N*Kd/1e9), flush=True)
# elementwise floor
x = torch.randn(144*256, device="cuda").to(torch.bfloat16)
t = graph_time(lambda: K.act_fwd(x, K.ACT_RELU))
print("tiny elementwise kernel: %.1f us" % t)
