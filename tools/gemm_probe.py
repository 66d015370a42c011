"""Micro-probe of gemm_bf16_kernel epilogue variants on BERT shapes (timing with CUDA events; ncu-friendly)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from layoutdetr_b200 import kernels as K

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]

M = 16*9*256
cases = []
for (N, Kd) in [(3072, 768), (2304, 768), (768, 768), (768, 3072)]:
    a = torch.randn((M, Kd), device="cuda").to(torch.bfloat16)
    w = torch.randn((N, Kd), device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    res = torch.randn((M, N), device="cuda").to(torch.bfloat16)
    o16 = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    o32 = torch.empty((M, N), dtype=torch.float32, device="cuda")
    fl = 2.0*M*N*Kd
    for name, fn in [
        ("plain bf16", lambda: K.linear(a, w, out=o16)),
        ("bias bf16", lambda: K.linear(a, w, bias, out=o16)),
        ("bias+gelu bf16", lambda: K.linear(a, w, bias, act=K.ACT_GELU, out=o16)),
        ("bias+relu bf16", lambda: K.linear(a, w, bias, act=K.ACT_RELU, out=o16)),
        ("bias+res f32out", lambda: K.linear(a, w, bias, residual=res, out=o32)),
        ("plain bn128", lambda: K.gemm(M, N, Kd, K.Op(a, Kd), K.Op(w, Kd), K.Out(o16, N), block_n=128)),
    ]:
        ms = timeit(fn)
        print("M%d N%d K%d %-18s %8.1f us  %7.1f TFLOP/s" % (M, N, Kd, name, ms*1e3, fl/(ms*1e-3)/1e12), flush=True)
# torch (cuBLAS) for context only
a = torch.randn((M, 768), device="cuda").to(torch.bfloat16); w = torch.randn((3072, 768), device="cuda").to(torch.bfloat16)
ms = timeit(lambda: a @ w.t())
print("cuBLAS context M%d N3072 K768: %.1f us %.1f TFLOP/s" % (M, ms*1e3, 2.0*M*3072*768/(ms*1e-3)/1e12))
