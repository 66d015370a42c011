"""ncu target: the three epilogue variants of gemm_bf16_kernel on the BERT K=768 shapes, 2 warm-up rounds then one round:
   ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 6 -c 3 -o gpurun_out/prof python tools/gemm_ncu.py
   launch order per round: [bias bf16 N=3072] [bias+GELU bf16 N=3072] [bias+residual fp32 N=768]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from layoutdetr_b200 import kernels as K
M, N, Kd = 16*9*256, 3072, 768
a = torch.randn((M, Kd), device="cuda").to(torch.bfloat16)
w = torch.randn((N, Kd), device="cuda").to(torch.bfloat16)
w2 = torch.randn((768, Kd), device="cuda").to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
res = torch.randn((M, 768), device="cuda").to(torch.bfloat16)
o16 = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
o32 = torch.empty((M, 768), dtype=torch.float32, device="cuda")
for _ in range(3):
    K.linear(a, w, bias, out=o16)
    K.linear(a, w, bias, act=K.ACT_GELU, out=o16)
    K.linear(a, w2, bias[:768].contiguous(), residual=res, out=o32)
torch.cuda.synchronize()
