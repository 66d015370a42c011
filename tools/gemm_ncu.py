import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from layoutdetr_b200 import kernels as K
M, N, Kd = 16*9*256, 3072, 768
a = torch.randn((M, Kd), device="cuda").to(torch.bfloat16)
w = torch.randn((N, Kd), device="cuda").to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
o16 = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
for _ in range(2):
    K.linear(a, w, out=o16)
    K.linear(a, w, bias, act=K.ACT_GELU, out=o16)
torch.cuda.synchronize()
