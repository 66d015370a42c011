"""ncu target for the non-GEMM kernels the north star names: fused attention (BERT shape and DETR cross-attention shape),
LayerNorm (+residual), StyleGAN2 ops (bias_act, upfirdn2d, demod+bias+act), im2col.  Three rounds; profile the last:

  ncu --set full --clock-control none --import-source on -k regex:'attention_fwd|layernorm_fwd|bias_act_kernel|upfirdn2d|demod_bias_act_fwd|im2col_vec8' \
      -s <2 x launches per round> -c <launches per round> -o gpurun_out/ops python tools/ncu_ops.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from layoutdetr_b200 import kernels as K
from layoutdetr_b200.torch_utils.ops import bias_act as ba, upfirdn2d as uf

dev = "cuda"
torch.manual_seed(0)
B, H, T, d = 144, 4, 256, 192                                     # one text-encoder call at bs16
qkv = torch.randn((B * T, 3 * H * d), device=dev).to(torch.bfloat16)
km = torch.zeros((B, T), dtype=torch.uint8, device=dev); km[:, 40:] = 1
q2 = torch.randn((16 * 10, 256), device=dev).to(torch.bfloat16)   # DETR decoder cross-attention: 10 queries x 64 image tokens
kv2 = torch.randn((16 * 64, 512), device=dev).to(torch.bfloat16)
x_ln = torch.randn((B * T, 768), device=dev).to(torch.bfloat16)
res = torch.randn((B * T, 768), device=dev).to(torch.bfloat16)
g = torch.ones(768, device=dev); b = torch.zeros(768, device=dev)
img = torch.randn((16, 32, 256, 256), device=dev)                 # bias_act on the largest bg_decoder activation (fp32 reference op API)
bias = torch.randn(32, device=dev)
f = uf.setup_filter([1, 3, 3, 1]).to(dev)
rgb = torch.randn((16, 3, 128, 128), device=dev)
act = torch.randn((16 * 256 * 256, 32), device=dev).to(torch.bfloat16)
dco = torch.rand((16, 32), device=dev) + 0.5
conv_in = torch.randn((16 * 64 * 64, 64), device=dev).to(torch.bfloat16)
for _ in range(3):
    K.attention_fwd(qkv, 0, qkv, H * d, qkv, 2 * H * d, B, H, T, T, d, d ** -0.5, key_mask=km)
    K.attention_fwd(q2, 0, kv2, 0, kv2, 256, 16, 8, 10, 64, 32, 32 ** -0.5)
    K.layernorm_fwd(x_ln, g, b, 1e-12, residual=res)
    ba.bias_act(img, bias, act="lrelu")
    uf.upsample2d(rgb, f)
    K.demod_bias_act_fwd(act, dco, bias, 16, 256 * 256, 32, K.ACT_LRELU, 2 ** 0.5)
    K.im2col(conv_in, 16, 64, 64, 64, 3, 3, 1, 1)
torch.cuda.synchronize()
