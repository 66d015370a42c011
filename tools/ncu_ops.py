"""ncu target for the kernels the north star names besides the plain GEMM: fused attention (BERT shape and DETR
cross-attention shape), the backbone conv path (im2col + tcgen05 GEMM with the FrozenBN/ReLU epilogue), LayerNorm
(+residual), the StyleGAN2 ops (bias_act, upfirdn2d, demod+bias+act), label-smoothed CE and the flat Adam step.
Three rounds; only the last one sits between cudaProfilerStart/Stop:

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r1_ops python tools/ncu_ops.py

Read back with `ncu -i gpurun_out/r1_ops.ncu-rep --page raw --csv` (tools/summarize_ncu_raw.py makes the table in profiles/).
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from layoutdetr_b200 import kernels as K, functional as F
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
# ResNet-50 layer1 3x3 conv at bs16: [16, 64, 64, 64] -> 64 channels, FrozenBN scale/shift + ReLU in the GEMM epilogue
conv_in = torch.randn((16 * 64 * 64, 64), device=dev).to(torch.bfloat16)
conv_w = torch.randn((64, 64, 3, 3), device=dev) * 0.05
bn_s = torch.rand(64, device=dev) + 0.5; bn_b = torch.randn(64, device=dev)
# layer3 3x3 conv: [16, 16, 16, 256] -> 256
conv3_in = torch.randn((16 * 16 * 16, 256), device=dev).to(torch.bfloat16)
conv3_w = torch.randn((256, 256, 3, 3), device=dev) * 0.02
bn3_s = torch.rand(256, device=dev) + 0.5; bn3_b = torch.randn(256, device=dev)
# LM-head CE rows: 16 sequences x 256 tokens over the 30524-way vocabulary (padded row pitch 30528)
logits = torch.randn((16 * 256, 30528), device=dev).to(torch.bfloat16)[:, :30524]
labels = torch.randint(0, 30524, (16 * 256,), device=dev)
dlog = logits                                                     # in place, as LMHeadCEFn does
# flat Adam over 16 M parameters
n = 16 * 1024 * 1024
p = torch.randn(n, device=dev); gr = torch.randn(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
p16 = torch.empty(n, dtype=torch.bfloat16, device=dev)


# LM text-decoder attention backward (128 sequences, causal): fused ld_attention_bwd
Bd = 128
qkv_d = torch.randn((Bd * T, 3 * H * d), device=dev).to(torch.bfloat16)
kmd = torch.zeros((Bd, T), dtype=torch.uint8, device=dev); kmd[:, 40:] = 1
lse_d = torch.empty((Bd * H, T), dtype=torch.float32, device=dev)
O_d = K.attention_fwd(qkv_d, 0, qkv_d, H * d, qkv_d, 2 * H * d, Bd, H, T, T, d, d ** -0.5, key_mask=kmd, causal=True, lse_out=lse_d)
dO_d = torch.randn_like(O_d)
dq_d = torch.empty_like(qkv_d)
Pd_d = torch.empty((Bd * H, T, T), dtype=torch.bfloat16, device=dev)
dS_d = torch.empty_like(Pd_d)

OPS = [
    ("attention_bert", lambda: K.attention_fwd(qkv, 0, qkv, H * d, qkv, 2 * H * d, B, H, T, T, d, d ** -0.5, key_mask=km)),
    ("attention_bert_dropout", lambda: K.attention_fwd(qkv, 0, qkv, H * d, qkv, 2 * H * d, B, H, T, T, d, d ** -0.5, key_mask=km,
                                                       dropout_p=0.1, rng_site=7)),
    ("attention_bert_bwd", lambda: K.attention_bwd(qkv_d, 0, qkv_d, H * d, qkv_d, 2 * H * d, O_d, dO_d, lse_d, dq_d, Pd_d, dS_d, Bd, H, T, T,
                                                   d, d ** -0.5, key_mask=kmd, causal=True)),
    ("attention_detr_cross", lambda: K.attention_fwd(q2, 0, kv2, 0, kv2, 256, 16, 8, 10, 64, 32, 32 ** -0.5)),
    ("layernorm_res", lambda: K.layernorm_fwd(x_ln, g, b, 1e-12, residual=res)),
    ("bias_act", lambda: ba.bias_act(img, bias, act="lrelu")),
    ("upsample2d", lambda: uf.upsample2d(rgb, f)),
    ("demod_bias_act", lambda: K.demod_bias_act_fwd(act, dco, bias, 16, 256 * 256, 32, K.ACT_LRELU, 2 ** 0.5)),
    ("conv3x3_layer1", lambda: F.conv2d(conv_in, conv_w, bn_s, bn_b, None, 16, 64, 64, 1, 1, K.ACT_RELU)),
    ("conv3x3_layer3", lambda: F.conv2d(conv3_in, conv3_w, bn3_s, bn3_b, None, 16, 16, 16, 1, 1, K.ACT_RELU)),
    ("cross_entropy", lambda: K.cross_entropy(logits, labels, label_smoothing=0.1, dlogits=dlog, grad_scale=1.0 / labels.numel())),
    ("adam_flat", lambda: K.adam_flat(p, gr, m, v, p16, 1e-5, 0.0, 0.99, 1e-8, 1)),
]


def one_round():
    with torch.no_grad():
        for name, fn in OPS:
            try:
                fn()
            except Exception as e:          # keep profiling the others
                print("ncu_ops: %s failed: %r" % (name, e), flush=True)


for i in range(3):
    if i == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    one_round()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
