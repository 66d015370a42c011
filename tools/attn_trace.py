"""Timeline of the persistent attention kernel (csrc/attention_fwd_sm100.cu) on the BERT shape: event clocks of CTA 0, per tile,
relative to the first event, in SM cycles.  Usage on a B200:  python tools/attn_trace.py > gpurun_out/attn_trace.txt"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from layoutdetr_b200 import kernels as K, _lib

B, H, T, d = 144, 4, 256, 192
torch.manual_seed(0)
qkv = torch.randn((B * T, 3 * H * d), device="cuda").to(torch.bfloat16)
km = torch.zeros((B, T), dtype=torch.uint8, device="cuda"); km[:, 40:] = 1
for _ in range(3):
    K.attention_fwd(qkv, 0, qkv, H * d, qkv, 2 * H * d, B, H, T, T, d, d ** -0.5, key_mask=km)
buf = torch.zeros((64, 16), dtype=torch.int64, device="cuda")
_lib.lib().ld_debug_attention_trace(ctypes.c_void_p(buf.data_ptr()))
K.attention_fwd(qkv, 0, qkv, H * d, qkv, 2 * H * d, B, H, T, T, d, d ** -0.5, key_mask=km)
torch.cuda.synchronize()
_lib.lib().ld_debug_attention_trace(ctypes.c_void_p(0))
t = buf.cpu()
names = ["qk_start", "qk_issued", "pv_start", "pv_issued", "s_seen", "max_done", "p_done", "o_seen", "epi_done", "staged_seen", "store_read", "q_issued"]
t0 = int(t[t > 0].min())
print("tile " + " ".join("%11s" % n for n in names))
for i in range(64):
    if int(t[i].max()) == 0:
        break
    print("%4d " % i + " ".join("%11s" % (int(t[i, e]) - t0 if int(t[i, e]) else "-") for e in range(len(names))))
