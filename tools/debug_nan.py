import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import build
from layoutdetr_b200 import functional as Fn
from layoutdetr_b200.synthetic import make_inputs
from layoutdetr_b200.training.loss import StyleGAN2Loss

counter = [0]
def wrap(name):
    orig = getattr(Fn, name)
    def f(*a, **k):
        out = orig(*a, **k)
        if torch.is_tensor(out) and out.requires_grad:
            idx = counter[0]; counter[0] += 1
            shape = tuple(out.shape)
            def hook(g, idx=idx, shape=shape):
                bad = (~torch.isfinite(g)).sum().item()
                if bad:
                    print("non-finite grad arriving at output #%d of %s shape %s: %d bad, absmax finite %.3e" % (idx, name, shape, bad, float(torch.nan_to_num(g.float(), nan=0, posinf=0, neginf=0).abs().max())))
            out.register_hook(hook)
        return out
    setattr(Fn, name, f)
for n in ["linear", "linear_ln", "layernorm", "attention", "conv2d", "maxpool3s2", "add_bcast", "fused_linear", "linear_f32", "lm_head_ce", "cross_entropy", "to_bf16_padded", "to_f32"]:
    wrap(n)

G = build("G").cuda(); D = build("D").cuda()
inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in make_inputs(2, n_valid=8, seed=2).items()}
loss = StyleGAN2Loss(device=torch.device("cuda"), G=G, D=D)
for m in (G, D):
    m.requires_grad_(False)
G.requires_grad_(True); G.text_encoder.requires_grad_(False)
loss.accumulate_gradients(phase="Gmain", bbox_real=inp["bbox_real"], bbox_class=inp["bbox_class"], bbox_text=inp["bbox_text"],
                          bbox_patch=inp["bbox_patch"], padding_mask=inp["padding_mask"], background=inp["background"],
                          real_c=inp["c"], gen_z=inp["z"], gen_c=inp["c"], gain=1.0, cur_nimg=0)
torch.cuda.synchronize()
print("total wrapped outputs", counter[0])
bad = [(k, int((~torch.isfinite(p.grad)).sum()), p.grad.numel()) for k, p in G.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
print("params with non-finite grads:", len(bad))
for b in bad[:60]:
    print("  ", b)
