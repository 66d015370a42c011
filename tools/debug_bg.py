"""bg_decoder alone: forward + backward twice (clean allocator vs NaN-poisoned free blocks); which gradients differ?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LAYOUTDETR_SYNTHETIC_TOKENIZER"] = "1"
import torch
from layoutdetr_b200 import engine
from layoutdetr_b200.lanes import LANES
from layoutdetr_b200.training import networks_stylegan2 as sg
LANES.configure(level=0)


def run(poison, fill=float("nan")):
    engine.clear_cache()
    torch.manual_seed(0)
    dec = sg.Decoder(z_dim=256, w_dim=512, channel_max=512, channel_base=8192, img_channels=3, img_resolution=256, use_noise=False,
                     num_fp16_res=0, conv_clamp=None, fused_modconv_default=False).cuda()
    dec.requires_grad_(True)
    x0 = torch.randn(2, 256, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)).to(torch.bfloat16).requires_grad_(True)
    tgt = torch.randn(2, 3, 256, 256, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    if poison:
        torch.cuda.empty_cache()
        junk = [torch.full((n,), fill, device="cuda") for n in (1 << 28, 1 << 26, 1 << 24, 1 << 22, 1 << 20, 1 << 18, 1 << 16, 1 << 14)]
        del junk
    img = dec(x0)
    loss = torch.nn.functional.mse_loss(img, tgt)
    loss.backward()
    torch.cuda.synchronize()
    g = {n: p.grad.clone() for n, p in dec.named_parameters() if p.grad is not None}
    g["__x0"] = x0.grad.float().clone()
    g["__img"] = img.detach().clone()
    return g


a = run(False)
for fill in (float("nan"), 1e30, 3.0):
    b = run(True, fill)
    rows = sorted(((float((b[n].float() - a[n].float()).norm() / (a[n].float().norm() + 1e-20)), n) for n in a), reverse=True)
    print("fill", fill, "| differing tensors:", sum(1 for r in rows if r[0] > 1e-6 or r[0] != r[0]), "of", len(rows))
    for r in rows[:24]:
        if r[0] > 1e-6 or r[0] != r[0]:
            print("   %.3e  %s" % r)
