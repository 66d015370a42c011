"""compute-sanitizer --tool initcheck target: one Dreal pass (forward + backward) of a small Discriminator."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["LAYOUTDETR_SYNTHETIC_TOKENIZER"] = "1"
import torch
from helpers import D_KWARGS
from layoutdetr_b200.lanes import LANES
from layoutdetr_b200.training import networks_stylegan2 as sg
from layoutdetr_b200 import functional as Fn
LANES.configure(level=0)
torch.manual_seed(0)
dec = sg.Decoder(z_dim=256, w_dim=512, channel_max=512, channel_base=8192, img_channels=3, img_resolution=256, use_noise=False,
                 num_fp16_res=0, conv_clamp=None, fused_modconv_default=False).cuda()
dec.requires_grad_(True)
x0 = torch.randn(2, 256, device="cuda").to(torch.bfloat16).requires_grad_(True)
img = dec(x0)
tgt = torch.randn_like(img)
loss = torch.nn.functional.mse_loss(img, tgt)
loss.backward()
torch.cuda.synchronize()
print("loss", float(loss), "dx0 norm", float(x0.grad.float().norm()))
