import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["LAYOUTDETR_SYNTHETIC_TOKENIZER"] = "1"
import torch
from helpers import build, golden
from layoutdetr_b200 import functional as Fn, kernels as K
from layoutdetr_b200.synthetic import make_inputs
G = build("G").cuda()
g = golden("model_b2_v8.pt")
inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in make_inputs(g["batch"], n_valid=g["n_valid"], seed=g["inputs_seed"]).items()}
outs = {}
calls = [0]
orig = K.layernorm_fwd
def counted(*a, **kw):
    if kw.get("residual") is not None:
        calls[0] += 1
    return orig(*a, **kw)
K.layernorm_fwd = counted
for flag in (False, True):
    Fn.LN_BF16_DENSE_NOGRAD = flag
    calls[0] = 0
    with torch.no_grad():
        text = G._front()(inp["bbox_text"], torch.device("cuda"))
        feat = G.text_encoder.cls_features(text["ids"], text["mask"]).float().clone()
        out = G(inp["z"], inp["bbox_class"], inp["bbox_real"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"], inp["background"], inp["c"], reconst=True)
    outs[flag] = (feat, out[0].float().clone())
    print("flag", flag, "residual-LN calls", calls[0], "bbox err vs golden", float((out[0].float().cpu() - g["G"]["bbox_fake"]).abs().max()))
print("CLS feature max abs diff on/off:", float((outs[True][0] - outs[False][0]).abs().max()), "scale", float(outs[False][0].abs().max()))
print("bbox max abs diff on/off:", float((outs[True][1] - outs[False][1]).abs().max()))
