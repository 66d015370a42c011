"""GPU-only time of each sub-module's forward(+backward), measured by capturing it into a CUDA graph (no CPU launch gaps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LAYOUTDETR_SYNTHETIC_TOKENIZER", "1")
import torch
import bench
from layoutdetr_b200 import functional as Fn, kernels as K
from layoutdetr_b200.synthetic import make_inputs
from layoutdetr_b200.training import networks_detr as nd

dev = torch.device("cuda", 0)
torch.manual_seed(0)
G = nd.Generator(**bench.G_KWARGS).to(dev); D = nd.Discriminator(**bench.D_KWARGS).to(dev)
B = 16
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in make_inputs(B, n_valid=8, seed=1).items()}

def graph_time(fn, reps=5):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def clear(m):
    for p in m.parameters(): p.grad = None

def fb(mod, fwd):
    def f():
        clear(mod)
        out = fwd()
        outs = out if isinstance(out, (tuple, list)) else [out]
        loss = sum(o.float().sum() for o in outs if torch.is_tensor(o) and o.requires_grad)
        loss.backward()
    return f

G.requires_grad_(True); G.text_encoder.requires_grad_(False); D.requires_grad_(True); D.text_encoder.requires_grad_(False)
text = G._front()(b["bbox_text"], dev)
ids, mask = text["ids"], text["mask"]
res = {}
with torch.no_grad():
    res["text_encoder fwd (frozen, dense T=256)"] = graph_time(lambda: G.text_encoder.cls_features(ids, mask))
res["backbone fwd"] = None
with torch.no_grad():
    res["backbone fwd"] = graph_time(lambda: G.backbone(b["background"]))
res["backbone fwd+bwd"] = graph_time(fb(G.backbone, lambda: G.backbone(b["background"])[0]))
feat, pos, h, w = G.backbone(b["background"]); feat = feat.detach()
src = Fn.conv2d(feat, G.input_proj.weight, None, G.input_proj.bias, None, B, h, w, 1, 0, K.ACT_NONE).detach()
x = torch.randn((B * 9, 256), device=dev).to(torch.bfloat16).requires_grad_(True)
with torch.no_grad():
    res["G.transformer fwd"] = graph_time(lambda: G.transformer(src, pos, x, b["padding_mask"], B, h * w, 9)[0])
res["G.transformer fwd+bwd"] = graph_time(fb(G.transformer, lambda: G.transformer(src, pos, x, b["padding_mask"], B, h * w, 9)[0]))
res["D.enc_transformer fwd+bwd"] = graph_time(fb(D.enc_transformer, lambda: D.enc_transformer(src, pos, x, b["padding_mask"], B, h * w, 9)[0]))
res["D.enc_transformer_uncond fwd+bwd"] = graph_time(fb(D.enc_transformer_uncond, lambda: D.enc_transformer_uncond(x, B, 9, b["padding_mask"])))
from layoutdetr_b200.training.detr_transformer import TransformerEncoderStack
res["D.dec_transformer fwd+bwd"] = graph_time(fb(D.dec_transformer, lambda: TransformerEncoderStack.run(D.dec_transformer, x, B, 9, b["padding_mask"])))
x0 = torch.randn((B, 256), device=dev).to(torch.bfloat16).requires_grad_(True)
with torch.no_grad():
    res["bg_decoder fwd"] = graph_time(lambda: D.bg_decoder(x0))
res["bg_decoder fwd+bwd"] = graph_time(fb(D.bg_decoder, lambda: D.bg_decoder(x0)))
vidx, vcpu = nd.valid_index(b["padding_mask"])
def dec():
    clear(G.text_decoder)
    l = nd._decode_text_loss(G, text, vidx, vcpu, G.tokenizer.bos_token_id, G.tokenizer.pad_token_id)
    l.backward()
nd._decode_text_loss(G, text, vidx, vcpu, G.tokenizer.bos_token_id, G.tokenizer.pad_token_id)
res["text_decoder+LM head fwd+bwd (dense)"] = graph_time(dec)
for k, v in res.items():
    print("%-45s %8.2f ms" % (k, v))
