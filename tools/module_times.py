"""Per-module GPU time of one training iteration (CUDA events around module calls, forward only + whole phases)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LAYOUTDETR_SYNTHETIC_TOKENIZER", "1")
import torch
import bench
from layoutdetr_b200.synthetic import make_inputs
from layoutdetr_b200.training import networks_detr as nd
from layoutdetr_b200.training.trainer import Trainer

dev = torch.device("cuda")
torch.manual_seed(0)
G = nd.Generator(**bench.G_KWARGS).to(dev); D = nd.Discriminator(**bench.D_KWARGS).to(dev)
tr = Trainer(G, D, dev, batch_size=16)
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in make_inputs(16, n_valid=8, seed=1).items()}
z = torch.randn((16, 9, 4), device=dev)
for _ in range(2): tr.iteration(b, z, z)
torch.cuda.synchronize()

times = {}
def wrap(mod, name):
    orig = mod.forward
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = orig(*a, **k); e1.record()
        times.setdefault(name, []).append((e0, e1))
        return out
    mod.forward = f
for m, n in [(G.backbone, "G.backbone"), (G.text_encoder, "G.text_encoder"), (G.transformer, "G.transformer"), (G.text_decoder, "G.text_decoder"), (G.fc_in, "G.fc_in"),
             (D.backbone, "D.backbone"), (D.text_encoder, "D.text_encoder"), (D.enc_transformer, "D.enc_transformer"), (D.text_decoder, "D.text_decoder"),
             (D.bg_decoder, "D.bg_decoder"), (D.enc_transformer_uncond, "D.enc_transformer_uncond")]:
    wrap(m, n)
# text_encoder is called through cls_features, wrap that too
for mod, n in [(G.text_encoder, "G.text_encoder.cls"), (D.text_encoder, "D.text_encoder.cls")]:
    orig = mod.cls_features
    def f(*a, _o=orig, _n=n, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = _o(*a, **k); e1.record(); times.setdefault(_n, []).append((e0, e1)); return out
    mod.cls_features = f

def phase_time(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
tG = phase_time(lambda: tr._phase("G", b, z))
tD = phase_time(lambda: tr._phase("D", b, z))
torch.cuda.synchronize()
print("phase G (fwd+bwd+adam) %.1f ms, phase D %.1f ms" % (tG, tD))
for n, evs in sorted(times.items()):
    ts = [a.elapsed_time(c) for a, c in evs]
    print("  %-28s fwd calls %d: %s ms" % (n, len(ts), ", ".join("%.2f" % t for t in ts)))
# forward-only vs backward split for phase G
with torch.no_grad():
    t = phase_time(lambda: G(z, b["bbox_class"], b["bbox_real"], b["bbox_text"], b["bbox_patch"], b["padding_mask"], b["background"], b["c"], reconst=True))
print("G forward(reconst) no-grad %.1f ms" % t)
with torch.no_grad():
    t = phase_time(lambda: D(b["bbox_real"], b["bbox_class"], b["bbox_text"], b["bbox_patch"], b["padding_mask"], b["background"], b["c"], reconst=True))
print("D forward(reconst) no-grad %.1f ms" % t)
