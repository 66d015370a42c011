"""Run-to-run determinism of the flat gradient buffers (single stream): which parameters differ between identical runs?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["LAYOUTDETR_SYNTHETIC_TOKENIZER"] = "1"
import torch
from helpers import G_KWARGS, D_KWARGS
from layoutdetr_b200 import engine
from layoutdetr_b200.lanes import LANES
from layoutdetr_b200.synthetic import make_inputs
from layoutdetr_b200.training import networks_detr as nd
from layoutdetr_b200.training.trainer import Trainer


def run(level, dry=0, poison=False, loss_kwargs=None):
    LANES.configure(level=level, dry=dry)
    engine.clear_cache()
    torch.manual_seed(0)
    G = nd.Generator(**dict(G_KWARGS, bert_num_encoder_layers=2, max_text_length=64)).cuda()
    D = nd.Discriminator(**dict(D_KWARGS, bert_num_encoder_layers=2, max_text_length=64)).cuda()
    tr = Trainer(G, D, torch.device("cuda"), batch_size=2, lr=0.0, loss_kwargs=loss_kwargs)
    hb = make_inputs(2, n_valid=8, seed=5)
    zs = [torch.randn((2, 9, 4), device="cuda", generator=torch.Generator(device="cuda").manual_seed(i)) for i in range(2)]
    dev_b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()}
    if poison:      # fill the allocator's free blocks with NaN patterns so reads of uninitialised memory show up
        junk = [torch.full((n,), float("nan"), device="cuda") for n in (1 << 28, 1 << 26, 1 << 24, 1 << 22, 1 << 20)]
        del junk
    for _ in range(3):
        tr.iteration(dev_b, zs[0], zs[1])
    torch.cuda.synchronize()
    out = {"loss": {ph + "/" + k: float(v.float().mean()) for ph in ("Gmain", "Dmain") for k, v in tr.loss.last[ph].items()}}
    for tag in ("G", "D"):
        f = tr.flat[tag]
        out[tag] = {n: f.g[o:o + p.numel()].clone() for n, p, o in zip(f.names, f.params, f.offsets)}
    return out


def compare(a, b, tag, top=12):
    rows = []
    for n in a[tag]:
        x, y = a[tag][n], b[tag][n]
        d = float((x - y).norm())
        rows.append((d / (float(y.norm()) + 1e-20), d, float(y.norm()), n))
    rows.sort(reverse=True)
    tot = sum(r[1] ** 2 for r in rows) ** 0.5 / (sum(r[2] ** 2 for r in rows) ** 0.5 + 1e-20)
    nbad = sum(1 for r in rows if r[0] > 1e-5)
    print("%s: overall rel L2 %.3e, params with rel diff > 1e-5: %d of %d" % (tag, tot, nbad, len(rows)))
    for r in rows[:top]:
        if r[0] > 1e-6:
            print("    %.3e  |d| %.3e  |g| %.3e  %s" % r)
    groups = {}
    for r in rows:
        g = groups.setdefault(r[3].split(".")[0], [0, 0])
        g[0] += 1
        g[1] += 1 if r[0] > 1e-5 else 0
    print("    affected / total per module:", {k: "%d/%d" % (v[1], v[0]) for k, v in groups.items() if v[1]})
    bgl = [(r[3], "%.2e" % r[0]) for r in rows if r[0] > 1e-5 and r[3].startswith("bg_decoder")]
    if bgl:
        print("    bg_decoder params affected:", bgl)
    nan = [n for n in a[tag] if not torch.isfinite(a[tag][n]).all()]
    if nan:
        print("    non-finite gradients:", nan[:10])


TRACE = None


def instrument():
    """Record a checksum of every tensor each hand-written autograd Function returns from backward (in execution order)."""
    from layoutdetr_b200 import functional as Fn
    for name in dir(Fn):
        cls = getattr(Fn, name)
        if isinstance(cls, type) and issubclass(cls, torch.autograd.Function) and cls is not torch.autograd.Function:
            orig = cls.backward

            def make(orig, name):
                def wrapped(ctx, *grads):
                    ins = [float(g.double().abs().sum()) if torch.is_tensor(g) else None for g in grads]
                    out = orig(ctx, *grads)
                    outs = out if isinstance(out, tuple) else (out,)
                    if TRACE is not None:
                        TRACE.append((name, ins, [float(t.double().abs().sum()) if torch.is_tensor(t) else None for t in outs],
                                      [tuple(t.shape) for t in outs if torch.is_tensor(t)]))
                    return out
                return staticmethod(wrapped)
            cls.backward = make(orig, name)


if __name__ == "__main__":
    instrument()
    traces = []
    for poison in (False, True):
        TRACE = []
        run(0, poison=poison)
        traces.append(TRACE)
    a, b = traces
    print("backward calls:", len(a), len(b))
    shown = 0
    for i, (x, y) in enumerate(zip(a, b)):
        if x[0] != y[0]:
            print("order differs at", i, x[0], y[0]); break
        if x[1] != y[1] or x[2] != y[2]:
            print("#%d %s shapes %s\n    in  clean %s\n    in  poisn %s\n    out clean %s\n    out poisn %s" % (i, x[0], x[3], x[1], y[1], x[2], y[2]))
            shown += 1
            if shown >= 6:
                break
