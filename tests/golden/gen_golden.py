"""Generate tests/golden/*.pt from the UNMODIFIED reference (build container only).

    python tests/golden/gen_golden.py [--skip-loss]

Imports /root/reference through oracle/ref_shim.py, overwrites every parameter/buffer with
layoutdetr_b200.synthetic.synth_tensor(name, shape) (a pure function of the state_dict key), runs the
reference on layoutdetr_b200.synthetic.make_inputs(...) in eval mode on CPU/fp32 and stores the (small)
outputs.  The GPU box has no reference: tests rebuild the same weights and inputs from the same pure
functions and compare against these files.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from layoutdetr_b200.synthetic import make_inputs, make_ragged_inputs, synth_state_dict  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _hook_outputs(module, store, name):
    def hook(_m, _inp, out):
        t = out[0] if isinstance(out, (tuple, list)) else out
        if isinstance(t, dict):
            t = list(t.values())[-1]
        if hasattr(t, "last_hidden_state"):
            t = t.last_hidden_state
        store[name] = t.detach().clone()
    return module.register_forward_hook(hook)


def gen_ops():
    """bias_act / upfirdn2d reference implementations (impl='ref') on small seeded tensors."""
    from torch_utils.ops import bias_act, upfirdn2d
    g = torch.Generator().manual_seed(123)
    out = {"bias_act": [], "upfirdn2d": []}
    for act in ["linear", "relu", "lrelu", "tanh", "sigmoid", "elu", "selu", "softplus", "swish"]:
        for (shape, dim, gain, clamp) in [((3, 5, 7, 6), 1, None, None), ((4, 16), 1, 0.7, 0.5)]:
            x = torch.randn(shape, generator=g) * 2
            b = torch.randn(shape[dim], generator=g)
            xr = x.clone().requires_grad_(True)
            br = b.clone().requires_grad_(True)
            y = bias_act.bias_act(xr, br, dim=dim, act=act, gain=gain, clamp=clamp, impl="ref")
            dy = torch.randn(shape, generator=g)
            dx, db = torch.autograd.grad(y, [xr, br], dy)
            out["bias_act"].append(dict(act=act, dim=dim, gain=gain, clamp=clamp, x=x, b=b, y=y.detach(), dy=dy, dx=dx, db=db))
    f4 = upfirdn2d.setup_filter([1, 3, 3, 1])
    f3 = upfirdn2d.setup_filter([1, 2, 1])
    fsep = torch.tensor([0.25, 0.5, 0.25, 0.1, 0.3])
    f2d = fsep.ger(fsep)
    cases = [
        dict(f=f4, up=1, down=1, padding=[1, 1, 1, 1], flip_filter=False, gain=4.0, shape=(2, 3, 9, 9)),     # post conv-transpose FIR
        dict(f=f4, up=2, down=1, padding=[2, 1, 2, 1], flip_filter=False, gain=4.0, shape=(2, 3, 8, 8)),     # upsample2d
        dict(f=f4, up=1, down=2, padding=[1, 1, 1, 1], flip_filter=False, gain=1.0, shape=(1, 4, 16, 12)),   # downsample2d
        dict(f=f3, up=2, down=3, padding=[3, 0, 1, 2], flip_filter=True, gain=0.5, shape=(2, 2, 7, 10)),
        dict(f=f2d, up=3, down=2, padding=[2, 2, 3, 1], flip_filter=False, gain=1.5, shape=(1, 3, 6, 5)),
        dict(f=f2d, up=1, down=1, padding=[-1, 3, 2, -1], flip_filter=True, gain=1.0, shape=(1, 2, 9, 9)),   # crop
    ]
    for c in cases:
        x = torch.randn(c["shape"], generator=g)
        xr = x.clone().requires_grad_(True)
        y = upfirdn2d.upfirdn2d(xr, c["f"], up=c["up"], down=c["down"], padding=c["padding"], flip_filter=c["flip_filter"],
                                gain=c["gain"], impl="ref")
        dy = torch.randn(y.shape, generator=g)
        (dx,) = torch.autograd.grad(y, [xr], dy)
        out["upfirdn2d"].append(dict(x=x, y=y.detach(), dy=dy, dx=dx, **{k: v for k, v in c.items() if k != "shape"}))
    torch.save(out, os.path.join(GOLD, "ops_ref.pt"))
    print("ops goldens:", len(out["bias_act"]), "bias_act,", len(out["upfirdn2d"]), "upfirdn2d")


def gen_hungarian():
    import numpy as np
    from scipy.optimize import linear_sum_assignment
    rng = np.random.RandomState(7)
    cases = []
    for n in range(1, 10):
        for rep in range(6):
            m = rng.rand(n, n)
            if rep == 3:
                m = np.zeros((n, n))
            if rep == 4:
                m = np.round(m * 3) / 3          # ties
            if rep == 5:
                m = (rng.rand(n, n) > 0.6).astype(np.float64) * rng.rand(n, n)
            r, c = linear_sum_assignment(m, maximize=True)
            cases.append(dict(cost=torch.from_numpy(m.copy()), row=torch.from_numpy(r), col=torch.from_numpy(c)))
    for (nr, nc) in [(3, 7), (7, 3), (9, 5), (1, 9)]:
        m = rng.rand(nr, nc)
        r, c = linear_sum_assignment(m, maximize=True)
        cases.append(dict(cost=torch.from_numpy(m.copy()), row=torch.from_numpy(r), col=torch.from_numpy(c)))
    torch.save(cases, os.path.join(GOLD, "hungarian_scipy.pt"))
    print("hungarian goldens:", len(cases))


def gen_vit():
    """Reference `training/networks_vit.py: VisionTransformer.forward(x, mask)` (ViT-B/16 on nn.TransformerEncoder, :139-221) with
    synthetic weights: a 64 x 96 image (4 x 6 patches, two of them masked out) and a 256 x 256 one (257 tokens)."""
    import training.networks_vit as nv
    out = {}
    for name, (h, w), seed in (("small", (64, 96), 5), ("bg256", (256, 256), 6)):
        torch.manual_seed(0)
        m = nv.VisionTransformer(img_height=h, img_width=w).eval()
        synth_state_dict(m)
        g = torch.Generator().manual_seed(seed)
        x = torch.randn((2, 3, h, w), generator=g)
        mask = torch.ones((2, 1, h, w))
        mask[1, :, :16, 16:48] = 0                       # sample 1: patches (0, 1) and (0, 2) carry no image content
        with torch.no_grad():
            y = m(x, mask)
        out[name] = dict(size=(h, w), seed=seed, y=y.clone(), y_nomask=m(x, torch.ones_like(mask)).detach().clone())
    torch.save(out, os.path.join(GOLD, "vit_ref.pt"))
    print("vit goldens:", {k: tuple(v["y"].shape) for k, v in out.items()})


def gen_maxiou():
    """Reference `compute_maximum_iou` (metrics/metric_layoutnet.py:140-150, scipy + a multiprocessing pool) on two seeded sets of
    layouts whose label multisets repeat, incl. groups of different sizes on the two sides (the reshape(N, M) quirk)."""
    import numpy as np
    import metrics.metric_layoutnet as mln          # the reference's (ref_shim.load() put the checkout first on sys.path)
    rng = np.random.RandomState(5)
    label_sets = [[0, 0, 1], [1, 2, 2, 2], [3], [0, 1, 2, 3, 4], [5, 5]]

    def layouts(counts):
        out = []
        for ls, k in zip(label_sets, counts):
            for _ in range(k):
                n = len(ls)
                b = np.stack([rng.uniform(0.2, 0.8, n), rng.uniform(0.2, 0.8, n), rng.uniform(0.1, 0.6, n), rng.uniform(0.05, 0.4, n)], 1)
                perm = rng.permutation(n)
                out.append((b[perm], np.array(ls)[perm]))
        return out

    l1, l2 = layouts([3, 2, 4, 1, 0]), layouts([2, 2, 4, 3, 2])
    score = mln.compute_maximum_iou(l1, l2, n_jobs=1)
    pack = lambda ls: [(torch.from_numpy(b.copy()), torch.from_numpy(l.copy())) for b, l in ls]
    torch.save(dict(layouts_1=pack(l1), layouts_2=pack(l2), score=float(score)), os.path.join(GOLD, "maxiou_ref.pt"))
    print("max-IoU golden:", score)


def gen_eval():
    """Evaluation-sweep pieces of the reference: LayoutNet.extract_features (FID features, synthetic weights), per-layout
    IoU / DocSim, overlap / alignment, FeatureStats moments + the layout-FID formula."""
    import numpy as np
    import scipy.linalg
    from training.networks_layoutnet import LayoutNet
    from metrics import metric_layoutnet as ml
    from metrics.metric_utils_layout import FeatureStats
    net = LayoutNet(13).eval()
    synth_state_dict(net)
    g = torch.Generator().manual_seed(11)
    B, N = 640, 9                                            # > 256 layouts so that the feature covariances have full rank
    scale, shift = torch.tensor([0.6, 0.6, 0.5, 0.3]), torch.tensor([0.2, 0.2, 0.05, 0.03])
    bbox_real = torch.rand((B, N, 4), generator=g) * scale + shift
    bbox_fake = (bbox_real + 0.08 * torch.randn((B, N, 4), generator=g)).clamp(0.01, 0.99)
    label = torch.randint(0, 13, (B, N), generator=g)
    n_valid = torch.randint(1, N + 1, (B,), generator=g)
    mask = torch.arange(N)[None, :] < n_valid[:, None]
    bbox_real = bbox_real * mask[..., None]
    with torch.no_grad():
        f_real = net.extract_features(bbox_real, label.clone(), ~mask)
        f_fake = net.extract_features(bbox_fake, label.clone(), ~mask)
        overlap = ml.compute_overlap(bbox_fake, mask)
        alignment = ml.compute_alignment(bbox_fake, mask)
    iou, docsim = [], []
    for j in range(B):
        m = mask[j].numpy()
        br, bf, l = bbox_real[j].numpy()[m], bbox_fake[j].numpy()[m], label[j].numpy()[m]
        iou.append(ml.compute_iou_for_layout((br, l), (bf, l)))
        docsim.append(ml.compute_docsim_for_layout((br, l), (bf, l)))
    st_r, st_f = FeatureStats(capture_mean_cov=True), FeatureStats(capture_mean_cov=True)
    st_r.append_torch(f_real)
    st_f.append_torch(f_fake)
    mu_r, s_r = st_r.get_mean_cov()
    mu_f, s_f = st_f.get_mean_cov()
    m = np.square(mu_f - mu_r).sum()                          # metrics/layout_frechet_inception_distance.py:36-39
    s = scipy.linalg.sqrtm(np.dot(s_f, s_r))                  # scipy >= 1.16 dropped `disp`; 1.6.3 returned (sqrtm, errest) with disp=False
    fid = float(np.real(m + np.trace(s_f + s_r - s * 2)))
    out = dict(bbox_real=bbox_real, bbox_fake=bbox_fake, label=label, mask=mask, f_real=f_real[:32].clone(), f_fake=f_fake[:32].clone(),
               overlap=overlap, alignment=alignment, iou=torch.tensor(iou, dtype=torch.float64),
               docsim=torch.tensor(docsim, dtype=torch.float64), fid=fid,
               layoutnet_keys={k: list(v.shape) for k, v in net.state_dict().items()})
    torch.save(out, os.path.join(GOLD, "eval_ref.pt"))
    print("eval goldens: fid %.6f, mean overlap %.5f alignment %.5f iou %.5f docsim %.5f" % (
        fid, float(overlap.mean()), float(alignment.mean()), float(np.mean(iou)), float(np.mean(docsim))))


def make_tiny_layout_zip(path, seed=3):
    """A tiny dataset in the reference's on-disk format (dataset_tool.py output: non_image.json + per-sample PNGs)."""
    import io, json, zipfile
    import numpy as np
    import PIL.Image
    rng = np.random.RandomState(seed)
    H, W = 48, 64                                              # page size (real data: 1024-class pages)

    def smooth(h, w, c=3):
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        img = np.stack([127 + 120 * np.sin(xx / (3 + 5 * rng.rand()) + 6 * rng.rand()) * np.cos(yy / (2 + 6 * rng.rand())) for _ in range(c)], -1)
        return np.clip(img, 0, 255).astype(np.uint8)

    samples = []
    with zipfile.ZipFile(path, "w", zipfile.ZIP_STORED) as z:
        def put(name, arr):
            buf = io.BytesIO()
            PIL.Image.fromarray(arr).save(buf, format="PNG")
            z.writestr(zipfile.ZipInfo(name, date_time=(2023, 1, 1, 0, 0, 0)), buf.getvalue())     # fixed stamp: reproducible bytes
        for si, n in enumerate([3, 9, 1]):
            base = "page_%02d" % si
            bboxes = np.round(np.stack([rng.uniform(0.2, 0.8, n), rng.uniform(0.2, 0.8, n), rng.uniform(0.1, 0.6, n), rng.uniform(0.03, 0.2, n)], -1), 4)
            for i in range(n):
                ph, pw = [(10, 30), (26, 12), (16, 16)][i % 3]
                put("%s_%d_patch.png" % (base, i), smooth(ph, pw))
                put("%s_%d_patch_orig.png" % (base, i), smooth(H, W))
                put("%s_%d_patch_mask.png" % (base, i), (smooth(H, W, 1)[:, :, 0] > 127).astype(np.uint8) * 255)
            put(base + "_background_orig.png", smooth(H, W))
            samples.append([base, dict(bboxes=bboxes.tolist(), labels=[int(v) for v in rng.randint(0, 8, n)],
                                       texts=["text %d of page %d" % (i, si) for i in range(n)], page_label=None,
                                       attr=dict(name=base, width=W, height=H, num_bbox_labels=8))])
        z.writestr(zipfile.ZipInfo("non_image.json", date_time=(2023, 1, 1, 0, 0, 0)), json.dumps(dict(samples=samples)))


def gen_sampler():
    """Index streams of the reference's InfiniteSampler (torch_utils/misc.py:114-148) for several partitions."""
    import itertools
    from torch_utils import misc
    _orig_init = torch.utils.data.Sampler.__init__
    torch.utils.data.Sampler.__init__ = lambda self, *a, **k: None      # torch >= 2.x: Sampler() takes no data_source (misc.py:120 passes one)
    out = []
    for (n, world, seed, shuffle, window) in [(37, 4, 5, True, 0.5), (8, 2, 0, True, 0.5), (3, 1, 1, True, 0.5), (10, 3, 2, False, 0.5), (50, 8, 7, True, 0.1)]:
        ds = list(range(n))
        streams = [list(int(i) for i in itertools.islice(iter(misc.InfiniteSampler(ds, rank=r, num_replicas=world, shuffle=shuffle, seed=seed, window_size=window)), 120))
                   for r in range(world)]
        out.append(dict(n=n, world=world, seed=seed, shuffle=shuffle, window=window, streams=streams))
    torch.utils.data.Sampler.__init__ = _orig_init
    torch.save(out, os.path.join(GOLD, "sampler_ref.pt"))
    print("sampler goldens:", len(out))


def gen_dataset():
    """tests/golden/tiny_layout.zip + what the reference's LayoutDataset returns for it (background_size 32)."""
    import numpy as np
    import PIL.Image
    np.bool = bool                                             # NumPy 2 / Pillow 12 drift (training/dataset_layoutganpp.py:37,296)
    PIL.Image.ANTIALIAS = PIL.Image.LANCZOS
    from training.dataset_layoutganpp import LayoutDataset
    path = os.path.join(GOLD, "tiny_layout.zip")
    make_tiny_layout_zip(path)
    ds = LayoutDataset(path=path, use_labels=False, max_size=None, xflip=False, background_size=32)
    out = dict(len=len(ds), patch_shape=ds.patch_shape, num_bbox_labels=ds.num_bbox_labels, label_dim=ds.label_dim, name=ds.name, items=[])
    for i in range(len(ds)):
        s, lab = ds[i]
        out["items"].append(dict(
            bboxes=torch.from_numpy(s["bboxes"]), labels=torch.from_numpy(s["labels"]), texts=s["texts"], mask=torch.from_numpy(s["mask"]),
            background=torch.from_numpy(s["background"]), name=s["name"], W_page=s["W_page"], H_page=s["H_page"], label=torch.from_numpy(lab),
            patches_sum=float(s["patches"].astype(np.float64).sum()), patches_abs=float(np.abs(s["patches"]).astype(np.float64).sum()),
            patches_sub=torch.from_numpy(s["patches"][:, :, ::16, ::16].copy()),
            patches_orig_sum=float(s["patches_orig"].astype(np.float64).sum()), patches_orig_shape=list(s["patches_orig"].shape),
            patch_masks_sum=float(s["patch_masks"].astype(np.float64).sum()), patch_masks_shape=list(s["patch_masks"].shape),
            background_orig_sum=float(s["background_orig"].astype(np.float64).sum())))
    torch.save(out, os.path.join(GOLD, "dataset_ref.pt"))
    print("dataset goldens:", out["len"], "samples, patch_shape", out["patch_shape"])


def run_model_goldens(nd, G, D, name, batch, n_valid, seed, ragged=None, light=False):
    inp = make_inputs(batch, n_valid=n_valid, seed=seed) if ragged is None else make_ragged_inputs(ragged, seed=seed)
    store = {}
    hooks = [_hook_outputs(G.text_encoder, store, "G.text_encoder"), _hook_outputs(G.input_proj, store, "G.input_proj"),
             _hook_outputs(G.transformer, store, "G.transformer"), _hook_outputs(G.fc_in, store, "G.fc_in"),
             _hook_outputs(G.backbone[0].body, store, "G.backbone_body")]
    out = {"inputs_seed": seed, "batch": batch, "n_valid": n_valid, "ragged": ragged}
    with torch.no_grad():
        t = time.time()
        res = G(inp["z"], inp["bbox_class"], inp["bbox_real"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"],
                inp["background"], inp["c"], reconst=True)
        print(name, "G fwd %.1fs" % (time.time() - t))
        out["G"] = dict(zip(["bbox_fake", "loss_z", "logit_cls", "loss_lm", "loss_text_len"], [r.clone() for r in res]))
        for h in hooks:
            h.remove()
        out["G_inter"] = {
            "text_cls": store["G.text_encoder"][:, 0, :].clone(),                       # [B*9, 768]
            "input_proj": store["G.input_proj"].to(torch.float16),                      # [B, 256, 8, 8]
            "fc_in": store["G.fc_in"].clone(),
            "hs": store["G.transformer"].clone(),                                       # [B, 9, 256]
            "backbone_mean_abs": float(store["G.backbone_body"].abs().mean()),
            "backbone_sub": store["G.backbone_body"][:, ::16].to(torch.float16),
        }
        t = time.time()
        dres = D(inp["bbox_real"], inp["bbox_class"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"],
                 inp["background"], inp["c"], reconst=True)
        print(name, "D fwd %.1fs" % (time.time() - t))
        names = ["logit_disc", "logit_disc_uncond", "bbox_pred", "logit_cls", "loss_lm", "loss_text_len", "bg_rec",
                 "bbox_pred_uncond", "logit_cls_uncond"]
        d = dict(zip(names, [r.clone() for r in dres]))
        bg = d.pop("bg_rec")
        d["bg_rec_sub"] = bg[:, :, ::8, ::8].clone() if not light else bg[:, :, ::16, ::16].to(torch.float16)
        d["bg_rec_mean"] = float(bg.mean())
        d["bg_rec_std"] = float(bg.std())
        out["D"] = d
        if light:                        # the benched batch size: outputs only (file size)
            out["G_inter"] = {"text_cls": out["G_inter"]["text_cls"].to(torch.float16), "hs": out["G_inter"]["hs"]}
        fres = D(out["G"]["bbox_fake"], inp["bbox_class"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"],
                 inp["background"], inp["c"])
        out["D_fake"] = dict(logit_disc=fres[0].clone(), logit_disc_uncond=fres[1].clone())
    torch.save(out, os.path.join(GOLD, name + ".pt"))
    print("saved", name)


def run_generator_1024(nd, G, name="model_b1_bg1024", seed=21):
    """BASELINE configs[3] geometry: a 1024 x 1024 background -> 32 x 32 = 1024 image tokens; G.forward only (eval)."""
    inp = make_inputs(1, n_valid=8, seed=seed, background_size=1024)
    with torch.no_grad():
        t = time.time()
        out = G(inp["z"], inp["bbox_class"], inp["bbox_real"], inp["bbox_text"], inp["bbox_patch"], inp["padding_mask"], inp["background"], inp["c"])
        print(name, "G fwd %.1fs" % (time.time() - t))
    torch.save({"inputs_seed": seed, "batch": 1, "n_valid": 8, "background_size": 1024, "bbox_fake": out.clone()}, os.path.join(GOLD, name + ".pt"))
    print("saved", name)


def run_loss_goldens(nd, G, D, name, batch, n_valid, seed, small_limit=4096):
    """Full reference loss + backward (training/loss.py accumulate_gradients) with dropout off."""
    import training.loss as ref_loss
    from torch_utils import training_stats
    training_stats.report = lambda name, value: value          # stats collector needs init_multiprocessing
    inp = make_inputs(batch, n_valid=n_valid, seed=seed)
    loss = ref_loss.StyleGAN2Loss(device=torch.device("cpu"), G=G, D=D, r1_gamma=0.0, pl_weight=0.0)
    out = {"inputs_seed": seed, "batch": batch, "n_valid": n_valid, "grads": {}}
    for phase, mod, other in [("Gmain", G, D), ("Dmain", D, G)]:
        for m in (G, D):
            m.requires_grad_(False)
            for p in m.parameters():
                p.grad = None
        mod.requires_grad_(True)
        mod.text_encoder.requires_grad_(False)
        t = time.time()
        loss.accumulate_gradients(phase=phase, bbox_real=inp["bbox_real"], bbox_class=inp["bbox_class"], bbox_text=inp["bbox_text"],
                                  bbox_patch=inp["bbox_patch"], padding_mask=inp["padding_mask"], background=inp["background"],
                                  real_c=inp["c"], gen_z=inp["z"], gen_c=inp["c"], gain=1.0, cur_nimg=0)
        print(name, phase, "fwd+bwd %.1fs" % (time.time() - t))
        norms, small = {}, {}
        for k, p in mod.named_parameters():
            if p.grad is None:
                continue
            norms[k] = float(p.grad.norm())
            if p.grad.numel() <= small_limit:
                small[k] = p.grad.clone()
        out["grads"][phase] = dict(norms=norms, small=small)
    torch.save(out, os.path.join(GOLD, name + ".pt"))
    print("saved", name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-loss", action="store_true")
    ap.add_argument("--skip-model", action="store_true")
    ap.add_argument("--only-eval", action="store_true", help="regenerate tests/golden/eval_ref.pt only")
    ap.add_argument("--only-ragged", action="store_true", help="regenerate tests/golden/model_b3_ragged.pt only")
    ap.add_argument("--only-dataset", action="store_true", help="regenerate tests/golden/tiny_layout.zip + dataset_ref.pt only")
    ap.add_argument("--only-maxiou", action="store_true", help="regenerate tests/golden/maxiou_ref.pt only")
    ap.add_argument("--only-vit", action="store_true", help="regenerate tests/golden/vit_ref.pt only")
    ap.add_argument("--only-bs16", action="store_true", help="model_b16_v8.pt + loss_b16_v8.pt only: the benched batch size (BASELINE configs[1])")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    nd = ref_shim.load()
    torch.set_num_threads(os.cpu_count())
    if args.only_eval:
        gen_eval()
        return
    if args.only_maxiou:
        gen_maxiou()
        return
    if args.only_vit:
        gen_vit()
        return
    if args.only_dataset:
        gen_dataset()
        gen_sampler()
        return
    gen_ops()
    gen_hungarian()
    gen_maxiou()
    gen_vit()
    gen_eval()
    gen_dataset()
    gen_sampler()
    if args.skip_model:
        return
    torch.manual_seed(0)
    G = nd.Generator(**ref_shim.G_KWARGS).eval()
    D = nd.Discriminator(**ref_shim.D_KWARGS).eval()
    synth_state_dict(G)
    synth_state_dict(D)
    if args.only_bs16:
        run_model_goldens(nd, G, D, "model_b16_v8", batch=16, n_valid=8, seed=16, light=True)
        run_loss_goldens(nd, G, D, "loss_b16_v8", batch=16, n_valid=8, seed=16, small_limit=1024)
        return
    if args.only_ragged:
        run_model_goldens(nd, G, D, "model_b3_ragged", batch=3, n_valid=9, seed=13, ragged=[1, 5, 9])
        run_generator_1024(nd, G)
        return
    manifest = {"G": {k: list(v.shape) for k, v in G.state_dict().items()},
                "D": {k: list(v.shape) for k, v in D.state_dict().items()}}
    with open(os.path.join(GOLD, "state_dict_manifest.json"), "w") as f:
        json.dump(manifest, f)
    run_model_goldens(nd, G, D, "model_b1_v4", batch=1, n_valid=4, seed=1)      # BASELINE configs[0]
    run_model_goldens(nd, G, D, "model_b2_v8", batch=2, n_valid=8, seed=2)
    run_model_goldens(nd, G, D, "model_b3_ragged", batch=3, n_valid=9, seed=13, ragged=[1, 5, 9])
    run_generator_1024(nd, G)
    run_model_goldens(nd, G, D, "model_b16_v8", batch=16, n_valid=8, seed=16, light=True)
    if not args.skip_loss:
        run_loss_goldens(nd, G, D, "loss_b2_v8", batch=2, n_valid=8, seed=2)
        run_loss_goldens(nd, G, D, "loss_b16_v8", batch=16, n_valid=8, seed=16, small_limit=1024)


if __name__ == "__main__":
    main()
