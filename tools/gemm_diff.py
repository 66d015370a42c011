"""Differential test of two builds of liblayoutdetr_sm100.so on the GEMM descriptors a real forward + backward issues.

    LD_OLD_LIB=/path/old.so python tools/gemm_diff.py [bg|bert|detr]

Logs every ld_gemm_bf16 call of a small model pass (shapes, majors, epilogue options), then replays each distinct
unbatched descriptor with random operands through BOTH libraries into NaN-prefilled outputs and compares bit patterns."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["LAYOUTDETR_SYNTHETIC_TOKENIZER"] = "1"
import torch
from layoutdetr_b200 import _lib, kernels as K
from layoutdetr_b200.lanes import LANES


def collect(which):
    LANES.configure(level=0)
    torch.manual_seed(0)
    K.GEMM_LOG = []
    if which == "bg":
        from layoutdetr_b200.training import networks_stylegan2 as sg
        dec = sg.Decoder(z_dim=256, w_dim=512, channel_max=512, channel_base=8192, img_channels=3, img_resolution=256,
                         use_noise=False, num_fp16_res=0, conv_clamp=None, fused_modconv_default=False).cuda()
        dec.requires_grad_(True)
        x0 = torch.randn(2, 256, device="cuda").to(torch.bfloat16).requires_grad_(True)
        img = dec(x0)
        torch.nn.functional.mse_loss(img, torch.randn_like(img)).backward()
    else:
        from helpers import G_KWARGS, D_KWARGS
        from layoutdetr_b200.synthetic import make_inputs
        from layoutdetr_b200.training import networks_detr as nd
        from layoutdetr_b200.training.trainer import Trainer
        G = nd.Generator(**dict(G_KWARGS, bert_num_encoder_layers=2, max_text_length=64)).cuda()
        D = nd.Discriminator(**dict(D_KWARGS, bert_num_encoder_layers=2, max_text_length=64)).cuda()
        tr = Trainer(G, D, torch.device("cuda"), batch_size=2, lr=0.0)
        hb = make_inputs(2, n_valid=8, seed=5)
        dev_b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()}
        z = torch.randn((2, 9, 4), device="cuda")
        tr.iteration(dev_b, z, z)
    torch.cuda.synchronize()
    log, K.GEMM_LOG = K.GEMM_LOG, None
    return log


def pad8(n):
    return (n + 7) // 8 * 8


def replay(entry, libs):
    M, N, Kd, nb, a_mn, b_mn, split_k, caller, o = entry
    if nb != 1 or o["softmax"] or o["aux"]:
        return None
    g = torch.Generator(device="cuda").manual_seed(1)

    def operand(rows, mn):
        if mn:      # element (row, k) at [k * ld + row]
            ld = pad8(rows)
            t = torch.randn((Kd, ld), device="cuda", generator=g).to(torch.bfloat16)
        else:
            ld = pad8(Kd)
            t = torch.randn((rows, ld), device="cuda", generator=g).to(torch.bfloat16)
        return t, ld
    A, lda = operand(M, a_mn)
    B, ldb = operand(N, b_mn)
    d_dt = torch.float32 if "float32" in o["d_dtype"] else torch.bfloat16
    ldd = max(o["ldd"], N)
    R = None
    if o["r"] is not None:
        r_dt = torch.float32 if "float32" in o["r"][0] else torch.bfloat16
        R = torch.randn((M, max(o["r"][1], N)), device="cuda", generator=g).to(r_dt)
    cs = torch.rand(N, device="cuda", generator=g) + 0.5 if o["cs"] else None
    cb = torch.randn(N, device="cuda", generator=g) if o["cb"] else None
    ad = torch.full((1,), 0.75, device="cuda") if o["alpha_dev"] else None
    outs = []
    for lib in libs:
        _lib._lib = lib
        if o["accumulate"]:
            D = torch.randn((M, ldd), device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)).to(d_dt)
        else:
            D = torch.full((M, ldd), float("nan"), device="cuda", dtype=d_dt)
        K.gemm(M, N, Kd, K.Op(A, lda, mn=bool(a_mn)), K.Op(B, ldb, mn=bool(b_mn)), K.Out(D, ldd), alpha=o["alpha"], act=o["act"],
               post_gain=o["post_gain"], accumulate=o["accumulate"], split_k=split_k,
               R=K.Out(R, R.stride(0)) if R is not None else None, col_scale=cs, col_bias=cb, block_n=o["block_n"], alpha_dev=ad)
        torch.cuda.synchronize()
        outs.append(D[:, :N].float().clone())
    a, b = outs
    unwritten = int(torch.isnan(b).sum())
    same = torch.equal(torch.nan_to_num(a, nan=12345.0), torch.nan_to_num(b, nan=12345.0))
    err = float((torch.nan_to_num(a) - torch.nan_to_num(b)).abs().max() / (torch.nan_to_num(a).abs().max() + 1e-20))
    return same, err, unwritten


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "bg"
    new = _lib.lib()
    old = ctypes.CDLL(os.environ["LD_OLD_LIB"])
    old.ld_last_error.restype = ctypes.c_char_p
    log = collect(which)
    seen, bad = set(), 0
    for e in log:
        key = (e[:7], tuple(sorted((k, str(v)) for k, v in e[8].items())))
        if key in seen:
            continue
        seen.add(key)
        r = replay(e, (old, new))
        if r is None:
            continue
        same, err, unw = r
        tol = 2e-2 if e[8]["accumulate"] == 2 or e[6] > 1 else (1e-2 if e[8]["act"] == 2 else 0.0)    # atomics order / new GELU form
        if (not same and err > tol) or unw:
            bad += 1
            print("DIFF err %.3e unwritten %d  %s %s" % (err, unw, e[:8], e[8]))
    _lib._lib = new
    print("%s: %d distinct descriptors replayed, %d differ" % (which, len(seen), bad))
