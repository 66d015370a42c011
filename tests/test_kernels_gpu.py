"""Bytes-bound kernels vs plain PyTorch fp32 references on the GPU (LayerNorm, softmax, CE, casts)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("C", [256, 768])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm_fwd_bwd(C, dtype):
    from layoutdetr_b200 import kernels as k
    g = _gen(C)
    rows = 1000
    x = (torch.randn((rows, C), generator=g, device="cuda") * 2 + 0.3).to(dtype)
    gamma = torch.rand(C, generator=g, device="cuda") + 0.5
    beta = torch.randn(C, generator=g, device="cuda")
    y16, y32, mean, rstd = k.layernorm_fwd(x, gamma, beta, 1e-5, out_bf16=True, out_f32=True, save_stats=True)
    xr = x.float().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (C,), gr, br, 1e-5)
    torch.testing.assert_close(y32, ref, atol=2e-5, rtol=1e-5)
    torch.testing.assert_close(y16.float(), ref, atol=3e-2, rtol=1e-2)
    dy = torch.randn((rows, C), generator=g, device="cuda")
    ref.backward(dy)
    dgamma = torch.zeros(C, device="cuda"); dbeta = torch.zeros(C, device="cuda")
    dx = k.layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, out_dtype=torch.float32)
    torch.testing.assert_close(dx, xr.grad, atol=2e-4, rtol=1e-4)
    torch.testing.assert_close(dgamma, gr.grad, atol=2e-3, rtol=1e-4)
    torch.testing.assert_close(dbeta, br.grad, atol=2e-3, rtol=1e-4)


def test_embed_ln():
    from layoutdetr_b200 import kernels as k
    g = _gen(3)
    V, T, C, n = 1000, 16, 768, 5
    word = torch.randn((V, C), generator=g, device="cuda") * 0.02
    pos = torch.randn((512, C), generator=g, device="cuda") * 0.02
    gamma = torch.rand(C, generator=g, device="cuda") + 0.5
    beta = torch.randn(C, generator=g, device="cuda") * 0.1
    ids = torch.randint(0, V, (n, T), generator=g, device="cuda")
    y, pre, mean, rstd = k.embed_ln_fwd(ids, word, pos, gamma, beta, T, 1e-12, save=True)
    ref_pre = word[ids] + pos[:T][None]
    ref = F.layer_norm(ref_pre, (C,), gamma, beta, 1e-12)
    torch.testing.assert_close(pre.view(n, T, C), ref_pre, atol=1e-7, rtol=0)
    torch.testing.assert_close(y.float().view(n, T, C), ref, atol=3e-2, rtol=1e-2)
    dword = torch.zeros_like(word); dpos = torch.zeros_like(pos)
    dpre = torch.randn((n * T, C), generator=g, device="cuda")
    k.embed_bwd(ids, dpre, dword, dpos, T, pad_id=0)
    ref_dword = torch.zeros_like(word).index_add_(0, ids.view(-1), dpre)
    ref_dword[0] = 0
    torch.testing.assert_close(dword, ref_dword, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(dpos[:T], dpre.view(n, T, C).sum(0), atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("cols,mask_inf,causal", [(256, False, False), (256, False, True), (64, True, False), (10, True, False)])
def test_softmax(cols, mask_inf, causal):
    from layoutdetr_b200 import kernels as k
    g = _gen(cols)
    nb1, nb2, rows = 3, 4, cols
    ldp = (cols + 7) // 8 * 8
    S = torch.randn((nb1 * nb2, rows, cols), generator=g, device="cuda") * 3
    km = torch.zeros((nb1, cols), dtype=torch.uint8, device="cuda")
    km[0, cols // 2:] = 1; km[2, -1] = 1
    P = torch.zeros((nb1 * nb2, rows, ldp), dtype=torch.bfloat16, device="cuda")
    k.softmax_fwd(S, P, nb1, nb2, rows, cols, 0.3, key_mask=km, mask_inf=mask_inf, causal=causal)
    add = torch.zeros((nb1, 1, rows, cols), device="cuda")
    neg = float("-inf") if mask_inf else -10000.0
    add = add.masked_fill(km.bool()[:, None, None, :], neg)
    if causal:
        cm = torch.ones(rows, cols, device="cuda").triu(1).bool()
        add = add.masked_fill(cm[None, None], neg)
    ref = torch.softmax(S.view(nb1, nb2, rows, cols) * 0.3 + add, -1).view(nb1 * nb2, rows, cols)
    torch.testing.assert_close(P[:, :, :cols].float(), ref, atol=4e-3, rtol=1e-2)
    dP = torch.randn((nb1 * nb2, rows, cols), generator=g, device="cuda")
    dS = torch.zeros_like(P)
    k.softmax_bwd(P, dP, dS, nb1 * nb2, rows, cols, 0.3)
    Pf = P[:, :, :cols].float()
    ref_dS = Pf * (dP - (dP * Pf).sum(-1, keepdim=True)) * 0.3
    torch.testing.assert_close(dS[:, :, :cols].float(), ref_dS, atol=1e-2, rtol=2e-2)


@pytest.mark.parametrize("V,eps,dtype", [(30524, 0.1, torch.bfloat16), (256, 0.0, torch.float32), (8, 0.0, torch.float32)])
def test_cross_entropy(V, eps, dtype):
    from layoutdetr_b200 import kernels as k
    g = _gen(V)
    rows = 300
    logits = (torch.randn((rows, V), generator=g, device="cuda") * 2).to(dtype)
    labels = torch.randint(0, V, (rows,), generator=g, device="cuda")
    labels[::7] = -100
    n_valid = int((labels != -100).sum())
    dl = torch.empty((rows, V), dtype=dtype, device="cuda")
    loss_rows = k.cross_entropy(logits, labels, eps, -100, True, dl, grad_scale=1.0 / n_valid)
    lr = logits.float().requires_grad_(True)
    ref = F.cross_entropy(lr, labels, label_smoothing=eps, ignore_index=-100)
    ref.backward()
    torch.testing.assert_close(loss_rows.sum() / n_valid, ref, atol=1e-4, rtol=1e-4)
    tol = 1e-6 if dtype == torch.float32 else 2e-5
    torch.testing.assert_close(dl.float(), lr.grad, atol=tol, rtol=2e-2)


def test_cast_and_axpby():
    from layoutdetr_b200 import kernels as k
    g = _gen(9)
    x = torch.randn((37, 36), generator=g, device="cuda")
    y = k.cast_pad(x, torch.bfloat16, 40)
    assert y.shape == (37, 40)
    torch.testing.assert_close(y[:, :36], x.to(torch.bfloat16), atol=0, rtol=0)
    assert float(y[:, 36:].abs().max()) == 0.0
    a = torch.randn((4, 64, 256), generator=g, device="cuda").to(torch.bfloat16)
    b = torch.randn((64, 256), generator=g, device="cuda")
    out = k.axpby_bcast(a, b, torch.bfloat16)
    torch.testing.assert_close(out.float(), (a.float() + b[None]).to(torch.bfloat16).float(), atol=0, rtol=0)
    z = torch.randn((1024, 1024), generator=g, device="cuda")
    torch.testing.assert_close(k.to_bf16(z), z.to(torch.bfloat16), atol=0, rtol=0)


@pytest.mark.parametrize("C,KH,stride,pad,H", [(3, 7, 2, 3, 32), (5, 3, 1, 1, 9), (12, 3, 2, 1, 10), (16, 3, 1, 1, 8)])
def test_im2col_matches_unfold(C, KH, stride, pad, H):
    """Patch matrix of the conv path (ResNet stem 7x7/2 with C = 3 uses the gather kernel, C % 8 == 0 the vector one):
    bit-exact against torch's unfold, K ordered (kh, kw, c), zero K-padding."""
    import torch.nn.functional as F
    from layoutdetr_b200 import kernels as K
    B, W = 2, H + 3
    x = torch.randn(B, C, H, W, device="cuda").to(torch.bfloat16)
    rows = x.permute(0, 2, 3, 1).contiguous().view(B * H * W, C)
    cols, Ho, Wo = K.im2col(rows, B, H, W, C, KH, KH, stride, pad)
    ref = F.unfold(x.float(), KH, padding=pad, stride=stride)                      # [B, C*KH*KW, L] with (c, kh, kw) order
    ref = ref.view(B, C, KH * KH, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, KH * KH * C)
    Kreal = KH * KH * C
    assert cols.shape == (B * Ho * Wo, (Kreal + 7) // 8 * 8)
    assert torch.equal(cols[:, :Kreal].float(), ref)
    assert float(cols[:, Kreal:].float().abs().max()) == 0.0 if cols.shape[1] > Kreal else True


def test_vectorised_elementwise_kernels_match_torch():
    """128-bit paths of the bytes-bound kernels (column sums, per-sample channel scaling, activation backward,
    demodulation + bias + leaky ReLU, fp32 bias_act) against plain fp32 torch math."""
    from layoutdetr_b200 import kernels as K
    from layoutdetr_b200.torch_utils.ops import bias_act as ba
    torch.manual_seed(0)
    B, P, C = 3, 37, 48
    x = torch.randn(B * P, C, device="cuda").to(torch.bfloat16)
    # column sums (rows not a multiple of the row tiling), accumulated onto existing values, on a strided column slice
    big = torch.randn(1000, 96, device="cuda").to(torch.bfloat16)
    out = torch.ones(48, device="cuda")
    K.colsum_accum(big[:, 48:], out)
    torch.testing.assert_close(out, 1.0 + big[:, 48:].float().sum(0), rtol=1e-5, atol=1e-3)
    # y[b, p, c] = x[b, p, c] * s[b, c]
    s = torch.randn(B, C, device="cuda")
    y = K.scale_channels(x, s, torch.bfloat16, P * C, C)
    ref = (x.float().view(B, P, C) * s[:, None, :]).view(B * P, C)
    torch.testing.assert_close(y.float(), ref.to(torch.bfloat16).float())
    # activation backward (leaky ReLU with gain, exact GELU)
    dy = torch.randn(B * P, C, device="cuda").to(torch.bfloat16)
    d = K.act_bwd(dy, x, K.ACT_LRELU, 2 ** 0.5).float()
    ref = dy.float() * 2 ** 0.5 * torch.where(x.float() > 0, 1.0, 0.2)
    torch.testing.assert_close(d, ref.to(torch.bfloat16).float(), rtol=1e-2, atol=1e-2)
    d = K.act_bwd(dy, x, K.ACT_GELU).float()
    xf = x.float().requires_grad_(True)
    torch.nn.functional.gelu(xf).backward(dy.float())
    torch.testing.assert_close(d, xf.grad.to(torch.bfloat16).float(), rtol=2e-2, atol=2e-2)
    # demodulation + bias + leaky ReLU * gain
    dc = torch.rand(B, C, device="cuda") + 0.5
    bias = torch.randn(C, device="cuda")
    y = K.demod_bias_act_fwd(x, dc, bias, B, P, C, K.ACT_LRELU, 2 ** 0.5).float()
    ref = torch.nn.functional.leaky_relu(x.float().view(B, P, C) * dc[:, None, :] + bias, 0.2).view(B * P, C) * 2 ** 0.5
    torch.testing.assert_close(y, ref.to(torch.bfloat16).float(), rtol=1e-2, atol=1e-2)
    # fp32 bias_act, NCHW with H*W % 4 == 0 (vector path) and % 4 != 0 (scalar path)
    for hw in ((8, 8), (5, 7)):
        img = torch.randn(2, 6, *hw, device="cuda")
        b6 = torch.randn(6, device="cuda")
        got = ba.bias_act(img, b6, act="lrelu", gain=1.5, clamp=2.0)
        ref = (torch.nn.functional.leaky_relu(img + b6[None, :, None, None], 0.2) * 1.5).clamp(-2.0, 2.0)
        torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("C,with_res", [(768, True), (256, True), (768, False), (384, True)])
def test_layernorm_residual_forward(C, with_res):
    """LayerNorm(x + residual) on bf16 rows (128-bit path for C % 256 == 0, generic path otherwise) vs fp32 torch."""
    from layoutdetr_b200 import kernels as k
    torch.manual_seed(1)
    rows = 133
    x = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    res = torch.randn(rows, C, device="cuda").to(torch.bfloat16) if with_res else None
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda")
    y, _, mean, rstd = k.layernorm_fwd(x, gamma, beta, 1e-12, save_stats=True, residual=res)
    pre = x.float() + (res.float() if with_res else 0.0)
    ref = torch.nn.functional.layer_norm(pre, (C,), gamma, beta, 1e-12)
    torch.testing.assert_close(y.float(), ref.to(torch.bfloat16).float(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(mean, pre.mean(1), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rstd, (pre.var(1, unbiased=False) + 1e-12).rsqrt(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("V,in_place", [(30524, False), (30524, True), (32000, False), (1028, True)])
def test_cross_entropy_register_row_path(V, in_place):
    """bf16 rows with a pitch of pad8(V) (the LM-head logits buffer) take the 128-bit register-row kernel; checked like the
    generic kernel against fp32 torch, separately and in place (d logits overwrites the logits), pad columns end up zero."""
    from layoutdetr_b200 import kernels as k
    g = _gen(V + 1)
    rows, eps = 77, 0.1
    Vp = (V + 7) // 8 * 8
    buf = torch.full((rows, Vp), float("nan"), dtype=torch.bfloat16, device="cuda")
    logits = buf[:, :V]
    logits.copy_((torch.randn((rows, V), generator=g, device="cuda") * 3).to(torch.bfloat16))
    labels = torch.randint(0, V, (rows,), generator=g, device="cuda")
    labels[::5] = -100
    labels[1] = V - 1
    labels[2] = 0
    n_valid = int((labels != -100).sum())
    lr = logits.float().clone().requires_grad_(True)
    ref = F.cross_entropy(lr, labels, label_smoothing=eps, ignore_index=-100)
    ref.backward()
    if in_place:
        dl = logits
    else:
        dbuf = torch.full((rows, Vp), float("nan"), dtype=torch.bfloat16, device="cuda")
        dl = dbuf[:, :V]
    loss_rows = k.cross_entropy(logits, labels, eps, -100, True, dl, grad_scale=1.0 / n_valid)
    torch.testing.assert_close(loss_rows.sum() / n_valid, ref, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(dl.float(), lr.grad, atol=2e-5, rtol=2e-2)
    if Vp > V:
        pad = (buf if in_place else dbuf)[:, V:]
        assert float(pad.float().abs().max()) == 0.0


@pytest.mark.parametrize("B,pixels,C", [(4, 64 * 64, 128), (2, 256 * 256, 32), (3, 16 * 16, 512)])
def test_style_gradient_reductions_are_deterministic(B, pixels, C):
    """ld_channel_dot_ws / ld_demod_bias_act_bwd_ws (fixed-order sums through a workspace): bit-identical from run to run and equal to
    the fp32 reference sums (reference: autograd of x * styles and of bias_act in training/networks_stylegan2.py:60-75, 307-325)."""
    from layoutdetr_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(C + B)
    a = torch.randn((B * pixels, C), generator=g, device="cuda").to(torch.bfloat16)
    dy = torch.randn((B * pixels, C), generator=g, device="cuda").to(torch.bfloat16)
    y = torch.randn((B * pixels, C), generator=g, device="cuda").to(torch.bfloat16)
    d = torch.rand((B, C), generator=g, device="cuda") + 0.5
    assert K.DETERMINISTIC_STYLE_SUMS
    outs = [K.channel_dot(a, dy, B, pixels, C) for _ in range(3)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = (a.float() * dy.float()).view(B, pixels, C).sum(1)
    assert float((outs[0] - ref).abs().max()) < 1e-3 * float(ref.abs().max()) + 1e-3
    runs = []
    for _ in range(3):
        dd = torch.zeros((B, C), device="cuda")
        db = torch.zeros(C, device="cuda")
        dx = K.demod_bias_act_bwd(dy, y, a, d, dd, db, B, pixels, C, K.ACT_LRELU, 2 ** 0.5)
        runs.append((dx, dd, db))
    for t0, t1, t2 in zip(*runs):
        assert torch.equal(t0, t1) and torch.equal(t0, t2)
    gact = dy.float() * (2 ** 0.5) * torch.where(y.float() < 0, 0.2, 1.0)
    dd_ref = (gact * a.float()).view(B, pixels, C).sum(1)
    db_ref = gact.view(B * pixels, C).sum(0)
    dx_ref = gact.view(B, pixels, C) * d[:, None, :]
    assert float((runs[0][1] - dd_ref).abs().max()) < 1e-3 * float(dd_ref.abs().max()) + 1e-3
    assert float((runs[0][2] - db_ref).abs().max()) < 1e-3 * float(db_ref.abs().max()) + 1e-2
    assert float((runs[0][0].float().view(B, pixels, C) - dx_ref).abs().max()) < 2e-2 * float(dx_ref.abs().max())
    # the atomics forms stay available (LD_DETERMINISTIC_STYLE_SUMS=0) and agree to summation-order noise
    K.DETERMINISTIC_STYLE_SUMS = False
    try:
        o2 = K.channel_dot(a, dy, B, pixels, C)
    finally:
        K.DETERMINISTIC_STYLE_SUMS = True
    assert float((o2 - outs[0]).abs().max()) < 1e-3 * float(ref.abs().max()) + 1e-3


@pytest.mark.parametrize("B,H,C,sep", [(2, 33, 32, True), (1, 129, 64, True), (3, 17, 512, True), (2, 21, 32, False)])
def test_upfirdn_nhwc_tiled_fir_matches_reference(B, H, C, sep):
    """The shared-memory tiled FIR (up = down = 1, channels-last bf16: the filter after a stride-2 transposed convolution,
    training/networks_stylegan2.py:307-325) vs the oracle's restatement of _upfirdn2d_ref (torch_utils/ops/upfirdn2d.py:168-214)
    on the same bf16-rounded input — rank-1 filter (two 1-D passes) and a non-separable filter (2-D loop), values and the
    backward pass (another upfirdn2d with the flipped filter)."""
    from layoutdetr_b200 import functional as Fn
    from oracle import layoutdetr_oracle as O
    g = torch.Generator(device="cuda").manual_seed(H * 7 + C)
    x = torch.randn((B * H * H, C), generator=g, device="cuda").to(torch.bfloat16)
    f1 = torch.tensor([1.0, 3.0, 3.0, 1.0])
    f = torch.outer(f1, f1)
    if not sep:
        f = f + torch.tensor([[0.0, 0.5, 0.0, 0.0], [0.0, 0.0, 0.0, 0.25], [0.3, 0.0, 0.0, 0.0], [0.0, 0.0, 0.1, 0.0]])
    f = (f / f.sum()).cuda()
    xi = x.clone().requires_grad_(True)
    y = Fn.upfirdn_nhwc(xi, f, B, H, H, pad=(1, 1, 1, 1), gain=4.0)
    oh = H - 1
    dy = torch.randn((B * oh * oh, C), generator=g, device="cuda").to(torch.bfloat16)
    y.backward(dy)
    torch.cuda.synchronize()
    x4 = x.float().cpu().view(B, H, H, C).permute(0, 3, 1, 2).clone().requires_grad_(True)
    ref = O.upfirdn2d(x4, f.cpu(), padding=(1, 1, 1, 1), gain=4.0)
    ref.backward(dy.float().cpu().view(B, oh, oh, C).permute(0, 3, 1, 2))
    y_ref = ref.detach().permute(0, 2, 3, 1).reshape(B * oh * oh, C)
    dx_ref = x4.grad.permute(0, 2, 3, 1).reshape(B * H * H, C)
    assert tuple(y.shape) == tuple(y_ref.shape)
    assert float((y.float().cpu() - y_ref).abs().max()) < 1e-2 * float(y_ref.abs().max())
    assert float((xi.grad.float().cpu() - dx_ref).abs().max()) < 1e-2 * float(dx_ref.abs().max())
