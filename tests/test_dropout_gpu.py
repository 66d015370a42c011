"""Training-mode dropout (reference nn.Dropout sites: training/med.py:96,213,240,318; training/detr_transformer.py:185-194,
210-214,270-285): Philox masks drawn in-kernel.  Checked against nn.Dropout's statistics (keep rate, 1 / (1 - p) scaling,
unbiased mean), for mask consistency between forward and backward, and for bit-equality with the deterministic path at p = 0."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_dropout_kernel_statistics_and_replay():
    from layoutdetr_b200 import kernels as K, rng
    rng.manual_seed(7)
    x = torch.ones((4096, 768), dtype=torch.bfloat16, device="cuda")
    s1, s2 = rng.next_site(), rng.next_site()
    y1 = K.dropout(x, 0.1, s1)
    y1b = K.dropout(x, 0.1, s1)
    y2 = K.dropout(x, 0.1, s2)
    assert torch.equal(y1, y1b) and not torch.equal(y1, y2)
    keep = (y1 != 0).float()
    assert abs(float(keep.mean()) - 0.9) < 2e-3
    assert abs(float(y1.float()[y1 != 0].mean()) - 1 / 0.9) < 1e-2
    # rows / columns are not correlated: per-column keep rates stay near 0.9
    assert float((keep.mean(0) - 0.9).abs().max()) < 0.03
    # nn.Dropout has the same first two moments
    ref = torch.nn.functional.dropout(x.float(), 0.1, training=True)
    assert abs(float(y1.float().mean()) - float(ref.mean())) < 5e-3
    assert abs(float(y1.float().var()) - float(ref.var())) < 5e-3
    # a new iteration (step + 1) draws a fresh mask for the same site; fp32 tensors use the same mask as bf16
    y32 = K.dropout(x.float(), 0.1, s1)
    assert torch.equal(y32 != 0, y1 != 0)
    rng.advance()
    y3 = K.dropout(x, 0.1, s1)
    assert not torch.equal(y1, y3)


def test_layernorm_residual_dropout_matches_composition():
    from layoutdetr_b200 import kernels as K, rng
    rng.manual_seed(3)
    g = torch.Generator(device="cuda").manual_seed(0)
    rows, C = 1024, 768
    x = torch.randn((rows, C), generator=g, device="cuda").to(torch.bfloat16)
    res = torch.randn((rows, C), generator=g, device="cuda").to(torch.bfloat16)
    gamma = torch.rand(C, generator=g, device="cuda") + 0.5
    beta = torch.randn(C, generator=g, device="cuda")
    site = rng.next_site()
    y, pre, mean, rstd = K.layernorm_res_dropout_fwd(x, res, gamma, beta, 1e-12, 0.1, site, save=True)
    mask = (K.dropout(torch.ones_like(x), 0.1, site) != 0).float()
    pre_ref = x.float() * mask * (65536.0 / (65536.0 - round(0.1 * 65536))) + res.float()
    torch.testing.assert_close(pre, pre_ref, atol=1e-5, rtol=1e-5)
    y_ref = torch.nn.functional.layer_norm(pre_ref, (C,), gamma, beta, 1e-12)
    torch.testing.assert_close(y.float(), y_ref, atol=3e-2, rtol=2e-2)
    # p = 0 is the deterministic kernel, bit for bit
    y0, _, _, _ = K.layernorm_res_dropout_fwd(x, res, gamma, beta, 1e-12, 0.0, 0)
    y_plain, _, _, _ = K.layernorm_fwd(x, gamma, beta, 1e-12, residual=res)
    assert torch.equal(y0, y_plain)


def test_linear_ln_dropout_gradients_under_replayed_mask():
    from layoutdetr_b200 import functional as Fn, kernels as K, rng
    rng.manual_seed(5)
    g = torch.Generator(device="cuda").manual_seed(1)
    M, Kin, N = 512, 256, 256
    lin = torch.nn.Linear(Kin, N).cuda()
    ln = torch.nn.LayerNorm(N).cuda()
    x = torch.randn((M, Kin), generator=g, device="cuda").to(torch.bfloat16).requires_grad_(True)
    res = torch.randn((M, N), generator=g, device="cuda").to(torch.bfloat16).requires_grad_(True)
    dy = torch.randn((M, N), generator=g, device="cuda").to(torch.bfloat16)
    site_next = rng._site[0] + 1
    y = Fn.linear_ln(x, res, lin.weight, lin.bias, ln.weight, ln.bias, 1e-5, dropout_p=0.1)
    y.backward(dy)
    mask = (K.dropout(torch.ones((M, N), dtype=torch.bfloat16, device="cuda"), 0.1, site_next) != 0).float() * (65536.0 / (65536.0 - 6554))
    xr, rr = x.detach().float().requires_grad_(True), res.detach().float().requires_grad_(True)
    w, b = lin.weight.detach().to(torch.bfloat16).float().requires_grad_(True), lin.bias.detach().clone().requires_grad_(True)
    dense = (xr @ w.t() + b)
    y_ref = torch.nn.functional.layer_norm(dense * mask + rr, (N,), ln.weight.detach(), ln.bias.detach(), 1e-5)
    y_ref.backward(dy.float())
    torch.testing.assert_close(y.float(), y_ref, atol=4e-2, rtol=3e-2)
    for got, ref in ((x.grad, xr.grad), (res.grad, rr.grad), (lin.weight.grad, w.grad), (lin.bias.grad, b.grad)):
        sc = float(ref.abs().max())
        assert float((got.float() - ref).abs().max()) < 3e-2 * sc


def test_modules_train_mode_is_stochastic_eval_is_deterministic():
    """BERT layer + DETR encoder layer holders: .eval() reproduces itself bit for bit, .train() draws fresh masks with
    matching first moment."""
    from layoutdetr_b200.training import med
    cfg = med.BertConfig.default(); cfg.num_hidden_layers = 2; cfg.num_attention_heads = 4
    model = med.BertModel(cfg).cuda()
    ids = torch.randint(1000, 20000, (8, 64), device="cuda")
    mask = torch.ones_like(ids)
    model.eval()
    with torch.no_grad():
        a = model.cls_features(ids, mask).clone(); b = model.cls_features(ids, mask).clone()
        assert torch.equal(a, b)
        model.train()
        c = model.cls_features(ids, mask).clone(); d = model.cls_features(ids, mask).clone()
    assert not torch.equal(c, d)
    assert float((c.float() - a.float()).abs().mean()) < 0.5 * float(a.float().abs().mean()) + 0.5
