"""Fused attention kernel (QK^T -> mask/softmax -> PV on chip) vs an fp32 torch restatement of the reference math."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, B, H, Lq, Lk, d, scale, km, mask_inf, causal):
    qf = q.float().view(B, Lq, H, d).permute(0, 2, 1, 3)
    kf = k.float().view(B, Lk, H, d).permute(0, 2, 1, 3)
    vf = v.float().view(B, Lk, H, d).permute(0, 2, 1, 3)
    s = qf @ kf.transpose(-1, -2) * scale
    neg = float("-inf") if mask_inf else -10000.0
    if km is not None:
        s = s + torch.zeros_like(s).masked_fill(km.bool()[:, None, None, :], neg)
    if causal:
        s = s + torch.zeros_like(s).masked_fill(torch.ones(Lq, Lk, device=s.device).triu(1).bool()[None, None], neg)
    p = torch.softmax(s, -1)
    o = (p @ vf).permute(0, 2, 1, 3).reshape(B * Lq, H * d)
    return o, p.reshape(B * H, Lq, Lk)


@pytest.mark.parametrize("Lq,Lk,d,H,mask_inf,causal,use_mask", [
    (256, 256, 192, 4, False, False, True), (256, 256, 192, 4, False, True, True), (48, 48, 192, 4, False, True, True),
    (64, 64, 32, 8, True, False, False), (10, 10, 32, 8, True, False, True), (9, 64, 32, 8, True, False, False),
    (200, 136, 64, 2, False, True, True), (130, 9, 32, 8, True, False, True), (1024, 256, 32, 8, True, False, False)])
def test_fused_attention_forward(Lq, Lk, d, H, mask_inf, causal, use_mask):
    from layoutdetr_b200 import kernels as k
    g = torch.Generator(device="cuda").manual_seed(Lq * 7 + Lk + d)
    B = 3
    qkv_q = (torch.randn((B * Lq, 3 * H * d), generator=g, device="cuda") * 0.8).to(torch.bfloat16)   # fused-buffer style strides
    kv = (torch.randn((B * Lk, 2 * H * d), generator=g, device="cuda") * 0.8).to(torch.bfloat16)
    km = None
    if use_mask:
        km = torch.zeros((B, Lk), dtype=torch.uint8, device="cuda")
        km[0, Lk // 2:] = 1
        km[2, -1] = 1
    scale = 1.0 / d ** 0.5
    Lkp = (Lk + 7) // 8 * 8
    lse = torch.zeros((B * H, Lq), dtype=torch.float32, device="cuda")
    O = k.attention_fwd(qkv_q, H * d, kv, 0, kv, H * d, B, H, Lq, Lk, d, scale, key_mask=km, mask_inf=mask_inf, causal=causal)
    O2 = k.attention_fwd(qkv_q, H * d, kv, 0, kv, H * d, B, H, Lq, Lk, d, scale, key_mask=km, mask_inf=mask_inf, causal=causal, lse_out=lse)
    # the probabilities themselves: the backward kernel recomputes them from lse (dropout off: Pd == P)
    P = torch.zeros((B * H, Lq, Lkp), dtype=torch.bfloat16, device="cuda")
    dS = torch.zeros_like(P)
    dq = torch.zeros_like(qkv_q)
    k.attention_bwd(qkv_q, H * d, kv, 0, kv, H * d, O2, torch.zeros_like(O2), lse, dq, P, dS, B, H, Lq, Lk, d, scale, key_mask=km,
                    mask_inf=mask_inf, causal=causal)
    torch.cuda.synchronize()
    assert float(dS.float().abs().max()) == 0.0 and float(dq.float().abs().max()) == 0.0       # zero upstream gradient
    q = qkv_q[:, H * d:2 * H * d]
    ref_o, ref_p = _ref(q, kv[:, :H * d], kv[:, H * d:], B, H, Lq, Lk, d, scale, km, mask_inf, causal)
    sc = float(ref_o.abs().max())
    assert float((O.float() - ref_o).abs().max()) < 2e-2 * sc, float((O.float() - ref_o).abs().max()) / sc
    assert torch.equal(O, O2)
    torch.testing.assert_close(P[:, :, :Lk].float(), ref_p, atol=4e-3, rtol=2e-2)
    # per-row log-sum-exp of the scaled + masked scores, log2 domain (what ld_attention_bwd recomputes P from)
    ref_lse = _ref_lse(q, kv[:, :H * d], B, H, Lq, Lk, d, scale, km, mask_inf, causal)
    torch.testing.assert_close(lse, ref_lse, atol=2e-2, rtol=1e-3)


def _ref_lse(q, k, B, H, Lq, Lk, d, scale, km, mask_inf, causal):
    qf = q.float().view(B, Lq, H, d).permute(0, 2, 1, 3)
    kf = k.float().view(B, Lk, H, d).permute(0, 2, 1, 3)
    s = qf @ kf.transpose(-1, -2) * scale
    neg = float("-inf") if mask_inf else -10000.0
    if km is not None:
        s = s + torch.zeros_like(s).masked_fill(km.bool()[:, None, None, :], neg)
    if causal:
        s = s + torch.zeros_like(s).masked_fill(torch.ones(Lq, Lk, device=s.device).triu(1).bool()[None, None], neg)
    return (torch.logsumexp(s, -1) * 1.4426950408889634).reshape(B * H, Lq)


@pytest.mark.parametrize("Lq,Lk,d,H,mask_inf,causal,use_mask", [
    (256, 256, 192, 4, False, True, True), (256, 256, 192, 4, False, False, True), (64, 64, 32, 8, True, False, False),
    (10, 64, 32, 8, True, False, False), (9, 9, 32, 8, True, False, True), (200, 136, 64, 2, False, True, True),
    (300, 40, 192, 4, False, False, True)])
def test_fused_attention_backward_vs_fp32(Lq, Lk, d, H, mask_inf, causal, use_mask):
    """ld_attention_bwd (P recomputed from the saved log-sum-exp, dQ on chip) + the two transposed GEMMs vs fp32 autograd."""
    from layoutdetr_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(Lq * 5 + Lk + d)
    B = 2
    qb = (torch.randn((B * Lq, 2 * H * d), generator=g, device="cuda") * 0.7).to(torch.bfloat16)        # q in the second half
    kvb = (torch.randn((B * Lk, 2 * H * d), generator=g, device="cuda") * 0.7).to(torch.bfloat16)
    dO = torch.randn((B * Lq, H * d), generator=g, device="cuda").to(torch.bfloat16)
    km = None
    if use_mask:
        km = torch.zeros((B, Lk), dtype=torch.uint8, device="cuda")
        km[0, Lk // 2:] = 1
        km[1, -1] = 1
    scale = 1.0 / d ** 0.5
    q_in, kv_in = qb.clone().requires_grad_(True), kvb.clone().requires_grad_(True)
    o = Fn.attention(q_in, kv_in, kv_in, H * d, 0, H * d, B, H, Lq, Lk, d, scale, key_mask=km, mask_inf=mask_inf, causal=causal)
    o.backward(dO)
    torch.cuda.synchronize()
    qr = qb[:, H * d:].float().clone().requires_grad_(True)
    kr = kvb[:, :H * d].float().clone().requires_grad_(True)
    vr = kvb[:, H * d:].float().clone().requires_grad_(True)
    ref_o, _ = _ref(qr, kr, vr, B, H, Lq, Lk, d, scale, km, mask_inf, causal)
    ref_o.backward(dO.float())
    for name, got, ref in (("dq", q_in.grad[:, H * d:], qr.grad), ("dk", kv_in.grad[:, :H * d], kr.grad), ("dv", kv_in.grad[:, H * d:], vr.grad)):
        sc = float(ref.abs().max())
        err = float((got.float() - ref).abs().max())
        assert err < 3e-2 * sc, (name, err / sc)
    assert float(q_in.grad[:, :H * d].abs().max()) == 0.0                    # columns no head reads get zero gradient


def test_fused_attention_dropout_mask_is_consistent_and_unbiased():
    """Forward and backward kernels regenerate the SAME Philox mask: O(dropout) == Pd @ V with the backward kernel's Pd;
    keep rate ~ 1 - p, kept probabilities scaled by 1 / (1 - p); gradients match fp32 autograd under the replayed mask."""
    from layoutdetr_b200 import kernels as k, rng
    B, H, L, d, p = 2, 4, 256, 192, 0.1
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = (torch.randn((B * L, 3 * H * d), generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    dO = torch.randn((B * L, H * d), generator=g, device="cuda").to(torch.bfloat16)
    km = torch.zeros((B, L), dtype=torch.uint8, device="cuda"); km[1, 200:] = 1
    scale = d ** -0.5
    rng.manual_seed(1234)
    site = rng.next_site()
    lse = torch.empty((B * H, L), dtype=torch.float32, device="cuda")
    O = k.attention_fwd(qkv, 0, qkv, H * d, qkv, 2 * H * d, B, H, L, L, d, scale, key_mask=km, lse_out=lse, dropout_p=p, rng_site=site)
    O_again = k.attention_fwd(qkv, 0, qkv, H * d, qkv, 2 * H * d, B, H, L, L, d, scale, key_mask=km, dropout_p=p, rng_site=site)
    O_other = k.attention_fwd(qkv, 0, qkv, H * d, qkv, 2 * H * d, B, H, L, L, d, scale, key_mask=km, dropout_p=p, rng_site=site + 1)
    assert torch.equal(O, O_again) and not torch.equal(O, O_other)
    dq = torch.zeros_like(qkv)
    Pd = torch.empty((B * H, L, L), dtype=torch.bfloat16, device="cuda")
    dS = torch.empty_like(Pd)
    k.attention_bwd(qkv, 0, qkv, H * d, qkv, 2 * H * d, O, dO, lse, dq, Pd, dS, B, H, L, L, d, scale, key_mask=km, dropout_p=p, rng_site=site)
    torch.cuda.synchronize()
    q, kk, v = qkv[:, :H * d], qkv[:, H * d:2 * H * d], qkv[:, 2 * H * d:]
    _, P_ref = _ref(q, kk, v, B, H, L, L, d, scale, km, False, False)
    big = P_ref > 1e-3
    keep = (Pd.float() != 0) & big
    rate = float(keep.sum()) / float(big.sum())
    assert abs(rate - (1 - p)) < 0.01, rate
    ratio = (Pd.float()[keep] / P_ref[keep])
    assert abs(float(ratio.mean()) - 1 / (1 - p)) < 0.02
    vf = v.float().view(B, L, H, d).permute(0, 2, 1, 3)
    O_from_Pd = (Pd.float().view(B, H, L, L) @ vf).permute(0, 2, 1, 3).reshape(B * L, H * d)
    sc = float(O_from_Pd.abs().max())
    assert float((O.float() - O_from_Pd).abs().max()) < 2e-2 * sc
    # gradients under the replayed mask
    M = ((Pd.float() != 0) | ~big).float().view(B, H, L, L)            # tiny probabilities: treat as kept (their weight is negligible)
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, kk, v))
    qf = qr.view(B, L, H, d).permute(0, 2, 1, 3); kf = kr.view(B, L, H, d).permute(0, 2, 1, 3); vf2 = vr.view(B, L, H, d).permute(0, 2, 1, 3)
    s = qf @ kf.transpose(-1, -2) * scale + torch.zeros(B, 1, 1, L, device="cuda").masked_fill(km.bool()[:, None, None, :], -10000.0)
    o_ref = ((torch.softmax(s, -1) * M / (1 - p)) @ vf2).permute(0, 2, 1, 3).reshape(B * L, H * d)
    o_ref.backward(dO.float())
    sc = float(qr.grad.abs().max())
    assert float((dq[:, :H * d].float() - qr.grad).abs().max()) < 4e-2 * sc
    ds_ref_dk = kr.grad                                                     # dK = dS^T Q through the kernel's dS
    dk = (dS.float().view(B, H, L, L).transpose(-1, -2) @ q.float().view(B, L, H, d).permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(B * L, H * d)
    assert float((dk - ds_ref_dk).abs().max()) < 4e-2 * float(ds_ref_dk.abs().max())


def test_attention_autograd_paths_agree():
    """Fused forward / backward kernels vs the unfused batched-GEMM path: same outputs and gradients."""
    from layoutdetr_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, L, d = 2, 4, 64, 192
    base = (torch.randn((B * L, 3 * H * d), generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    km = torch.zeros((B, L), dtype=torch.uint8, device="cuda"); km[1, 40:] = 1
    dO = (torch.randn((B * L, H * d), generator=g, device="cuda")).to(torch.bfloat16)
    outs = []
    for fused in (True, False):
        Fn.FUSED_ATTENTION = fused
        x = base.clone().requires_grad_(True)
        o = Fn.attention(x, x, x, 0, H * d, 2 * H * d, B, H, L, L, d, key_mask=km, mask_inf=False, causal=True)
        o.backward(dO)
        outs.append((o.detach().float(), x.grad.float()))
    Fn.FUSED_ATTENTION = True
    torch.testing.assert_close(outs[0][0], outs[1][0], atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(outs[0][1], outs[1][1], atol=3e-2, rtol=3e-2)


