"""tcgen05 GEMM parity on the GPU: every operand-major combination, tails, batching, epilogues.
Reference = fp32 matmul of the same bf16-rounded inputs (tolerance: bf16 inputs are exact in the
reference, fp32 accumulation order differs -> 2e-3 relative to the output scale)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).to(torch.bfloat16)


def _check(out, ref, tol=2e-3, what=""):
    scale = ref.abs().max().item() + 1e-6
    err = (out.float() - ref).abs().max().item() / scale
    assert err < tol, "%s: rel err %.3e (scale %.3e)" % (what, err, scale)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 256), (300, 200, 136), (1024, 768, 768),
                                   (9, 8, 40), (4096, 3072, 768), (144, 30524, 768)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_majors(M, N, K, a_mn, b_mn):
    from layoutdetr_b200 import kernels as k
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operand needs ld % 8 == 0")
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = _rand((M, K), g)
    B = _rand((N, K), g)
    ref = A.float() @ B.float().t()
    At = A.t().contiguous() if a_mn else A
    Bt = B.t().contiguous() if b_mn else B
    D = torch.empty((M, N), dtype=torch.float32, device="cuda")
    k.gemm(M, N, K, k.Op(At, At.stride(0), mn=a_mn), k.Op(Bt, Bt.stride(0), mn=b_mn), k.Out(D, N))
    torch.cuda.synchronize()
    _check(D, ref, what="gemm M%d N%d K%d a_mn=%s b_mn=%s" % (M, N, K, a_mn, b_mn))


@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_gemm_epilogue(block_n):
    from layoutdetr_b200 import kernels as k
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 520, 392, 264
    A, B = _rand((M, K), g, 0.5), _rand((N, K), g, 0.5)
    bias = torch.randn(N, generator=g, device="cuda")
    scale = torch.rand(N, generator=g, device="cuda") + 0.5
    R = _rand((M, N), g)
    pre = (A.float() @ B.float().t()) * 0.5 * scale + bias + R.float()
    for act, fn in [(k.ACT_NONE, lambda v: v), (k.ACT_RELU, torch.relu),
                    (k.ACT_GELU, torch.nn.functional.gelu), (k.ACT_LRELU, lambda v: torch.nn.functional.leaky_relu(v, 0.2)),
                    (k.ACT_SIGMOID, torch.sigmoid)]:
        D = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
        aux = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
        k.gemm(M, N, K, k.Op(A, K), k.Op(B, K), k.Out(D, N), alpha=0.5, act=act, post_gain=1.5, R=k.Out(R, N),
               col_scale=scale, col_bias=bias, aux=aux, block_n=block_n)
        torch.cuda.synchronize()
        _check(D, fn(pre) * 1.5, tol=1e-2, what="epilogue act %d" % act)
        _check(aux, pre, tol=1e-2, what="aux act %d" % act)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
def test_gemm_cta_pair_path_with_epilogue(a_mn, b_mn):
    """Shapes with >= 74 tiles of 256x256 take the cta_group::2 kernel; check tails, bias/residual/activation, fp32 out."""
    from layoutdetr_b200 import kernels as k
    g = torch.Generator(device="cuda").manual_seed(23)
    M, N, K = 2304 + 72, 2560 + 40, 328
    A, B = _rand((M, K), g, 0.5), _rand((N, K), g, 0.5)
    bias = torch.randn(N, generator=g, device="cuda")
    R = _rand((M, N), g)
    At = A.t().contiguous() if a_mn else A
    Bt = B.t().contiguous() if b_mn else B
    pre = A.float() @ B.float().t() + bias + R.float()
    D = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    k.gemm(M, N, K, k.Op(At, At.stride(0), mn=a_mn), k.Op(Bt, Bt.stride(0), mn=b_mn), k.Out(D, N), act=k.ACT_RELU,
           R=k.Out(R, N), col_bias=bias)
    torch.cuda.synchronize()
    _check(D, torch.relu(pre), tol=1e-2, what="2sm relu")
    D32 = torch.empty((M, N), dtype=torch.float32, device="cuda")
    k.gemm(M, N, K, k.Op(At, At.stride(0), mn=a_mn), k.Op(Bt, Bt.stride(0), mn=b_mn), k.Out(D32, N), R=k.Out(R, N), col_bias=bias)
    torch.cuda.synchronize()
    _check(D32, pre, tol=2e-3, what="2sm f32")
    # batched: 8 x (512 x 512 x 192) -> 8*2*2 = 32 pair tiles < 74 stays single-CTA; 40 batches -> 160 pair tiles
    nb = 40
    Ab, Bb = _rand((nb * 512, 192), g), _rand((nb * 512, 192), g)
    Db = torch.empty((nb, 512, 512), dtype=torch.float32, device="cuda")
    k.gemm(512, 512, 192, k.Op(Ab, 192, sb1=512 * 192), k.Op(Bb, 192, sb1=512 * 192), k.Out(Db, 512, sb1=512 * 512), nb1=nb)
    torch.cuda.synchronize()
    _check(Db, Ab.float().view(nb, 512, 192) @ Bb.float().view(nb, 512, 192).transpose(1, 2), what="2sm batched")


def test_gemm_batched_strided_attention_layout():
    """Q K^T and P V straight from a fused [B*T, 3*H*d] QKV buffer with two batch dims."""
    from layoutdetr_b200 import kernels as k
    g = torch.Generator(device="cuda").manual_seed(11)
    Bn, T, H, d = 3, 256, 4, 192
    qkv = _rand((Bn * T, 3 * H * d), g, 0.3)
    ld = 3 * H * d
    S = torch.empty((Bn * H, T, T), dtype=torch.float32, device="cuda")
    k.gemm(T, T, d, k.Op(qkv, ld, off=0, sb1=T * ld, sb2=d), k.Op(qkv, ld, off=H * d, sb1=T * ld, sb2=d),
           k.Out(S, T, sb1=H * T * T, sb2=T * T), nb1=Bn, nb2=H, alpha=0.125)
    q = qkv[:, :H * d].float().view(Bn, T, H, d).permute(0, 2, 1, 3)
    kk = qkv[:, H * d:2 * H * d].float().view(Bn, T, H, d).permute(0, 2, 1, 3)
    v = qkv[:, 2 * H * d:].float().view(Bn, T, H, d).permute(0, 2, 1, 3)
    ref = (q @ kk.transpose(-1, -2)) * 0.125
    torch.cuda.synchronize()
    _check(S.view(Bn, H, T, T), ref, what="QK^T")
    P = torch.softmax(ref, -1).to(torch.bfloat16).contiguous().view(Bn * H, T, T)
    O = torch.empty((Bn * T, H * d), dtype=torch.bfloat16, device="cuda")
    k.gemm(T, d, T, k.Op(P, T, sb1=H * T * T, sb2=T * T), k.Op(qkv, ld, off=2 * H * d, sb1=T * ld, sb2=d, mn=True),
           k.Out(O, H * d, sb1=T * H * d, sb2=d), nb1=Bn, nb2=H)
    refO = (P.float().view(Bn, H, T, T) @ v).permute(0, 2, 1, 3).reshape(Bn * T, H * d)
    torch.cuda.synchronize()
    _check(O, refO, tol=1e-2, what="PV")


def test_gemm_small_keys():
    """DETR decoder shapes: 9/10 keys, head_dim 32 (K tails below one 64-wide k-block)."""
    from layoutdetr_b200 import kernels as k
    g = torch.Generator(device="cuda").manual_seed(13)
    Bn, L, H, d = 5, 10, 8, 32
    q = _rand((Bn * L, H * d), g); kk = _rand((Bn * L, H * d), g); v = _rand((Bn * L, H * d), g)
    S = torch.empty((Bn * H, L, L), dtype=torch.float32, device="cuda")
    k.gemm(L, L, d, k.Op(q, H * d, sb1=L * H * d, sb2=d), k.Op(kk, H * d, sb1=L * H * d, sb2=d),
           k.Out(S, L, sb1=H * L * L, sb2=L * L), nb1=Bn, nb2=H)
    qf = q.float().view(Bn, L, H, d).permute(0, 2, 1, 3); kf = kk.float().view(Bn, L, H, d).permute(0, 2, 1, 3)
    vf = v.float().view(Bn, L, H, d).permute(0, 2, 1, 3)
    torch.cuda.synchronize()
    _check(S.view(Bn, H, L, L), qf @ kf.transpose(-1, -2), what="small QK^T")
    Lp = 16
    P = torch.zeros((Bn * H, L, Lp), dtype=torch.bfloat16, device="cuda")
    P[:, :, :L] = torch.softmax(S, -1).to(torch.bfloat16)
    O = torch.empty((Bn * L, H * d), dtype=torch.bfloat16, device="cuda")
    k.gemm(L, d, L, k.Op(P, Lp, sb1=H * L * Lp, sb2=L * Lp), k.Op(v, H * d, sb1=L * H * d, sb2=d, mn=True),
           k.Out(O, H * d, sb1=L * H * d, sb2=d), nb1=Bn, nb2=H)
    refO = (P[:, :, :L].float().view(Bn, H, L, L) @ vf).permute(0, 2, 1, 3).reshape(Bn * L, H * d)
    torch.cuda.synchronize()
    _check(O, refO, tol=1e-2, what="small PV")


def test_gemm_splitk_accumulate():
    from layoutdetr_b200 import kernels as k
    g = torch.Generator(device="cuda").manual_seed(17)
    M, N, K = 768, 768, 8192     # wgrad-like: long K, few tiles
    At = _rand((K, M), g, 0.2); Bt = _rand((K, N), g, 0.2)
    base = torch.randn((M, N), generator=g, device="cuda")
    D = base.clone()
    k.gemm(M, N, K, k.Op(At, M, mn=True), k.Op(Bt, N, mn=True), k.Out(D, N), accumulate=2, split_k=8)
    torch.cuda.synchronize()
    _check(D, base + At.float().t() @ Bt.float(), what="split-k")
    D2 = base.clone()
    k.gemm(M, N, K, k.Op(At, M, mn=True), k.Op(Bt, N, mn=True), k.Out(D2, N), accumulate=1)
    torch.cuda.synchronize()
    _check(D2, base + At.float().t() @ Bt.float(), what="accumulate=1")


@pytest.mark.parametrize("Lq,Lk,d,H,mask_inf,causal", [(256, 256, 192, 4, False, False), (256, 256, 192, 4, False, True),
                                                      (64, 64, 32, 8, True, False), (10, 10, 32, 8, True, False),
                                                      (9, 64, 32, 8, True, False), (200, 136, 64, 2, False, True)])
def test_gemm_fused_softmax_epilogue(Lq, Lk, d, H, mask_inf, causal):
    """QK^T with scale + key/causal mask + softmax + bf16 cast fused into the TMEM drain."""
    from layoutdetr_b200 import kernels as k
    g = torch.Generator(device="cuda").manual_seed(Lq * 3 + Lk)
    Bn = 3
    q = _rand((Bn * Lq, H * d), g, 0.7)
    kk = _rand((Bn * Lk, H * d), g, 0.7)
    km = torch.zeros((Bn, Lk), dtype=torch.uint8, device="cuda")
    km[0, Lk // 2:] = 1
    km[2, -1] = 1
    Lkp = (Lk + 7) // 8 * 8
    P = torch.full((Bn * H, Lq, Lkp), 7.0, dtype=torch.bfloat16, device="cuda")
    scale = 1.0 / d ** 0.5
    k.gemm(Lq, Lk, d, k.Op(q, H * d, sb1=Lq * H * d, sb2=d), k.Op(kk, H * d, sb1=Lk * H * d, sb2=d),
           k.Out(P, Lkp, sb1=H * Lq * Lkp, sb2=Lq * Lkp), nb1=Bn, nb2=H, alpha=scale,
           softmax=dict(key_mask=km, mask_inf=mask_inf, causal=causal))
    qf = q.float().view(Bn, Lq, H, d).permute(0, 2, 1, 3)
    kf = kk.float().view(Bn, Lk, H, d).permute(0, 2, 1, 3)
    s = qf @ kf.transpose(-1, -2) * scale
    neg = float("-inf") if mask_inf else -10000.0
    s = s + torch.zeros_like(s).masked_fill(km.bool()[:, None, None, :], neg)
    if causal:
        s = s + torch.zeros_like(s).masked_fill(torch.ones(Lq, Lk, device="cuda").triu(1).bool()[None, None], neg)
    ref = torch.softmax(s, -1).reshape(Bn * H, Lq, Lk)
    torch.cuda.synchronize()
    torch.testing.assert_close(P[:, :, :Lk].float(), ref, atol=4e-3, rtol=2e-2)
    if Lkp > Lk:
        assert float(P[:, :, Lk:].abs().max()) == 0.0


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("N", [64, 40, 24])
def test_gemm_narrow_n_takes_the_64_wide_tile(N, a_mn, b_mn):
    """N <= 64 (64-channel convolution layers, narrow heads): the 128 x 64 tile (UMMA N = 64) in every operand-major form, fast bf16
    epilogue with bias + ReLU and the fp32 accumulate path."""
    from layoutdetr_b200 import kernels as k
    g = torch.Generator(device="cuda").manual_seed(N + 7)
    M, K = 1000, 200
    A, B = _rand((M, K), g, 0.5), _rand((N, K), g, 0.5)
    At = A.t().contiguous() if a_mn else A
    Bt = B.t().contiguous() if b_mn else B
    if b_mn and N % 8:
        pytest.skip("MN-major operands need a leading dimension that is a multiple of 8")
    bias = torch.randn(N, generator=g, device="cuda")
    ref = A.float() @ B.float().t()
    D = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    k.gemm(M, N, K, k.Op(At, At.stride(0), mn=a_mn), k.Op(Bt, Bt.stride(0), mn=b_mn), k.Out(D, N), act=k.ACT_RELU, col_bias=bias)
    acc = torch.ones((M, N), dtype=torch.float32, device="cuda")
    k.gemm(M, N, K, k.Op(At, At.stride(0), mn=a_mn), k.Op(Bt, Bt.stride(0), mn=b_mn), k.Out(acc, N), accumulate=1)
    torch.cuda.synchronize()
    _check(D, torch.relu(ref + bias), tol=1e-2, what="narrow N=%d" % N)
    _check(acc, ref + 1.0, tol=1e-2, what="narrow N=%d accumulate" % N)


@pytest.mark.parametrize("M,N,K", [(256, 256, 160), (296, 200, 144), (768, 768, 4096), (2048, 256, 160)])
def test_gemm_fp32_accumulate_fast_path(M, N, K):
    """Weight-gradient form (both operands MN-major) accumulated into an fp32 buffer (accumulate = 1): served by the vectorised
    epilogue with the residual aliased to the output; tails in M and N, a device-side alpha, two accumulations in a row."""
    from layoutdetr_b200 import kernels as k
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    At, Bt = _rand((K, M), g, 0.5), _rand((K, N), g, 0.5)
    base = torch.randn((M, N), generator=g, device="cuda")
    D = base.clone()
    alpha_dev = torch.tensor([0.5], device="cuda")
    k.gemm(M, N, K, k.Op(At, M, mn=True), k.Op(Bt, N, mn=True), k.Out(D, N), accumulate=1, alpha_dev=alpha_dev)
    k.gemm(M, N, K, k.Op(At, M, mn=True), k.Op(Bt, N, mn=True), k.Out(D, N), accumulate=1)
    torch.cuda.synchronize()
    prod = At.float().t() @ Bt.float()
    _check(D, base + 1.5 * prod, tol=5e-3, what="fp32 accumulate M%d N%d K%d" % (M, N, K))
