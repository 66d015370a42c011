"""GPU: implicit-GEMM convolution (ld_conv_gemm_bf16 — TMA boxes of pixels x 64 channels straight from the NHWC image, no patch
matrix) behind functional.Conv2dFn: forward, data gradient and weight gradient against torch's fp32 convolution on the same
bf16-rounded operands, for the ResNet-50 / StyleGAN2 geometries of the path (training/detr_backbone.py:82-95,
training/networks_stylegan2.py:30-83) and the shapes of its box rule: rows wider than a box, several rows per box, several images
per box, odd batch, stride 2, 1x1 stride 2, Cout not a multiple of the tile."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # B, H, W, Cin, Cout, k, stride, pad
    (4, 64, 64, 64, 64, 3, 1, 1),       # layer1 conv2: two rows per 128-pixel box
    (2, 64, 64, 128, 128, 3, 2, 1),     # layer2.0 conv2: stride 2 (traversal stride of the tensor map)
    (3, 16, 16, 256, 256, 3, 1, 1),     # layer3: 8 rows per box, odd batch
    (3, 8, 8, 512, 512, 3, 1, 1),       # layer4: two images per box, odd batch (last box half out of range)
    (2, 64, 64, 256, 512, 1, 2, 0),     # 1x1 stride-2 downsample
    (1, 128, 128, 64, 96, 3, 1, 1),     # a row is exactly one box; Cout with a ragged N tile
    (1, 256, 256, 64, 32, 3, 1, 1),     # StyleGAN2 256^2 layer: rows of two boxes
    (2, 32, 32, 128, 64, 3, 1, 1),
    (1, 256, 256, 32, 32, 3, 1, 1),     # StyleGAN2 256^2 layer with 32 channels: 64-byte-swizzle K blocks (forward + dgrad; wgrad explicit)
    (2, 64, 64, 32, 64, 3, 1, 1),
    (3, 16, 16, 64, 32, 3, 1, 1),       # Cout = 32: the data gradient runs over a 32-channel dy
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride,pad", CASES)
def test_implicit_conv_matches_torch(B, H, W, Cin, Cout, k, stride, pad):
    from layoutdetr_b200 import functional as Fn, kernels as K
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H + Cin + Cout + k)
    x = torch.randn((B * H * W, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((Cout, Cin, k, k), generator=g, device="cuda") * (1.0 / (k * k * Cin) ** 0.5)).requires_grad_(True)
    scale = torch.rand(Cout, generator=g, device="cuda") + 0.5
    shift = torch.randn(Cout, generator=g, device="cuda")
    Ho, Wo = K.conv_out_size(H, k, stride, pad), K.conv_out_size(W, k, stride, pad)
    dy = torch.randn((B * Ho * Wo, Cout), generator=g, device="cuda").to(torch.bfloat16)
    used = []
    real_gemm = K.gemm

    def spy(*a, **kw):
        used.append(kw.get("conv", None) and kw["conv"]["mode"])
        return real_gemm(*a, **kw)

    K.gemm = spy
    try:
        assert Fn.IMPLICIT_CONV
        xi = x.clone().requires_grad_(True)
        y = Fn.conv2d(xi, w, scale, shift, None, B, H, W, stride, pad, K.ACT_RELU)
        y.backward(dy)
    finally:
        K.gemm = real_gemm
    torch.cuda.synchronize()
    assert 1 in used, used                                    # the forward went implicit
    if Cin % 64 == 0:
        assert 2 in used, used                                # and so did the weight gradient (whole 64-channel blocks only)
    if stride == 1 and (Cout % 64 == 0 or Cout == 32):        # dy has Cout channels: the implicit data gradient needs 64-channel blocks or exactly 32
        assert used.count(1) == 2, used
    # fp32 reference on the same bf16-rounded operands
    xr = x.float().view(B, H, W, Cin).permute(0, 3, 1, 2).clone().requires_grad_(True)
    wr = w.detach().to(torch.bfloat16).float().requires_grad_(True)
    yr = torch.relu(F.conv2d(xr, wr, stride=stride, padding=pad) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    yr.backward(dy.float().view(B, Ho, Wo, Cout).permute(0, 3, 1, 2))
    y_ref = yr.permute(0, 2, 3, 1).reshape(B * Ho * Wo, Cout)
    dx_ref = xr.grad.permute(0, 2, 3, 1).reshape(B * H * W, Cin)
    # dx / dw: the ReLU mask comes from the bf16 output here and from the fp32 output in the reference — where the pre-activation is
    # within a rounding of 0 one term of the sum flips, which the max norm sees for short sums (288 terms at 32 channels) — so the
    # gradients are held to a relative L2 bound and a looser max bound
    for name, got, ref, tol, tol_l2 in (("y", y, y_ref, 1e-2, 3e-3), ("dx", xi.grad, dx_ref, 6e-2, 1e-2), ("dw", w.grad, wr.grad, 2e-2, 1e-2)):
        ref = ref.detach()
        sc = float(ref.abs().max()) + 1e-12
        err = float((got.float() - ref).abs().max()) / sc
        l2 = float((got.float() - ref).norm() / (ref.norm() + 1e-12))
        assert err < tol and l2 < tol_l2, (name, err, l2)


def test_implicit_and_explicit_paths_agree():
    """Same convolution through the patch-matrix path (LD_CONV_IMPLICIT=0 semantics) and the implicit one: identical operands, so the
    results agree to accumulation-order noise of the bf16 store."""
    from layoutdetr_b200 import functional as Fn, kernels as K
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, W, Cin, Cout = 2, 32, 32, 128, 128
    x = torch.randn((B * H * W, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((Cout, Cin, 3, 3), generator=g, device="cuda") * 0.03)
    dy = torch.randn((B * H * W, Cout), generator=g, device="cuda").to(torch.bfloat16)
    outs = []
    for flag in (True, False):
        Fn.IMPLICIT_CONV = flag
        try:
            xi = x.clone().requires_grad_(True)
            wi = w.clone().requires_grad_(True)
            y = Fn.conv2d(xi, wi, None, None, None, B, H, W, 1, 1, K.ACT_NONE)
            y.backward(dy)
            outs.append((y.detach().float(), xi.grad.float(), wi.grad.float()))
        finally:
            Fn.IMPLICIT_CONV = True
    for a, b in zip(*outs):
        assert float((a - b).abs().max()) <= 1e-2 * float(b.abs().max())


def test_conv_gemm_rejects_bad_geometry():
    from layoutdetr_b200 import kernels as K
    from layoutdetr_b200._lib import LayoutDetrKernelError
    x = torch.zeros((1 * 12 * 12, 64), dtype=torch.bfloat16, device="cuda")       # 12-pixel rows: 128 % 12 != 0
    w = torch.zeros((64, 9 * 64), dtype=torch.bfloat16, device="cuda")
    out = torch.empty((144, 64), dtype=torch.bfloat16, device="cuda")
    assert not K.conv_box_ok(12, 12, 1, 128)
    with pytest.raises(LayoutDetrKernelError, match="rectangle of whole rows"):
        K.gemm(144, 64, 576, K.Op(x, 64), K.Op(w, 576), K.Out(out, 64),
               conv=dict(img=x, mode=1, B=1, H=12, W=12, C=64, Ho=12, Wo=12, KH=3, KW=3, stride=1, pad=1))
