"""GPU: the op-level conv2d_gradfix API (reference torch_utils/ops/conv2d_gradfix.py:37-42) vs torch's fp32 convolutions —
values, first-order gradients and the SECOND-order gradients R1 / path-length regularisation take (reference training/loss.py:132,210:
grad(..., create_graph=True) followed by backward through the penalty)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    sc = float(b.abs().max()) + 1e-12
    err = float((a.float() - b.float()).abs().max()) / sc
    assert err < tol, (what, err)


@pytest.mark.parametrize("k,stride,pad,cin,cout,transpose", [
    (3, 1, 1, 16, 24, False), (1, 1, 0, 32, 8, False), (3, 2, 1, 8, 16, False), (3, 1, 1, 16, 3, False),
    (3, 2, 0, 16, 8, True), (4, 2, 1, 8, 8, True)])
def test_conv_values_first_and_second_order(k, stride, pad, cin, cout, transpose):
    from layoutdetr_b200.torch_utils.ops import conv2d_gradfix as cg
    g = torch.Generator(device="cuda").manual_seed(k * 100 + stride * 10 + cin)
    x = torch.randn((2, cin, 12, 12), generator=g, device="cuda")
    w = torch.randn((cin, cout, k, k) if transpose else (cout, cin, k, k), generator=g, device="cuda") * 0.2
    b = torch.randn(cout, generator=g, device="cuda")
    ours = (lambda xx, ww: cg.conv_transpose2d(xx, ww, b, stride=stride, padding=pad)) if transpose else \
        (lambda xx, ww: cg.conv2d(xx, ww, b, stride=stride, padding=pad))
    ref = (lambda xx, ww: F.conv_transpose2d(xx, ww, b, stride=stride, padding=pad)) if transpose else \
        (lambda xx, ww: F.conv2d(xx, ww, b, stride=stride, padding=pad))
    # bf16-rounded operands on the reference side: what the tensor cores see
    xq, wq = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()
    outs = []
    for fn, xi, wi in ((ours, x, w), (ref, xq, wq)):
        xi = xi.clone().requires_grad_(True); wi = wi.clone().requires_grad_(True)
        y = fn(xi, wi)
        (gx,) = torch.autograd.grad(y.square().sum(), xi, create_graph=True)          # first order, kept differentiable
        penalty = gx.square().sum()                                                   # R1-style penalty on the input gradient
        ggx, ggw = torch.autograd.grad(penalty, (xi, wi))                             # second order
        gw1, = torch.autograd.grad(fn(xi, wi).sum(), wi)
        outs.append((y.detach(), gx.detach(), gw1, ggx, ggw))
    for name, a, r in zip(("y", "dx", "dw", "d2x", "d2w"), outs[0], outs[1]):
        _close(a, r, 4e-2, name)


def test_no_weight_gradients_context():
    """`no_weight_gradients()` (reference torch_utils/ops/conv2d_gradfix.py:37-42, used by R1 in training/loss.py:209-215):
    backward passes executed INSIDE the context skip the first-order weight gradient; the penalty's own backward, outside
    the context, still reaches the weights through the second-order terms."""
    from layoutdetr_b200.torch_utils.ops import conv2d_gradfix as cg
    x = torch.randn((1, 8, 8, 8), device="cuda", requires_grad=True)
    w = torch.randn((8, 8, 3, 3), device="cuda", requires_grad=True)
    y = cg.conv2d(x, w, padding=1)
    with cg.no_weight_gradients():
        gx, gw = torch.autograd.grad(y.square().sum(), [x, w], create_graph=True, allow_unused=True)
    assert gw is None and gx is not None
    gx.square().sum().backward()
    assert w.grad is not None and x.grad is not None
    # against torch's own double backward of the same expression (bf16 tensor-core contraction: loose tolerance)
    x2, w2 = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    y2 = torch.nn.functional.conv2d(x2, w2, padding=1)
    gx2, = torch.autograd.grad(y2.square().sum(), [x2], create_graph=True)
    gx2.square().sum().backward()
    for a, b in ((x.grad, x2.grad), (w.grad, w2.grad)):
        assert float((a - b).norm() / b.norm()) < 5e-2
