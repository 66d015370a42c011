"""Box-loss kernels (csrc/box_loss.cu) through the C-ABI against the CPU oracle: values and gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(seed, B=6, N=9):
    g = torch.Generator().manual_seed(seed)
    bbox = torch.rand((B, N, 4), generator=g) * torch.tensor([0.6, 0.6, 0.5, 0.3]) + torch.tensor([0.2, 0.2, 0.05, 0.03])
    mask = torch.ones((B, N), dtype=torch.bool)
    mask[:, N - 1:] = False
    mask[1, 3:] = False
    mask[2, 1:] = False
    if seed % 2 == 0:
        bbox[:, N - 1:] = torch.rand((B, 1, 4), generator=g)
    return bbox, mask, g


@pytest.mark.parametrize("seed,N", [(0, 9), (1, 9), (2, 12), (3, 33)])
def test_layout_losses_match_oracle(seed, N):
    from layoutdetr_b200 import box_ops
    from layoutdetr_b200.metrics import metric_layoutnet as ml
    from oracle import layoutdetr_oracle as O
    bbox, mask, g = _case(seed, N=N)
    B = bbox.shape[0]
    w1, w2 = torch.rand(B, generator=g), torch.rand(B, generator=g)
    b_ref = bbox.clone().requires_grad_(True)
    ov_ref, al_ref = O.compute_overlap(b_ref, mask), O.compute_alignment(b_ref, mask)
    ((ov_ref * w1).sum() + (al_ref * w2).sum()).backward()
    b_dev = bbox.cuda().requires_grad_(True)
    ov, al = ml.layout_overlap_alignment(b_dev, mask.cuda())
    ((ov * w1.cuda()).sum() + (al * w2.cuda()).sum()).backward()
    torch.testing.assert_close(ov.detach().cpu(), ov_ref.detach(), atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(al.detach().cpu(), al_ref.detach(), atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(b_dev.grad.cpu(), b_ref.grad, atol=2e-5, rtol=1e-4)
    # the two single-loss entry points of the reference API, and the no-grad path
    with torch.no_grad():
        torch.testing.assert_close(ml.compute_overlap(bbox.cuda(), mask.cuda()).cpu(), ov_ref.detach(), atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(ml.compute_alignment(bbox.cuda(), mask.cuda()).cpu(), al_ref.detach(), atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("M", [1, 37, 128, 1000])
def test_giou_loss_matches_oracle(M):
    from layoutdetr_b200.metrics import metric_layoutnet as ml
    from oracle import layoutdetr_oracle as O
    g = torch.Generator().manual_seed(M)
    mk = lambda: torch.rand((M, 4), generator=g) * torch.tensor([0.6, 0.6, 0.5, 0.3]) + torch.tensor([0.2, 0.2, 0.05, 0.03])
    fake, real = mk(), mk()
    f_ref = fake.clone().requires_grad_(True)
    ref = O.generalized_iou_loss(f_ref, real)
    (ref * 4.0).backward()
    f_dev = fake.cuda().requires_grad_(True)
    out = ml.generalized_iou_loss(f_dev, real.cuda())
    (out * 4.0).backward()
    torch.testing.assert_close(out.detach().cpu(), ref.detach(), atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(f_dev.grad.cpu(), f_ref.grad, atol=1e-6, rtol=1e-4)


def test_layout_losses_inside_cuda_graph():
    """The loss kernels allocate only through the caching allocator and read no host state: capturable."""
    from layoutdetr_b200 import box_ops
    bbox, mask, _ = _case(5)
    b = bbox.cuda()
    m = mask.cuda()
    with torch.no_grad():
        eager = torch.stack(box_ops.layout_losses(b, m))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        box_ops.layout_losses(b, m)
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr), torch.no_grad():
        out = torch.stack(box_ops.layout_losses(b, m))
    gr.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
