"""Hungarian matching: the C oracle is pinned to scipy's goldens (CPU); the CUDA kernel must match both bit-exactly (GPU)."""
import numpy as np
import pytest
import torch

from helpers import golden


def test_lsap_oracle_matches_scipy_goldens():
    from oracle import lsap
    cases = golden("hungarian_scipy.pt")
    for c in cases:
        r, col = lsap.linear_sum_assignment(c["cost"].numpy(), maximize=True)
        assert np.array_equal(r, c["row"].numpy()) and np.array_equal(col, c["col"].numpy()), c["cost"]


def test_lsap_oracle_matches_installed_scipy_random():
    scipy_opt = pytest.importorskip("scipy.optimize")
    from oracle import lsap
    rng = np.random.RandomState(3)
    for _ in range(300):
        nr, nc = rng.randint(1, 10), rng.randint(1, 10)
        m = rng.rand(nr, nc)
        if rng.rand() < 0.3:
            m = np.round(m * 2) / 2
        for mx in (True, False):
            r0, c0 = scipy_opt.linear_sum_assignment(m, maximize=mx)
            r1, c1 = lsap.linear_sum_assignment(m, maximize=mx)
            assert np.array_equal(r0, r1) and np.array_equal(c0, c1)


@pytest.mark.gpu
def test_lsap_cuda_bit_exact():
    from layoutdetr_b200 import kernels as K
    from oracle import lsap
    cases = golden("hungarian_scipy.pt")
    for c in cases:
        r, col, st = K.lsap(c["cost"].cuda(), maximize=True)
        assert int(st) == 0
        assert torch.equal(r.cpu(), c["row"]) and torch.equal(col.cpu(), c["col"]), c["cost"]
    rng = np.random.RandomState(11)
    for (nr, nc) in [(9, 9), (5, 8), (8, 5), (1, 1), (16, 16)]:
        m = np.round(rng.rand(500, nr, nc) * 4) / 4
        r, col, st = K.lsap(torch.from_numpy(m).cuda(), maximize=True)
        assert int(st.abs().sum()) == 0
        for i in range(0, 500, 7):
            r0, c0 = lsap.linear_sum_assignment(m[i], maximize=True)
            assert np.array_equal(r[i].cpu().numpy(), r0) and np.array_equal(col[i].cpu().numpy(), c0)


def test_maximum_iou_oracle_matches_reference_golden():
    """oracle restatement of compute_maximum_iou vs the value the reference's own function produced (gen_golden.py --only-maxiou)."""
    from oracle import layoutdetr_oracle as O
    g = golden("maxiou_ref.pt")
    l1 = [(b.numpy(), l.numpy()) for b, l in g["layouts_1"]]
    l2 = [(b.numpy(), l.numpy()) for b, l in g["layouts_2"]]
    assert abs(O.compute_maximum_iou(l1, l2) - g["score"]) < 1e-12


@pytest.mark.gpu
def test_maximum_iou_cuda_matches_reference_golden():
    """Product compute_maximum_iou (batched ld_lsap) vs the reference's value: assignments are bit-identical to scipy, the score
    differs only by fp64 summation order."""
    from layoutdetr_b200.metrics.metric_layoutnet import compute_maximum_iou
    g = golden("maxiou_ref.pt")
    l1 = [(b.numpy(), l.numpy()) for b, l in g["layouts_1"]]
    l2 = [(b.numpy(), l.numpy()) for b, l in g["layouts_2"]]
    assert abs(compute_maximum_iou(l1, l2) - g["score"]) < 1e-9
