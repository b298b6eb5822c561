"""disp_to_depth / compute_errors: the numpy oracle against the fixture written by the UNMODIFIED
reference functions (CPU), and the CUDA kernels against both (GPU, through the C ABI)."""
import os
import warnings

import numpy as np
import pytest

from oracle import metrics_oracle as mo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz")


def _gold():
    return dict(np.load(GOLD))


def test_oracle_matches_reference_fixture():
    g = _gold()
    disp, gt = mo.make_case()
    scaled, depth = mo.disp_to_depth(disp, 0.1, 150.0)
    assert scaled.dtype == np.float32 and np.array_equal(scaled, g["scaled"]) and np.array_equal(depth, g["depth"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        errs = mo.evaluate_frames(gt, depth, 1e-3, 150, 1.0)
    assert np.array_equal(np.isnan(errs), np.isnan(g["errors"]))
    assert np.isnan(errs[-1]).all()                      # the frame without valid ground truth
    np.testing.assert_allclose(errs[:-1], g["errors"][:-1], rtol=0, atol=0)


@pytest.mark.gpu
def test_gpu_disp_to_depth_is_bit_exact():
    from endodav_b200 import metrics

    g = _gold()
    disp, _ = mo.make_case()
    scaled, depth = metrics.disp_to_depth(disp, 0.1, 150.0)
    assert scaled.dtype == np.float32 and depth.dtype == np.float32
    assert np.array_equal(scaled, g["scaled"]) and np.array_equal(depth, g["depth"])
    for lo, hi in ((1e-3, 150.0), (0.1, 100.0), (0.5, 80.0)):
        s2, d2 = metrics.disp_to_depth(disp, lo, hi)
        rs, rd = mo.disp_to_depth(disp, lo, hi)
        assert np.array_equal(s2, rs) and np.array_equal(d2, rd), (lo, hi)


@pytest.mark.gpu
def test_gpu_compute_errors_matches_reference():
    import torch

    from endodav_b200 import metrics

    g = _gold()
    disp, gt = mo.make_case()
    depth = g["depth"]
    out = metrics.evaluate_frames(gt, depth, 1e-3, 150.0, 1.0)
    assert out.shape == (gt.shape[0], 8)
    assert np.isnan(out[-1, :7]).all() and out[-1, 7] == 0
    # numpy reduces float32 arrays pairwise in float32, the kernel in float64: agreement to float32 round-off
    np.testing.assert_allclose(out[:-1, :7], g["errors"][:-1], rtol=2e-6)
    valid = np.logical_and(gt > 1e-3, gt < 150)
    assert np.array_equal(out[:, 7], valid.reshape(gt.shape[0], -1).sum(1))
    # single-frame reference signature with an explicit mask, and device tensors in / out
    p = np.clip(depth[0], 1e-3, 150)
    one = metrics.compute_errors(gt[0], p, valid[0])
    np.testing.assert_allclose(one, g["errors"][0], rtol=2e-6)
    dev = metrics.evaluate_frames(torch.from_numpy(gt).cuda(), torch.from_numpy(depth).cuda(), 1e-3, 150.0, 2.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref2 = mo.evaluate_frames(gt, depth, 1e-3, 150, 2.0)
    assert dev.is_cuda
    np.testing.assert_allclose(dev[:-1, :7].cpu().numpy(), ref2[:-1], rtol=2e-6)
    # determinism (fixed summation order)
    again = metrics.evaluate_frames(gt, depth, 1e-3, 150.0, 1.0)
    assert np.array_equal(out[:-1], again[:-1])
