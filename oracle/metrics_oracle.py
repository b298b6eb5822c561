"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the evaluation helpers the reference's evaluate scripts apply
to the output of ``infer_video_depth``.  Pinned by tests/golden/metrics.npz, generated from the UNMODIFIED reference
functions (oracle/make_golden.py --metrics-only).  The product never imports this file."""
import numpy as np


def disp_to_depth(disp, min_depth, max_depth):
    """utils/layers.py:11-20."""
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    scaled_disp = min_disp + (max_disp - min_disp) * disp
    depth = 1 / scaled_disp
    return scaled_disp, depth


def compute_errors(gt, pred, mask=None):
    """utils/utils.py:112-133."""
    if mask is not None:
        pred = pred[mask]
        gt = gt[mask]
    thresh = np.maximum((gt / pred), (pred / gt))
    a1 = (thresh < 1.25).mean()
    a2 = (thresh < 1.25 ** 2).mean()
    a3 = (thresh < 1.25 ** 3).mean()
    rmse = np.sqrt(((gt - pred) ** 2).mean())
    rmse_log = np.sqrt(((np.log(gt) - np.log(pred)) ** 2).mean())
    abs_rel = np.mean(np.abs(gt - pred) / gt)
    sq_rel = np.mean(((gt - pred) ** 2) / gt)
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3


def evaluate_frames(gt_depths, pred_depths, min_depth=1e-3, max_depth=150, pred_depth_scale_factor=1.0):
    """The per-frame loop of evaluate_depth_video.py:197-204 (mask, scale, clamp, compute_errors) -> [N,7]."""
    out = []
    for gt_depth, pred_depth in zip(gt_depths, pred_depths):
        pred_depth = pred_depth.copy()
        valid_mask = np.logical_and(gt_depth > min_depth, gt_depth < max_depth)
        pred_depth *= pred_depth_scale_factor
        pred_depth[pred_depth < min_depth] = min_depth
        pred_depth[pred_depth > max_depth] = max_depth
        out.append(compute_errors(gt_depth, pred_depth, valid_mask))
    return np.array(out, dtype=np.float64)


def make_case(seed=5, n=6, h=48, w=64):
    """Seeded synthetic disparity / ground-truth pair shared by the fixture generator and the tests."""
    rng = np.random.default_rng(seed)
    disp = rng.random((n, h, w), dtype=np.float32) * 1.5
    gt = (rng.random((n, h, w), dtype=np.float32) * 0.2 + 0.01).astype(np.float32)
    gt[rng.random((n, h, w)) < 0.2] = 0.0          # invalid ground-truth pixels
    gt[rng.random((n, h, w)) < 0.02] = 200.0       # beyond MAX_DEPTH
    gt[n - 1] = 0.0                                # a frame without any valid pixel
    return disp, gt
