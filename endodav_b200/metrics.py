"""GPU versions of the two evaluation helpers the reference's evaluate scripts apply to the output of
``infer_video_depth``, with the reference's own signatures:

  * ``disp_to_depth(disp, min_depth, max_depth)``   -- utils/layers.py:11-20
  * ``compute_errors(gt, pred, mask=None)``         -- utils/utils.py:112-133
  * ``evaluate_frames(...)``                        -- the per-frame masking / scaling / clamping loop around
                                                       compute_errors (evaluate_depth_video.py:197-204), batched

They accept numpy arrays (returned as numpy, like the reference) or CUDA tensors (returned as tensors, nothing leaves
the device).  The arithmetic runs in the library's CUDA kernels (csrc/metrics.cuh); there is no CPU fallback."""
import ctypes

import numpy as np
import torch

from . import engine as _engine


def _dev(x, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        if x.device.type != "cuda":
            x = x.cuda()
        return x.to(dtype).contiguous(), False
    return torch.from_numpy(np.ascontiguousarray(x)).to(device="cuda", dtype=dtype), True


def disp_to_depth(disp, min_depth, max_depth):
    """-> (scaled_disp, depth), bit-identical to the reference's float32 numpy expression."""
    lib = _engine.load_library()
    d, as_np = _dev(disp)
    scaled, depth = torch.empty_like(d), torch.empty_like(d)
    _engine._check(lib.edv_op_disp_to_depth(_engine._ptr(d), _engine._ptr(scaled), _engine._ptr(depth), d.numel(),
                                            float(min_depth), float(max_depth), _engine._stream()), None, "edv_op_disp_to_depth")
    if as_np:
        return scaled.cpu().numpy(), depth.cpu().numpy()
    return scaled, depth


def _errors(gt, pred, mask, gt_lo, gt_hi, scale, clamp_lo, clamp_hi):
    lib = _engine.load_library()
    g, as_np = _dev(gt)
    p, _ = _dev(pred)
    if g.shape != p.shape:
        raise ValueError("gt and pred must have the same shape, got %s and %s" % (tuple(g.shape), tuple(p.shape)))
    frames = 1 if g.dim() <= 2 else int(np.prod(g.shape[:-2]))
    hw = g.numel() // frames
    m = None
    if mask is not None:
        m, _ = _dev(mask, torch.uint8)
        if m.numel() != g.numel():
            raise ValueError("mask must have the shape of gt")
    out = torch.empty(frames, 8, dtype=torch.float64, device=g.device)
    _engine._check(lib.edv_op_compute_errors(_engine._ptr(g), _engine._ptr(p), _engine._ptr(m), frames, hw,
                                             ctypes.c_float(gt_lo), ctypes.c_float(gt_hi), ctypes.c_float(scale),
                                             ctypes.c_float(clamp_lo), ctypes.c_float(clamp_hi), _engine._ptr(out),
                                             _engine._stream()), None, "edv_op_compute_errors")
    return out, as_np


def compute_errors(gt, pred, mask=None):
    """abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 over the masked pixels of ONE frame (reference signature)."""
    out, _ = _errors(gt, pred, mask if mask is not None else np.ones(np.shape(gt), dtype=np.uint8), 0.0, 0.0, 1.0, 1.0, 0.0)
    return tuple(float(v) for v in out[0, :7].cpu())


def evaluate_frames(gt_depths, pred_depths, min_depth=1e-3, max_depth=150.0, pred_depth_scale_factor=1.0):
    """evaluate_depth_video.py:197-204 for all frames at once: valid = gt in (min_depth, max_depth), pred scaled and
    clamped to [min_depth, max_depth], then compute_errors.  -> float64 [N,7] (+ the valid-pixel count as column 7)."""
    out, as_np = _errors(gt_depths, pred_depths, None, min_depth, max_depth, pred_depth_scale_factor, min_depth, max_depth)
    return out.cpu().numpy() if as_np else out
