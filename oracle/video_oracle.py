"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the sliding-window video driver.

Restates ``endodav.infer_video_depth`` (models/endodav/endodav.py:162-254) and its helpers
``compute_scale_and_shift_full`` / ``get_interpolate_frames`` (utils/util.py:40-74) and the
aspect-keeping ``Resize`` size rule (models/endodav/util/transform.py:49-107), with the
network abstracted as ``forward_fn(clip[1,32,3,h,w] float32) -> disp[32,1,h0,w0]``.

Integer/index work (window starts, padding, keyframe slots) must be bit-exact against the
reference; the float stitching follows the reference's float32 numpy arithmetic op by op.
"""
import numpy as np

INFER_LEN = 32  # endodav.py:47-50
OVERLAP = 10
KEYFRAMES = [6, 12, 24, 25, 26, 27, 28, 29, 30, 31]
INTERP_LEN = 8
STEP = INFER_LEN - OVERLAP


def resize_target(width, height, want_w, want_h, multiple=14):
    """Resize.get_size with keep_aspect_ratio=True, resize_method='lower_bound'
    (util/transform.py:62-107).  Returns (new_w, new_h)."""
    sh, sw = want_h / height, want_w / width
    if sw > sh:
        sh = sw
    else:
        sw = sh

    def constrain(x, min_val):
        y = int(np.round(x / multiple) * multiple)
        if y < min_val:
            y = int(np.ceil(x / multiple) * multiple)
        return y

    return constrain(sw * width, want_w), constrain(sh * height, want_h)


def window_slots(n_frames):
    """For every window k: the list of 32 source-frame indices it reads.

    Follows endodav.py:185-199 literally: pad the frame list with copies of the last
    frame, slide with stride 22, then overwrite slots 0..9 by the *previous window's
    input* slots KEYFRAMES."""
    pad = (STEP - (n_frames % STEP)) % STEP + (INFER_LEN - STEP)
    src = list(range(n_frames)) + [n_frames - 1] * pad
    out, prev = [], None
    for start in range(0, n_frames, STEP):
        cur = [src[start + i] for i in range(INFER_LEN)]
        if prev is not None:
            for i in range(OVERLAP):
                cur[i] = prev[KEYFRAMES[i]]
        out.append(cur)
        prev = cur
    return out


def scale_and_shift(pred, target):
    """compute_scale_and_shift_full with an all-ones mask (utils/util.py:40-62)."""
    pred = pred.astype(np.float32)
    target = target.astype(np.float32)
    mask = np.ones_like(target, dtype=np.float32)
    a00 = np.sum(mask * pred * pred)
    a01 = np.sum(mask * pred)
    a11 = np.sum(mask)
    b0 = np.sum(mask * pred * target)
    b1 = np.sum(mask * target)
    x0, x1 = 1, 0
    det = a00 * a11 - a01 * a01
    if det != 0:
        x0 = (a11 * b0 - a01 * b1) / det
        x1 = (-a01 * b0 + a00 * b1) / det
    return x0, x1


def stitch(per_window, n_frames):
    """per_window: list over windows of float32 [32,H,W] -> float32 [n_frames,H,W]
    (endodav.py:213-254; the dead ``ref_align`` bookkeeping is omitted)."""
    aligned = []
    for k, win in enumerate(per_window):
        frames = [win[i] for i in range(INFER_LEN)]
        if k == 0:
            aligned += frames
            continue
        pre = aligned[-INTERP_LEN:]
        post = frames[OVERLAP - INTERP_LEN:OVERLAP]
        scale, shift = scale_and_shift(np.concatenate(post), np.concatenate(pre))
        post2 = []
        for f in post:
            g = f * scale + shift
            g[g < 0] = 0
            post2.append(g)
        step = 1.0 / (INTERP_LEN - 1)
        w = [0.0] + [i * step for i in range(1, INTERP_LEN - 1)] + [1.0]
        aligned[-INTERP_LEN:] = [pre[i] * (1 - w[i]) + post2[i] * w[i] for i in range(INTERP_LEN)]
        for i in range(OVERLAP, INFER_LEN):
            g = frames[i] * scale + shift
            g[g < 0] = 0
            aligned.append(g)
    return np.stack(aligned[:n_frames], axis=0)


def infer_video_depth(frames_u8, image_shape, forward_fn):
    """frames_u8 [N,H,W,3] uint8 -> float32 [N,H,W].  ``forward_fn`` receives the window
    as a float32 torch tensor [1,32,3,h,w] and returns disparity [32,1,h0,w0]."""
    import cv2
    import torch
    import torch.nn.functional as F

    n, H, W = frames_u8.shape[:3]
    new_w, new_h = resize_target(W, H, image_shape[1], image_shape[0])
    resized = {}

    def prep(idx):
        if idx not in resized:
            img = frames_u8[idx].astype(np.float32) / 255.0
            img = cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_CUBIC)
            resized[idx] = np.ascontiguousarray(np.transpose(img, (2, 0, 1))).astype(np.float32)
        return resized[idx]

    per_window = []
    for slots in window_slots(n):
        clip = torch.from_numpy(np.stack([prep(i) for i in slots], 0)).unsqueeze(0)
        disp = forward_fn(clip)
        disp = F.interpolate(disp.reshape(INFER_LEN, 1, *disp.shape[-2:]).float(), size=(H, W),
                             mode="bilinear", align_corners=True)
        per_window.append(disp[:, 0].cpu().numpy())
    return stitch(per_window, n)
