"""On-GPU window stitching (edv_op_stitch_window) against the reference's numpy chain
(video.stitch_windows == endodav.py:213-254, itself pinned bit-exactly by tests/golden/video_stub.npz)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from endodav_b200 import engine as eng  # noqa: E402
from endodav_b200 import video as V  # noqa: E402
from oracle import weights  # noqa: E402
from golden_util import load_case  # noqa: E402


def _windows(nwin, H, W, seed, spread=0.3):
    rng = np.random.default_rng(seed)
    base = rng.random((H, W), dtype=np.float32) * spread + 0.5
    wins = []
    for k in range(nwin):
        a, b = 0.7 + 0.6 * rng.random(), 0.2 * rng.random() - 0.1
        w = (base[None] * (1 + 0.3 * np.sin(np.arange(32, dtype=np.float32) * 0.3 + k))[:, None, None]
             + 0.05 * rng.standard_normal((32, H, W)).astype(np.float32))
        wins.append((a * w + b).astype(np.float32))
    return wins


def ref_tail(wins, k):
    """depth_list_aligned[-8:] right before window k is stitched (numpy chain)."""
    seq = V.stitch_windows(wins[:k], 32 + 22 * (k - 1))
    return seq[-8:]


def _gpu_stitch(wins, n):
    H, W = wins[0].shape[1:]
    st = V._GpuStitcher(len(wins), n, H, W, torch.device("cuda"))
    for w in wins:
        st.push(torch.from_numpy(w).cuda())
    out = st.finish()
    return out, st.scale_shift.cpu().numpy()


@pytest.mark.parametrize("n,H,W", [(32, 8, 12), (33, 7, 9), (54, 30, 44), (100, 64, 80), (120, 37, 53), (230, 256, 320)])
def test_stitch_matches_numpy_chain(n, H, W):
    nwin = V.num_windows(n)
    wins = _windows(nwin, H, W, n)
    ref = V.stitch_windows(wins, n)
    got, ss = _gpu_stitch(wins, n)
    assert got.shape == ref.shape and got.dtype == np.float32
    # bit-exact: numpy's pairwise sums are evaluated in the same association order (edv_op_stitch_plan) and
    # the float32 solve / clamp / cross-fade op for op without FMA contraction
    assert np.array_equal(got, ref), float(np.abs(got - ref).max())
    assert ss[0, 0] == 1.0 and ss[0, 1] == 0.0
    for k in range(1, nwin):
        pos = 32 + 22 * (k - 1)
        sc, sh = V.lsq_scale_shift(np.concatenate(list(wins[k][2:10])), np.concatenate(list(ref_tail(wins, k))))
        assert np.float32(sc) == ss[k, 0] and np.float32(sh) == ss[k, 1], (k, sc, sh, ss[k])


def test_stitch_given_same_scale_shift_is_bit_exact():
    """Feeding windows that are already aligned (scale 1, shift 0 solves exactly when post == pre) the apply
    step must reproduce numpy bit for bit: float32 multiply / add without FMA contraction, clamp, cross-fade."""
    H, W, nwin = 16, 20, 4
    rng = np.random.default_rng(3)
    wins = [rng.random((32, H, W), dtype=np.float32) for _ in range(nwin)]
    ref_scales = []
    # make every solve degenerate (det == 0 -> scale 1, shift 0): constant overlap frames
    for w in wins:
        w[2:10] = 0.5
    wins[0][24:32] = 0.5
    for k in range(1, nwin):
        wins[k][24:32] = 0.5
    ref = V.stitch_windows(wins, 32 + 22 * (nwin - 1))
    got, ss = _gpu_stitch(wins, 32 + 22 * (nwin - 1))
    assert np.array_equal(ss, np.tile(np.array([[1.0, 0.0]], np.float32), (nwin, 1)))
    assert np.array_equal(got, ref)


def test_negative_values_are_clamped():
    H, W = 8, 8
    wins = _windows(3, H, W, 5)
    wins[1][12:20] -= 5.0                      # fresh frames far below zero after alignment
    ref = V.stitch_windows(wins, 76)
    got, _ = _gpu_stitch(wins, 76)
    assert (got >= 0).all() and (ref[40:48] == 0).any()
    assert np.array_equal(got, ref)


def test_infer_video_depth_gpu_stitch_equals_host_stitch(monkeypatch):
    import endodav_b200 as E
    from golden_util import oracle_cfg

    m, arrays = load_case("video_n45")
    kw = dict(m["ctor"])
    kw["image_shape"] = tuple(kw["image_shape"])
    model = E.endodav(dtype="fp32", **kw)
    model.load_state_dict(weights.make_state_dict(oracle_cfg(m["ctor"]), m["weight_seed"]), strict=True)
    model = model.cuda().eval()
    N, H, W = m["input"]
    v = weights.make_video_u8(N, H, W, m["frame_seed"])
    a = model.infer_video_depth(v)
    monkeypatch.setenv("ENDODAV_STITCH", "host")
    b = model.infer_video_depth(v)
    assert a.shape == b.shape == arrays["depth"].shape
    assert np.array_equal(a, b)


@pytest.mark.parametrize("seg_bytes", [1, 22 * 30 * 44 * 4 * 3])
def test_stitch_bounded_device_segment_is_bit_exact(monkeypatch, seg_bytes):
    """The stitched sequence lives on the GPU only as a sliding segment (long HD videos must not keep
    N*H*W*4 bytes on the device).  Forcing the smallest segments (1 and 3 windows) must not change a single bit."""
    n, H, W = 230, 30, 44
    nwin = V.num_windows(n)
    wins = _windows(nwin, H, W, 17)
    ref = V.stitch_windows(wins, n)
    monkeypatch.setattr(V._GpuStitcher, "SEGMENT_BYTES", seg_bytes)
    got, _ = _gpu_stitch(wins, n)
    assert np.array_equal(got, ref)


def test_large_frames_fall_back_to_host_stitching():
    """8*H*W >= 2^24 (2048 x 1024 frames): edv_op_stitch_plan refuses the shape and infer_video_depth must run the
    reference's numpy chain instead of raising."""
    import endodav_b200 as E

    assert V._stitch_plan_or_none(1024, 2048) is None and V._stitch_plan_or_none(256, 320) is not None
    kw = dict(encoder="vits", features=64, out_channels=[48, 96, 192, 384], r=4, lora_type="dvlora",
              image_shape=(28, 56), disable_conv_head=True, residual_block_indexes=[])
    model = E.endodav(dtype="fp16", **kw)
    from endodav_b200 import synthetic
    synthetic.randomize_(model, 5)
    model = model.cuda().eval()
    v = weights.make_video_u8(24, 1024, 2048, 3)
    out = model.infer_video_depth(v)
    assert out.shape == (24, 1024, 2048) and out.dtype == np.float32 and np.isfinite(out).all()


def test_non_uint8_frames_take_the_host_preprocessing_path():
    """The reference accepts any frame dtype through .astype(np.float32)/255 (endodav.py:195)."""
    import endodav_b200 as E
    from golden_util import oracle_cfg

    m, arrays = load_case("video_n5")
    kw = dict(m["ctor"])
    kw["image_shape"] = tuple(kw["image_shape"])
    model = E.endodav(dtype="fp32", **kw)
    model.load_state_dict(weights.make_state_dict(oracle_cfg(m["ctor"]), m["weight_seed"]), strict=True)
    model = model.cuda().eval()
    N, H, W = m["input"]
    v = weights.make_video_u8(N, H, W, m["frame_seed"])
    a = model.infer_video_depth(v.astype(np.int32))
    err = np.abs(a - arrays["depth"]) / np.maximum(np.abs(arrays["depth"]), 1.0)
    assert err.max() <= 5e-4


def test_data_writes_need_invalidate_weights():
    """writes through .data bump neither the version counter nor the pointer; invalidate_weights()
    forces the re-pack, in-place ops and load_state_dict trigger it by themselves."""
    import endodav_b200 as E
    from endodav_b200 import synthetic

    kw = dict(encoder="vits", features=64, out_channels=[48, 96, 192, 384], r=4, lora_type="dvlora",
              image_shape=(28, 42), disable_conv_head=True, residual_block_indexes=[])
    model = E.endodav(dtype="fp16", **kw)
    synthetic.randomize_(model, 7)
    model = model.cuda().eval()
    x = weights.make_frames(1, 2, 28, 42, 3).cuda()
    a = model(x)[("disp", 0)].clone()
    p = dict(model.named_parameters())["head.scratch.output_conv2.0.bias"]
    p.data.add_(0.25)                                   # invisible to the automatic check
    model.invalidate_weights()
    b = model(x)[("disp", 0)].clone()
    assert not torch.equal(a, b)
    with torch.no_grad():
        p.add_(0.25)                                    # bumps _version: picked up automatically
    c = model(x)[("disp", 0)]
    assert not torch.equal(b, c)
