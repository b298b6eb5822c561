"""endodac image model on a B200 (drop-in class -> ctypes -> C ABI, edv_config.no_motion) against golden
vectors written by the UNMODIFIED reference (models/endodac/endodac.py).  Same gates as test_gpu_forward.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import endodav_b200 as E  # noqa: E402
from oracle import endodav_oracle as orc  # noqa: E402
from oracle import weights  # noqa: E402
from golden_util import load_case, manifest  # noqa: E402
from test_endodac_cpu import DAC, dac_cfg, dac_input  # noqa: E402
from test_gpu_forward import FP32_RTOL, GATES, _metrics, _rel  # noqa: E402


def _build(m, dtype):
    ctor = dict(m["ctor"])
    ctor["image_shape"] = tuple(ctor["image_shape"])
    model = E.endodac(dtype=dtype, **ctor)
    model.load_state_dict(weights.to_endodac_keys(weights.make_state_dict(dac_cfg(ctor), m["weight_seed"])), strict=True)
    return model.cuda().eval()


@pytest.mark.parametrize("name", DAC)
def test_endodac_fp32_matches_reference_golden(name):
    m, arrays = load_case(name)
    model = _build(m, "fp32")
    out = model(dac_input(m).cuda())
    assert model._eng.launch_count() > 0
    for s in range(4):
        got, ref = out[("disp", s)].cpu().numpy(), arrays["disp%d" % s]
        assert got.shape == ref.shape
        err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
        assert err.max() <= FP32_RTOL, (name, s, float(err.max()))


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("name", DAC)
def test_endodac_16bit_within_tolerance(name, dtype):
    m, arrays = load_case(name)
    model = _build(m, dtype)
    out = model(dac_input(m).cuda())
    got, ref = out[("disp", 0)].cpu().numpy(), arrays["disp0"]
    gate = GATES[dtype]
    absrel, a1 = _metrics(got, ref)
    assert _rel(got, ref).max() <= gate["rel"], (name, dtype, float(_rel(got, ref).max()))
    assert absrel <= gate["absrel"] and a1 >= gate["a1"], (name, dtype, absrel, a1)
    for s in range(1, 4):
        assert _rel(out[("disp", s)].cpu().numpy(), arrays["disp%d" % s]).max() <= gate["rel"], (name, dtype, s)


@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-4), ("fp16", 1e-2)])
def test_endodac_infer_video_depth_matches_reference_golden(dtype, tol):
    m, arrays = load_case("dac_video_n11")
    model = _build(m, dtype)
    N, H, W = m["input"]
    v = weights.make_video_u8(N, H, W, m["frame_seed"])
    got = model.infer_video_depth(v, batch_size=m["batch_size"])
    ref = arrays["depth"]
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert _rel(got, ref).max() <= tol, float(_rel(got, ref).max())


def test_endodac_base_full_size_batch():
    """ViT-B at 518x518, 8 frames: runs on the tensor-core path, frames are independent (no temporal mixing)."""
    ctor = dict(backbone_size="base", lora_type="dvlora", image_shape=(518, 518), disable_conv_head=True)
    model = E.endodac(dtype="fp16", **ctor)
    model.load_state_dict(weights.to_endodac_keys(weights.make_state_dict(dac_cfg(ctor), 5)), strict=True)
    model = model.cuda().eval()
    x = weights.make_frames(1, 8, 518, 518, 6)[0].cuda()
    a = model(x)[("disp", 0)]
    assert a.shape == (8, 1, 518, 518) and torch.isfinite(a).all() and float(a.mean()) > 0.05
    b = model(x.flip(0))[("disp", 0)].flip(0)
    assert torch.equal(a, b)
    # one frame against the CPU oracle
    sd = weights.to_endodac_keys(weights.make_state_dict(dac_cfg(ctor), 5))
    ref = orc.forward_endodac(sd, x[:1].cpu(), dac_cfg(ctor), (518, 518))[("disp", 0)].numpy()
    assert _rel(a[:1].cpu().numpy(), ref).max() <= 1e-2
