"""endodac image model on CPU: the oracle restatement against golden vectors written
by the UNMODIFIED reference (models/endodac/endodac.py, oracle/make_golden.py --endodac-only), the
drop-in class's checkpoint layout, and loud failure without a GPU."""
import json
import os

import numpy as np
import pytest
import torch

import endodav_b200 as E
from oracle import endodav_oracle as orc
from oracle import weights
from golden_util import GOLDEN_DIR, load_case, manifest

ATOL = 5e-5
DAC = [k for k, v in manifest().items() if v["kind"] == "endodac_forward"]


def dac_cfg(ctor):
    return weights.endodac_cfg(ctor["backbone_size"], ctor.get("lora_type", "lora"), 4,
                               ctor.get("residual_block_indexes", []), ctor.get("disable_conv_head", False),
                               ctor.get("include_cls_token", True), ctor.get("use_cls_token", False), ctor.get("use_bn", False))


def dac_input(m):
    shape = m["input"]
    n = int(np.prod(shape[:-3]))
    return weights.make_frames(1, n, shape[-2], shape[-1], m["frame_seed"])[0].reshape(shape)


@pytest.mark.parametrize("name", DAC)
def test_endodac_oracle_matches_reference_golden(name):
    m, arrays = load_case(name)
    ctor = m["ctor"]
    cfg = dac_cfg(ctor)
    sd = weights.to_endodac_keys(weights.make_state_dict(cfg, m["weight_seed"]))
    assert len(sd) == m["keys"]
    out = orc.forward_endodac(sd, dac_input(m), cfg, tuple(ctor["image_shape"]), pre_norm=ctor.get("pre_norm", False),
                              inv_sigmoid=ctor.get("inv_sigmoid", False))
    for s in range(4):
        got, ref = out[("disp", s)].numpy(), arrays["disp%d" % s]
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= ATOL, (name, s, float(np.abs(got - ref).max()))
    assert float(np.abs(arrays["disp0"]).mean()) > 0.05


@pytest.mark.parametrize("name", DAC)
def test_endodac_packed_graph_matches_reference_golden(name):
    """pack.py (lora_scale 1/r, no motion weights) + the engine's graph algebra with edv_config.no_motion /
    no_normalize / taps = last four blocks reproduce the reference endodac outputs (fp32, CPU emulation)."""
    from endodav_b200 import pack
    import packed_emulator as emu

    m, arrays = load_case(name)
    ctor = m["ctor"]
    cfg = dac_cfg(ctor)
    sd = weights.make_state_dict(cfg, m["weight_seed"])          # endodav-layout keys, as model._pack_state_dict yields
    ishape = tuple(ctor["image_shape"])
    pk = pack.pack_state_dict(sd, cfg, torch.float32)
    assert not any(k.startswith("mm") for k in pk)
    pk.update(pack.pos_tables(sd, cfg, ishape[0] // 14, ishape[1] // 14))
    enc = dict(pack.ENCODERS[cfg["encoder"]], taps=cfg["taps"])
    x = dac_input(m)
    x = x.flatten(0, 1) if x.dim() == 5 else x
    ecfg = dict(cfg, normalize=ctor.get("pre_norm", False), inv_sigmoid=ctor.get("inv_sigmoid", False))
    with torch.no_grad():
        out = emu.forward(pk, ecfg, enc, x[None], ishape)
    for s in range(4):
        got, ref = out[("disp", s)].numpy(), arrays["disp%d" % s]
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-4, (name, s, float(np.abs(got - ref).max()))


def test_endodac_video_oracle_matches_reference_golden():
    m, arrays = load_case("dac_video_n11")
    ctor = m["ctor"]
    cfg = dac_cfg(ctor)
    sd = weights.to_endodac_keys(weights.make_state_dict(cfg, m["weight_seed"]))
    N, H, W = m["input"]
    v = weights.make_video_u8(N, H, W, m["frame_seed"])
    got = orc.infer_video_depth_endodac(sd, v, cfg, tuple(ctor["image_shape"]), batch_size=m["batch_size"])
    assert got.dtype == np.float32 and got.shape == arrays["depth"].shape
    assert np.abs(got - arrays["depth"]).max() <= ATOL


@pytest.mark.parametrize("name", ["dac_small_dvlora", "dac_base_lora_convhead"])
def test_endodac_state_dict_layout_matches_reference(name):
    with open(os.path.join(GOLDEN_DIR, "state_dict_keys_%s.json" % name)) as f:
        ref = json.load(f)
    ctor = dict(manifest()[name]["ctor"])
    ctor["image_shape"] = tuple(ctor["image_shape"])
    model = E.endodac(**ctor)
    sd = model.state_dict()
    assert list(sd.keys()) == [k for k, _ in ref]
    for k, shape in ref:
        assert list(sd[k].shape) == shape, k
    model.load_state_dict(weights.to_endodac_keys(weights.make_state_dict(dac_cfg(ctor), 1)), strict=True)
    assert hasattr(model, "pretrained") and hasattr(model, "depth_head") and not hasattr(model, "head")


def test_endodac_ctor_defaults_and_errors():
    m = E.endodac()                                     # reference defaults: base, lora, conv head (endodac.py:153-165)
    assert m.backbone_size == "base" and m.embedding_dim == 768 and m.image_shape == (224, 280)
    assert "depth_head.conv_depth_1.head.0.weight" in m.state_dict()
    with pytest.raises(KeyError):
        E.endodac(backbone_size="large")
    with pytest.raises(AssertionError):
        E.endodac(r=0)


def test_endodac_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    model = E.endodac(backbone_size="small", image_shape=(28, 28), disable_conv_head=True)
    with pytest.raises(E.EndoDAVError):
        model(torch.rand(2, 3, 28, 28))
    with pytest.raises(E.EndoDAVError):
        model.infer_video_depth(np.zeros((3, 28, 28, 3), np.uint8))
