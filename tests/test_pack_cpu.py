"""Host-side logic on CPU: checkpoint layout of the drop-in class, the weight packer and the
engine's graph algebra (emulated in torch on the packed weights) against the oracle."""
import json
import os

import numpy as np
import pytest
import torch

import endodav_b200 as E
from endodav_b200 import pack
from oracle import endodav_oracle as orc
from oracle import weights
from golden_util import GOLDEN_DIR, load_case, manifest, oracle_cfg, subsample_like_golden
import packed_emulator as emu

ENC = {"vits": dict(dim=384, depth=12, heads=6, taps=[2, 5, 8, 11]), "vitl": dict(dim=1024, depth=24, heads=16, taps=[4, 11, 17, 23])}


def _ctor(m):
    kw = dict(m["ctor"])
    kw["image_shape"] = tuple(kw["image_shape"])
    return kw


@pytest.mark.parametrize("keys_file,name", [("state_dict_keys_vits.json", "fwd_vits_dvlora"),
                                            ("state_dict_keys_vits_lora_res_convhead.json", "fwd_vits_lora_res_convhead"),
                                            ("state_dict_keys_vitl.json", "fwd_vitl")])
def test_state_dict_layout_matches_reference(keys_file, name):
    """Keys, order and shapes equal the reference's state_dict() (captured by make_golden.py)."""
    with open(os.path.join(GOLDEN_DIR, keys_file)) as f:
        ref = json.load(f)
    model = E.endodav(**_ctor(manifest()[name]))
    sd = model.state_dict()
    assert list(sd.keys()) == [k for k, _ in ref]
    for k, shape in ref:
        assert list(sd[k].shape) == shape, k
    # the synthetic checkpoint loads strictly, like a real one would
    cfg = oracle_cfg(manifest()[name]["ctor"])
    model.load_state_dict(weights.make_state_dict(cfg, 1), strict=True)
    assert isinstance(model.head.motion_modules, torch.nn.ModuleList) and len(model.head.motion_modules) == 4
    assert hasattr(model, "pretrained") and hasattr(model, "head")


def test_filtered_load_idiom():
    """evaluate_depth_video.py:91-93: load_state_dict({k: v for k in ckpt if k in model_dict})."""
    model = E.endodav(encoder="vits", features=64, out_channels=[48, 96, 192, 384], lora_type="dvlora", disable_conv_head=True)
    ckpt = weights.make_state_dict(weights.full_cfg(), 5)
    ckpt.update({"height": torch.tensor(256), "width": torch.tensor(320), "use_stereo": torch.tensor(False)})
    model_dict = model.state_dict()
    model.load_state_dict({k: v for k, v in ckpt.items() if k in model_dict})
    assert torch.equal(model.state_dict()["pretrained.norm.weight"], ckpt["pretrained.norm.weight"])


def test_forward_without_gpu_fails_loudly():
    model = E.endodav(encoder="vits", features=64, out_channels=[48, 96, 192, 384], lora_type="dvlora", disable_conv_head=True)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.EndoDAVError):
        model(torch.rand(1, 2, 3, 28, 28))


@pytest.mark.parametrize("name", ["fwd_vits_dvlora", "fwd_vits_b2", "fwd_vits_ssb_tlora", "fwd_vits_dash_tlora",
                                  "fwd_vits_lora_res_convhead", "fwd_vits_rope", "fwd_vitl"])
def test_packed_graph_matches_reference_golden(name):
    """pack.py + the engine's graph algebra reproduce the reference outputs (fp32, CPU emulation)."""
    m, arrays = load_case(name)
    ctor = m["ctor"]
    cfg = oracle_cfg(ctor)
    sd = weights.make_state_dict(cfg, m["weight_seed"])
    ishape = tuple(ctor["image_shape"])
    pk = pack.pack_state_dict(sd, cfg, torch.float32)
    pk.update(pack.pos_tables(sd, cfg, ishape[0] // 14, ishape[1] // 14))
    B, T, H, W = m["input"]
    x = weights.make_frames(B, T, H, W, m["frame_seed"])
    with torch.no_grad():
        out = emu.forward(pk, cfg, ENC[cfg["encoder"]], x, ishape)
    for s in range(4):
        got = subsample_like_golden(name, s, out[("disp", s)].numpy())
        ref = arrays["disp%d" % s]
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-4, (name, s, float(np.abs(got - ref).max()))


def test_packed_stage_records_match_oracle():
    """Every stage the CUDA engine can snapshot agrees with the oracle's record of the same stage."""
    m, _ = load_case("fwd_vits_dvlora")
    cfg = oracle_cfg(m["ctor"])
    sd = weights.make_state_dict(cfg, m["weight_seed"])
    ishape = tuple(m["ctor"]["image_shape"])
    B, T, H, W = m["input"]
    x = weights.make_frames(B, T, H, W, m["frame_seed"])
    r_or, r_em = {}, {}
    orc.forward(sd, x, cfg, ishape, record=r_or)
    pk = pack.pack_state_dict(sd, cfg, torch.float32)
    pk.update(pack.pos_tables(sd, cfg, ishape[0] // 14, ishape[1] // 14))
    with torch.no_grad():
        emu.forward(pk, cfg, ENC["vits"], x, ishape, record=r_em)
    assert set(r_or) == set(r_em)
    for k in r_or:
        a, b = r_or[k], r_em[k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        assert float((a - b).abs().max()) <= 2e-4 * max(1.0, float(a.abs().max())), k


def test_pack_dtypes_and_alignment():
    cfg = weights.full_cfg()
    sd = weights.make_state_dict(cfg, 3)
    for dt in (torch.bfloat16, torch.float16, torch.float32):
        pk = pack.pack_state_dict(sd, cfg, dt)
        for k, v in pk.items():
            assert v.is_contiguous()
            if v.dim() == 2 and not k.endswith((".pe", ".rope")):
                assert v.dtype == dt, k
                assert v.shape[1] % 32 == 0, (k, v.shape)   # K of every GEMM is a multiple of 32
                assert v.shape[0] % 32 == 0, (k, v.shape)   # N of every GEMM is a multiple of 32
            else:
                assert v.dtype == torch.float32, k


def test_weight_invalidation_hooks():
    """load_state_dict / .to() / invalidate_weights() drop the packed-weight key so the next forward re-packs;
    the cached tensor list follows the module's current parameters (model.py)."""
    import endodav_b200 as E

    model = E.endodav(encoder="vits", features=64, out_channels=[48, 96, 192, 384], lora_type="dvlora", image_shape=(28, 42),
                      disable_conv_head=True)
    v0 = model._versions()
    model._packed_versions = v0
    model.load_state_dict(model.state_dict())
    assert model._packed_versions is None
    model._packed_versions = model._versions()
    model.double()
    assert model._packed_versions is None
    assert model._versions() != v0                      # new storages after the dtype change
    model._packed_versions = model._versions()
    p = next(model.parameters())
    with torch.no_grad():
        p.add_(1.0)
    assert model._versions() != model._packed_versions  # in-place update bumps the version counter
    model._packed_versions = model._versions()
    p.data.add_(1.0)
    assert model._versions() == model._packed_versions  # .data writes are invisible ...
    model.invalidate_weights()
    assert model._packed_versions is None               # ... hence the explicit hook
