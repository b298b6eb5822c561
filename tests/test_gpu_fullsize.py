"""Parity of the CUDA path at the BASELINE configurations' REAL sizes.

  * BASELINE config 2 -- ViT-S, one 32-frame 518 x 518 clip (S = 1370 tokens, 37x37 -> 19x19 maps, T = 32): the CUDA
    path in fp32 / fp16 / bf16 against the CPU oracle on the FULL maps, and against the strided sample of the
    UNMODIFIED reference's output committed under tests/golden (full_vits_518_t32, oracle/make_golden.py);
  * BASELINE config 4's model -- ViT-L at 518 x 518, two frames;
  * fp16 range safety: every 16-bit intermediate of the headline configuration is scanned for inf / near-overflow,
    and the encoder is re-run with its 16-bit operands scaled by 8x and 64x (function-preserving power-of-two
    re-parameterisation), which must reproduce the unscaled result unless something saturates.

Same gates as tests/test_gpu_forward.py.  Beside the floored relative error the tests record the UNFLOORED
distribution (|d - ref| / |ref| percentiles and how many pixels the floor touches) in
gpurun_out/fullsize_parity.json so the numbers can be quoted in DESIGN.md.  /root/reference is never read here."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import endodav_b200 as E  # noqa: E402
from oracle import endodav_oracle as orc  # noqa: E402
from oracle import weights  # noqa: E402
from golden_util import load_case, oracle_cfg  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "fullsize_parity.json")
FP32_RTOL = 2e-4
GATES = {"fp16": dict(rel=1e-2, absrel=1e-3, a1=0.999), "bf16": dict(rel=6e-2, absrel=1e-2, a1=0.999)}


def _report(key, value):
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        data = {}
        if os.path.exists(REPORT):
            with open(REPORT) as f:
                data = json.load(f)
        data[key] = value
        with open(REPORT, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass


def _metrics(pred, gt):
    """compute_errors (utils/utils.py:112-133) with the reference output as gt."""
    pred, gt = pred.astype(np.float64).ravel(), gt.astype(np.float64).ravel()
    m = gt > 0.25 * gt.mean()
    pred, gt = np.maximum(pred[m], 1e-9), gt[m]
    thresh = np.maximum(gt / pred, pred / gt)
    return float(np.mean(np.abs(gt - pred) / gt)), float((thresh < 1.25).mean())


def error_summary(got, ref):
    """Floored gate metric + the unfloored per-pixel relative error distribution."""
    got, ref = got.astype(np.float64), ref.astype(np.float64)
    mean = float(np.abs(ref).mean())
    diff = np.abs(got - ref)
    floored = diff / np.maximum(np.abs(ref), 0.5 * mean)
    nz = np.abs(ref) > 0
    raw = diff[nz] / np.abs(ref[nz])
    absrel, a1 = _metrics(got, ref)
    return dict(
        floored_max=float(floored.max()), absrel=absrel, a1=a1, mean_disp=mean,
        abs_err_max=float(diff.max()),
        unfloored_p50=float(np.percentile(raw, 50)), unfloored_p99=float(np.percentile(raw, 99)),
        unfloored_p999=float(np.percentile(raw, 99.9)), unfloored_max=float(raw.max()),
        pixels=int(ref.size), pixels_ref_zero=int((~nz).sum()),
        pixels_below_floor=int((np.abs(ref) < 0.5 * mean).sum()),
        abs_err_max_where_ref_zero=float(diff[~nz].max()) if (~nz).any() else 0.0)


def _ctor(m):
    kw = dict(m["ctor"])
    kw["image_shape"] = tuple(kw["image_shape"])
    return kw


class _Case:
    def __init__(self, name):
        self.m, self.gold = load_case(name)
        self.ctor = _ctor(self.m)
        self.cfg = oracle_cfg(self.m["ctor"])
        self.sd = weights.make_state_dict(self.cfg, self.m["weight_seed"])
        B, T, H, W = self.m["input"]
        self.x = weights.make_frames(B, T, H, W, self.m["frame_seed"])
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            out = orc.forward(self.sd, self.x, self.cfg, self.ctor["image_shape"])
        self.ref = {s: out[("disp", s)].numpy() for s in range(4)}
        # the oracle itself against the committed sample of the unmodified reference
        st = self.m["stride"]
        samp = self.ref[0][self.m["frames"]][:, :, ::st, ::st]
        assert np.abs(samp - self.gold["disp0"]).max() <= 5e-5

    def run(self, dtype):
        model = E.endodav(dtype=dtype, **self.ctor)
        model.load_state_dict(self.sd, strict=True)
        model = model.cuda().eval()
        out = model(self.x.cuda())
        assert model._eng.launch_count() > 0
        return model, {s: out[("disp", s)].cpu().numpy() for s in range(4)}


@pytest.fixture(scope="module")
def vits_full():
    return _Case("full_vits_518_t32")


@pytest.fixture(scope="module")
def vitl_full():
    return _Case("full_vitl_518_t2")


def _check_fp32(case, tag):
    _, got = case.run("fp32")
    worst = 0.0
    for s in range(4):
        assert got[s].shape == case.ref[s].shape
        err = np.abs(got[s] - case.ref[s]) / np.maximum(np.abs(case.ref[s]), 1.0)
        worst = max(worst, float(err.max()))
    # directly against the reference's own numbers (strided sample)
    st = case.m["stride"]
    samp = got[0][case.m["frames"]][:, :, ::st, ::st]
    gerr = float((np.abs(samp - case.gold["disp0"]) / np.maximum(np.abs(case.gold["disp0"]), 1.0)).max())
    _report(tag + ".fp32", dict(max_err_vs_oracle=worst, max_err_vs_reference_sample=gerr))
    assert worst <= FP32_RTOL, worst
    assert gerr <= FP32_RTOL, gerr


def _check_16(case, dtype, tag):
    _, got = case.run(dtype)
    gate = GATES[dtype]
    summ = error_summary(got[0], case.ref[0])
    st = case.m["stride"]
    samp = got[0][case.m["frames"]][:, :, ::st, ::st]
    g = case.gold["disp0"]
    summ["floored_max_vs_reference_sample"] = float(
        (np.abs(samp - g) / np.maximum(np.abs(g), 0.5 * float(np.abs(g).mean()))).max())
    _report(tag + "." + dtype, summ)
    assert summ["floored_max"] <= gate["rel"], summ
    assert summ["floored_max_vs_reference_sample"] <= gate["rel"], summ
    assert summ["absrel"] <= gate["absrel"] and summ["a1"] >= gate["a1"], summ
    for s in range(1, 4):
        e = np.abs(got[s] - case.ref[s]) / np.maximum(np.abs(case.ref[s]), 0.5 * float(np.abs(case.ref[s]).mean()))
        assert e.max() <= gate["rel"], (dtype, s, float(e.max()))


def test_vits_518_t32_fp32(vits_full):
    _check_fp32(vits_full, "vits_518_t32")


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_vits_518_t32_16bit(vits_full, dtype):
    _check_16(vits_full, dtype, "vits_518_t32")


def test_vitl_518_t2_fp32(vitl_full):
    _check_fp32(vitl_full, "vitl_518_t2")


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_vitl_518_t2_16bit(vitl_full, dtype):
    _check_16(vitl_full, dtype, "vitl_518_t2")


# ---- fp16 range safety ----------------------------------------------------------------------------
def _scan(model):
    """max |x| and count of non-finite / near-overflow (> 3e4) entries of every plan buffer after a forward."""
    out = {}
    for name, t in model._eng.plan_buffers().items():
        if name in ("mm.stats",):
            continue
        tf = t.float()
        fin = torch.isfinite(tf)
        out[name] = dict(max_abs=float(tf[fin].abs().max()) if bool(fin.any()) else 0.0,
                         nonfinite=int((~fin).sum()), over_3e4=int((tf[fin].abs() > 3e4).sum()), fp32=t.dtype == torch.float32)
    return out


def test_fp16_intermediates_have_headroom(vits_full):
    """Every 16-bit buffer of the headline configuration: no inf / NaN, nothing above 3e4 (fp16 max 65504)."""
    model, got = vits_full.run("fp16")
    scan = _scan(model)
    _report("vits_518_t32.fp16.range", scan)
    assert np.isfinite(got[0]).all()
    for name, st in scan.items():
        assert st["nonfinite"] == 0, (name, st)
        if not st["fp32"]:
            assert st["over_3e4"] == 0, (name, st)


def _scaled_packed(packed, s, depth, D):
    """Power-of-two re-parameterisation of the PACKED weights (endodav_b200/pack.py names) that leaves the network
    function unchanged but makes 16-bit intermediates s times larger: LayerNorm outputs (xn, the four taps) via
    gamma / beta * s with the consuming weight matrices / s; V and the attention output via the V bias * s (the V
    rows keep their weights) with proj.w / s.  Powers of two commute with fp16 rounding, so the fp16 result must
    be the unscaled one unless an activation overflows -- up to fp16 SUBNORMAL effects: a few LayerNorm outputs below
    6e-5 carry fewer mantissa bits in the unscaled run than in the scaled one, and those last-bit differences are
    amplified like any other rounding noise (tools/scale_ops_diag.py: every kernel is scale-exact except for
    subnormal outputs).  Weights whose scaled value would be an fp16 subnormal are flushed to zero in both versions.
    Returns (base, scaled)."""
    tiny = s * 2.0 ** -14
    base = {k: v.clone() for k, v in packed.items()}
    down = ["proj%d.w" % i for i in range(4)]
    for i in range(depth):
        down += ["blk%d.proj.w" % i, "blk%d.fc1.w" % i]
    for k in down:
        base[k][base[k].abs().float() < tiny] = 0
    for i in range(depth):
        w = base["blk%d.qkv.w" % i]
        qk = w[:2 * D]
        qk[qk.abs().float() < tiny] = 0
    scaled = {k: v.clone() for k, v in base.items()}
    for k in down:
        scaled[k] = (scaled[k].float() / s).to(scaled[k].dtype)
    for i in range(depth):
        b = "blk%d." % i
        for n in ("ln1", "ln2"):
            scaled[b + n + ".w"] = scaled[b + n + ".w"] * s
            scaled[b + n + ".b"] = scaled[b + n + ".b"] * s
        w = scaled[b + "qkv.w"].float()
        w[:2 * D] /= s                                   # q, k rows see the s-times larger LN output; V rows keep their weights
        scaled[b + "qkv.w"] = w.to(scaled[b + "qkv.w"].dtype)
        scaled[b + "qkv.b"][2 * D:] *= s                 # v (hence P v) comes out s times larger
    scaled["norm.w"] = scaled["norm.w"] * s
    scaled["norm.b"] = scaled["norm.b"] * s
    return base, scaled


@pytest.mark.parametrize("scale", [8.0, 64.0])
def test_fp16_scaled_operands_do_not_saturate(scale):
    """LN outputs, the four taps, V and the attention output scaled by 8x / 64x (DINOv2 checkpoints carry
    large-magnitude channels): the fp16 path must stay finite, nothing may come near the fp16 limit, and the result
    must agree with the unscaled fp16 run to within fp16 rounding noise (the same gates as against the oracle)."""
    from endodav_b200 import pack

    ctor = dict(encoder="vits", features=64, out_channels=[48, 96, 192, 384], r=4, lora_type="dvlora",
                image_shape=(518, 518), disable_conv_head=True, residual_block_indexes=[])
    cfg = oracle_cfg(ctor)
    sd = weights.make_state_dict(cfg, 1234)
    x = weights.make_frames(1, 4, 518, 518, 4321).cuda()
    model = E.endodav(dtype="fp16", **ctor)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    model(x)                                             # creates the engine and packs the weights
    eng = model._eng
    base_w, scaled_w = _scaled_packed(pack.pack_state_dict(sd, model._cfg, torch.float16), scale, 12, 384)
    eng.set_weights(base_w)
    base = model(x)[("disp", 0)].cpu().numpy()
    eng.set_weights(scaled_w)
    got = model(x)[("disp", 0)].cpu().numpy()
    scan = _scan(model)
    rel = np.abs(got - base) / np.maximum(np.abs(base), 0.5 * float(np.abs(base).mean()))
    _report("vits_518_t4.fp16.scaled_x%d" % int(scale), dict(
        max_abs_16bit={k: v["max_abs"] for k, v in scan.items() if not v["fp32"]},
        over_3e4={k: v["over_3e4"] for k, v in scan.items() if v["over_3e4"]},
        max_rel_vs_unscaled=float(rel.max()), bit_identical=bool(np.array_equal(got, base))))
    assert np.isfinite(got).all()
    for name, st in scan.items():
        assert st["nonfinite"] == 0, (name, st)
        if not st["fp32"]:
            assert st["over_3e4"] == 0, (name, st)
    assert float(np.abs(base).mean()) > 0.05
    absrel, a1 = _metrics(got, base)
    assert rel.max() <= GATES["fp16"]["rel"] and absrel <= GATES["fp16"]["absrel"] and a1 >= GATES["fp16"]["a1"], (float(rel.max()), absrel, a1)
