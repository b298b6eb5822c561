"""Diagnostic: per-stage difference between the fp16 forward on weights W and on the function-preserving
power-of-two re-parameterisation of tests/test_gpu_fullsize.py (which stage stops being scale-invariant?)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import endodav_b200 as E  # noqa: E402
from oracle import weights  # noqa: E402
from golden_util import oracle_cfg  # noqa: E402

src = open(os.path.join(ROOT, "tests", "test_gpu_fullsize.py")).read()
ns = {}
exec(src[src.index("def _scaled_encoder_sd"):src.index('@pytest.mark.parametrize("scale"')], ns)

ctor = dict(encoder="vits", features=64, out_channels=[48, 96, 192, 384], r=4, lora_type="dvlora",
            image_shape=(70, 98), disable_conv_head=True, residual_block_indexes=[])
cfg = oracle_cfg(ctor)
sd = weights.make_state_dict(cfg, 1234)
x = weights.make_frames(1, 4, 70, 98, 4321).cuda()


def run(state):
    m = E.endodav(dtype="fp16", **ctor)
    m.load_state_dict(state, strict=True)
    m = m.cuda().eval()
    eng = m._ensure_engine(5, 7)
    eng.set_debug(True)
    d = m(x)[("disp", 0)].float().cpu()
    taps = {k: eng.debug_tap(k).cpu() for k in ("tokens0", "block0", "tap0", "tap1", "tap2", "tap3", "layer1", "layer2", "layer3", "layer4", "mm0", "path1")}
    return d, taps


s = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
d0, t0 = run(sd)
d1, t1 = run(ns["_scaled_encoder_sd"](sd, s))
for k in t0:
    a, b = t0[k], t1[k]
    if k.startswith("tap"):
        b = b / s
    print("%-8s max|a| %.4g  max diff %.4g" % (k, float(a.abs().max()), float((a - b).abs().max())))
print("disp0 max diff %.4g (mean %.4g)" % (float((d0 - d1).abs().max()), float(d0.mean())))
