"""Diagnostic (not collected by pytest): per-stage error of the CUDA path against the oracle.

  python tests/stage_diff.py <golden case> <dtype> [engine]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import endodav_b200 as E  # noqa: E402
from oracle import endodav_oracle as orc  # noqa: E402
from oracle import weights  # noqa: E402
from golden_util import load_case, oracle_cfg  # noqa: E402


def main():
    name, dtype = sys.argv[1], sys.argv[2]
    if len(sys.argv) > 3:
        os.environ["ENDODAV_ENGINE"] = sys.argv[3]
    m, arrays = load_case(name)
    ctor = dict(m["ctor"])
    ctor["image_shape"] = tuple(ctor["image_shape"])
    cfg = oracle_cfg(m["ctor"])
    sd = weights.make_state_dict(cfg, m["weight_seed"])
    B, T, H, W = m["input"]
    x = weights.make_frames(B, T, H, W, m["frame_seed"])
    rec, rec16 = {}, {}
    ref = orc.forward(sd, x, cfg, ctor["image_shape"], record=rec)
    emu = orc.forward(sd, x, cfg, ctor["image_shape"], record=rec16, emulate_bf16="f16" if dtype == "fp16" else True)
    model = E.endodav(dtype=dtype, **ctor)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    h, w = ctor["image_shape"]
    eng = model._ensure_engine(h // 14, w // 14)
    eng.set_debug(True)
    out = model(x.cuda())
    print("%-10s %12s %12s %12s %12s" % ("stage", "ref absmax", "cuda err", "emu16 err", "nan?"))
    for k, r in rec.items():
        got = eng.debug_tap(k).cpu()
        r2 = r.permute(0, 2, 3, 1) if r.dim() == 4 else r
        r2 = r2.reshape(-1, r2.shape[-1])
        e2 = rec16[k].permute(0, 2, 3, 1) if r.dim() == 4 else rec16[k]
        e2 = e2.reshape(-1, e2.shape[-1])
        print("%-10s %12.4g %12.4g %12.4g %12s" % (k, float(r2.abs().max()), float((got - r2).abs().max()),
                                                  float((e2 - r2).abs().max()), bool(torch.isnan(got).any())))
    for s in range(4):
        g = out[("disp", s)].cpu()
        r = ref[("disp", s)]
        e = emu[("disp", s)]
        rel = ((g - r).abs() / r.abs().clamp_min(1e-3)).max()
        rele = ((e - r).abs() / r.abs().clamp_min(1e-3)).max()
        print("disp%d ref[%.4g,%.4g] cuda rel %.4g  emu16 rel %.4g" % (s, float(r.min()), float(r.max()), float(rel), float(rele)))


if __name__ == "__main__":
    main()
