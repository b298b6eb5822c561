"""BASELINE config 5: Hamlyn-shaped clip-batch sweep -- B in 1..64 clips x T in {8,16,32} frames at 256x320
(network resolution 224x280).  Prints frames/s of endodav.forward and the share of the temporal-attention
kernel, and writes profiles-style JSON.   python tools/sweep_clips.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import endodav_b200 as E  # noqa: E402
from endodav_b200 import synthetic  # noqa: E402

model = E.endodav(encoder="vits", features=64, out_channels=[48, 96, 192, 384], r=4, lora_type="dvlora", image_shape=(224, 280),
                  disable_conv_head=True, residual_block_indexes=[])
synthetic.randomize_(model, 1234)
model = model.cuda().eval()
rows = []
for T in (8, 16, 32):
    for B in (1, 2, 4, 8, 16, 32, 64):
        x = torch.rand(B, T, 3, 256, 320, device="cuda")
        for _ in range(2):
            model(x)
        eng = model._eng
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if B * T <= 512 else 2
        e0.record()
        for _ in range(reps):
            model(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        eng.profile(True)
        model(x)
        tab = eng.profile_collect()
        eng.profile(False)
        tot = sum(r["ms"] for r in tab)
        ta = sum(r["ms"] for r in tab if r["name"].startswith("temporal_attention"))
        tb = sum(r["bytes"] for r in tab if r["name"].startswith("temporal_attention"))
        rows.append(dict(B=B, T=T, frames=B * T, ms=ms, frames_per_s=B * T / ms * 1e3, temporal_attention_ms=ta,
                         temporal_attention_share=ta / tot, temporal_attention_gbs=tb / (ta * 1e-3) / 1e9 if ta else 0.0))
        print("B=%2d T=%2d: %8.2f ms  %8.0f frames/s   temporal attention %.3f ms (%.1f%% of kernels, %.0f GB/s)" % (
            B, T, ms, rows[-1]["frames_per_s"], ta, 100 * ta / tot, rows[-1]["temporal_attention_gbs"]))
        del x
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump(dict(workload="ViT-S, 256x320 frames, network 224x280, fp16", rows=rows), f, indent=1)
