"""Times the long-video driver on a SCARED-shaped synthetic video (BASELINE config 3) with the GPU
and the host preprocessing paths.   python tools/time_video.py [frames]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import endodav_b200 as E  # noqa: E402
from endodav_b200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
model = E.endodav(encoder="vits", features=64, out_channels=[48, 96, 192, 384], r=4, lora_type="dvlora", image_shape=(224, 280),
                  disable_conv_head=True, residual_block_indexes=[])
synthetic.randomize_(model, 1234)
model = model.cuda().eval()
rng = np.random.default_rng(0)
video = rng.integers(0, 256, size=(n, 256, 320, 3), dtype=np.uint8)
for mode in ("gpu", "host"):
    os.environ["ENDODAV_PREPROCESS"] = mode
    model.infer_video_depth(video[:64])
    torch.cuda.synchronize()
    for rep in range(2):     # the first full-size call also pins its 655 MB output array (recycled by torch afterwards)
        t0 = time.perf_counter()
        out = model.infer_video_depth(video)
        dt = time.perf_counter() - t0
        if rep == 0:
            print("  first call %.2f s" % dt)
            del out
    print("infer_video_depth %d frames 256x320, preprocessing=%s: %.2f s -> %.0f frames/s (output %s)" % (n, mode, dt, n / dt, out.shape))

if "--wb-sweep" in sys.argv:
    os.environ["ENDODAV_PREPROCESS"] = "gpu"
    for wb in (1, 2, 4, 6, 8):
        os.environ["ENDODAV_WINDOW_BATCH"] = str(wb)
        model.infer_video_depth(video[:400])
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            model.infer_video_depth(video)
            best = min(best, time.perf_counter() - t0)
        print("window batch %d: %.3f s -> %.0f frames/s" % (wb, best, n / best))
    del os.environ["ENDODAV_WINDOW_BATCH"]

if "--profile" in sys.argv:
    import cProfile
    import pstats

    os.environ["ENDODAV_PREPROCESS"] = "gpu"
    pr = cProfile.Profile()
    pr.enable()
    model.infer_video_depth(video)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
