"""In-kernel timeline of the tcgen05 GEMM (edv_op_linear_timeline) at the ViT-S encoder shapes.
Usage: python tools/gemm_timeline.py [qkv|fc1|fc2|proj ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from endodav_b200 import engine as eng  # noqa: E402

M = 32 * 1370
SHAPES = {"qkv": (1152, 384, 0), "fc1": (1536, 384, 1), "fc2": (384, 1536, 0), "proj": (384, 384, 0)}
g = torch.Generator().manual_seed(0)


def main():
    for name in (sys.argv[1:] or list(SHAPES)):
        N, K, act = SHAPES[name]
        A = (torch.randn(M, K, generator=g)).half().cuda()
        W = (torch.randn(N, K, generator=g) * 0.05).half().cuda()
        b = torch.zeros(N).cuda()
        for _ in range(3):
            eng.op_linear(A, W, b, act)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.op_linear(A, W, b, act)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print("== %s  M=%d N=%d K=%d: %.1f us, %.1f TFLOP/s" % (name, M, N, K, us, 2.0 * M * N * K / us / 1e6))
        _, tl = eng.op_linear_timeline(A, W, b, act)
        t = tl.cpu().numpy()[0]
        t0 = t[0]
        print("   setup %d" % (t[1] - t0))
        rows = []
        for i in range(14):
            if t[9 + 4 * i] == 0:
                break
            rows.append((i, t[8 + 4 * i] - t0, t[9 + 4 * i] - t0, t[10 + 4 * i] - t0 if t[10 + 4 * i] else -1,
                         t[11 + 4 * i] - t0 if t[11 + 4 * i] else -1, t[100 + i] - t0))
        print("   tile:  acc-free   issued   (mainloop)   epi-start  epi-done  (epilogue)   producer-done")
        for (i, a, bq, c, d, p) in rows:
            print("   %3d  %9d %9d   (%6d)   %9d %9d   (%6d)   %9d" % (i, a, bq, bq - a, c, d, d - c if d > 0 else -1, p))
        if t[170] > 0:
            e0 = t[10 + 4 * 5]
            print("   tile 5 epilogue phases (from accumulator-complete): staging-free wait %d, first sub-tile computed %d, barrier %d, "
                  "all sub-tiles staged %d, fence.proxy %d, barrier %d, stores issued %d" % (
                      t[170] - e0, t[171] - e0, t[172] - e0, t[173] - e0, t[174] - e0, t[175] - e0, t[11 + 4 * 5] - e0))
        kb = [int(t[130 + k] - t0) for k in range(30) if t[130 + k] > 0]
        print("   tile 4 k-block ready times: %s" % kb)
        print("   deltas: %s" % [kb[i] - kb[i - 1] for i in range(1, len(kb))])
        if t[230] > 0:
            wt = [int(t[230 + k] - t[200 + k]) for k in range(30) if t[230 + k] > 0]
        else:
            wt = [int(t[130 + k] - t[200 + k]) for k in range(30) if t[130 + k] > 0]
        print("   time spent waiting for each k-block (MMA warp): %s" % wt)


if __name__ == "__main__":
    main()
