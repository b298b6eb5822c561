"""Print the in-kernel timeline of the tcgen05 flash-attention kernel (edv_op_attention_timeline) for the headline
shape (32 frames x 6 heads x 1370 tokens) and time the launch with CUDA events.  Usage: python tools/fa_timeline.py [dtype]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from endodav_b200 import engine as eng  # noqa: E402


def main():
    dt = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float16
    F, S, H = 32, 1370, 6
    g = torch.Generator().manual_seed(1)
    qkv = (torch.randn(F * S, 3 * H * 64, generator=g) * 0.7).to(dt).cuda()
    for _ in range(3):
        eng.op_attention(qkv, F, S, H)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.op_attention(qkv, F, S, H)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print("flash_attention_tc %s: %.1f us per launch, %.1f TFLOP/s" % (dt, us, 4.0 * F * H * S * S * 64 / us / 1e6))
    _, tl = eng.op_attention_timeline(qkv, F, S, H)
    tl = tl.cpu().numpy()
    for c in range(4):
        t = tl[c]
        t0 = t[0]
        it = [int(t[3 + j] - t0) for j in range(11)]
        print("cta %d: setup %d, first S %d, softmax iteration ends (tile A) %s" % (c, t[1] - t0, t[2] - t0, it))
        print("        per-iteration: %s" % [it[j] - it[j - 1] for j in range(1, 11)])
        print("        PV issue times %s" % [int(t[24 + j] - t0) for j in range(11)])
        print("        O complete %d, stored %d;  iteration 5 phases: ld %d, max %d, exp %d, wait PV(j-1) %d, st P %d" % (
            t[20] - t0, t[21] - t0, t[44] - t[3 + 4], t[45] - t[44], t[46] - t[45], t[47] - t[46], t[48] - t[47]))
        ends = [int(t[50 + i] - t0) for i in range(12) if t[50 + i] > 0]
        print("        item ends (tile A) %s -> per item %s" % (ends, [ends[i] - ends[i - 1] for i in range(1, len(ends))]))


if __name__ == "__main__":
    main()
