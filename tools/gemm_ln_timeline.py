"""In-kernel epilogue timeline of the fused GEMM + residual + LayerNorm kernel (gemm_ln.cuh) at the ViT-S shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from endodav_b200 import engine as eng  # noqa: E402

M, N = 32 * 1370, 384
g = torch.Generator().manual_seed(0)
for name, K in (("proj", 384), ("fc2", 1536)):
    A = torch.randn(M, K, generator=g).half().cuda()
    W = (torch.randn(N, K, generator=g) * K ** -0.5).half().cuda()
    b = torch.zeros(N).cuda()
    gam, bet = torch.ones(N).cuda(), torch.zeros(N).cuda()
    x = torch.randn(M, N, generator=g).cuda()
    for _ in range(3):
        eng.op_linear_residual_ln(A, W, b, x, gam, bet)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.op_linear_residual_ln(A, W, b, x, gam, bet)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    byts = (M * K + N * K) * 2 + M * N * 10
    print("== %s K=%d: %.1f us, %.1f TFLOP/s, %.0f GB/s algorithmic" % (name, K, us, 2.0 * M * N * K / us / 1e6, byts / us / 1e3))
    tl = torch.zeros(64, dtype=torch.int64, device="cuda")
    eng.op_linear_residual_ln(A, W, b, x, gam, bet, timeline=tl)
    t = tl.cpu().numpy()
    base = t[0]
    print("   tile: res-loads-issued  acc-complete  pass1  mean  pass2  peer-stats  pass3   (cycles from the first stamp; deltas)")
    for i in range(6):
        if t[8 * i + 6] == 0:
            break
        row = [int(t[8 * i + k] - base) for k in range(7)]
        print("   %d  %s   deltas %s" % (i, row, [row[k] - row[k - 1] for k in range(1, 7)]))
