"""Runs the hot kernels at the BASELINE config-2 shapes through the C ABI (for ncu).

  python tools/prof_ops.py [ops...]            # two plain repetitions of each op
  ncu --profile-from-start off ... python tools/prof_ops.py --range [ops...]
                                               # warm every op once, profile ONE repetition of each
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from endodav_b200 import engine as eng  # noqa: E402

dt = torch.float16
M = 32 * 1370
g = torch.Generator().manual_seed(0)
ALL = ["fc1", "qkv", "fc2", "attn", "rcu", "tattn", "head", "oc", "ln", "proj_ln", "fc2_ln", "ups"]


def rnd(*s, scale=1.0):
    return (torch.randn(*s, generator=g) * scale).to(dt).cuda()


def run(w):
    if w == "fc1":
        eng.op_linear(rnd(M, 384), rnd(1536, 384, scale=0.05), torch.zeros(1536).cuda(), 1)
    elif w == "qkv":
        eng.op_linear(rnd(M, 384), rnd(1152, 384, scale=0.05), torch.zeros(1152).cuda(), 0)
    elif w == "fc2":
        eng.op_linear(rnd(M, 1536), rnd(384, 1536, scale=0.03), torch.zeros(384).cuda(), 0)
    elif w == "attn":
        eng.op_attention(rnd(M, 1152, scale=0.5), 32, 1370, 6)
    elif w == "rcu":
        eng.op_conv3x3(rnd(32, 148, 148, 64), rnd(64, 576, scale=0.04), torch.zeros(64).cuda(), True)
    elif w == "head":
        eng.op_disp_head(rnd(32, 296, 296, 32), rnd(32, 288, scale=0.06), torch.zeros(32).cuda(), torch.ones(33).cuda() * 0.1, 518, 518)
    elif w == "oc":
        eng.op_conv3x3(rnd(32, 296, 296, 64), rnd(32, 576, scale=0.04), torch.zeros(32).cuda(), False)
    elif w == "tattn":
        eng.op_temporal_attention(rnd(32 * 5476, 192, scale=0.7), 1, 32, 5476, 64)
    elif w == "proj_ln":   # blk.proj: x += A W^T + b (fp32 residual stream), LayerNorm(x) -> 16 bit, one kernel (gemm_ln.cuh)
        eng.op_linear_residual_ln(rnd(M, 384), rnd(384, 384, scale=0.05), torch.zeros(384).cuda(), torch.randn(M, 384, generator=g).cuda(),
                                  torch.ones(384).cuda(), torch.zeros(384).cuda())
    elif w == "fc2_ln":
        eng.op_linear_residual_ln(rnd(M, 1536), rnd(384, 1536, scale=0.03), torch.zeros(384).cuda(), torch.randn(M, 384, generator=g).cuda(),
                                  torch.ones(384).cuda(), torch.zeros(384).cuda())
    elif w == "ups":       # fusion-block upsample 148^2 -> 296^2, 64 channels
        eng.op_upsample(rnd(32, 148, 148, 64), 296, 296)
    elif w == "ln":
        eng.op_layernorm(torch.randn(M, 384, generator=g).cuda(), torch.ones(384).cuda(), torch.zeros(384).cuda(), 1e-6, dt)
    else:
        raise SystemExit("unknown op %r (known: %s)" % (w, ALL))


which = [a for a in sys.argv[1:] if not a.startswith("--")] or ALL
use_range = "--range" in sys.argv
for w in which:
    run(w)
    if not use_range:
        run(w)
torch.cuda.synchronize()
if use_range:
    torch.cuda.cudart().cudaProfilerStart()
    for w in which:
        run(w)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok: " + " ".join(which))
