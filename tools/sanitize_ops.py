"""One small launch of every hand-written kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):

  compute-sanitizer --tool memcheck python tools/sanitize_ops.py

Sizes are small (sanitizer slow-down is 10-100x) but cover every warp-specialised tcgen05 kernel with partial tiles."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import endodav_b200 as E  # noqa: E402
from endodav_b200 import engine as eng  # noqa: E402
from endodav_b200 import synthetic  # noqa: E402

g = torch.Generator().manual_seed(0)
dt = torch.float16


def rnd(*s, scale=1.0, dtype=dt):
    return (torch.randn(*s, generator=g) * scale).to(dtype).cuda()


# streaming GEMM (K = 1536 -> gemm_tc), B-resident GEMM (K = 384), GELU epilogue, ragged M
eng.op_linear(rnd(300, 1536), rnd(384, 1536, scale=0.03), torch.zeros(384).cuda(), 0)
eng.op_linear(rnd(2100, 384), rnd(1152, 384, scale=0.05), torch.zeros(1152).cuda(), 0)
eng.op_linear(rnd(2100, 384), rnd(1536, 384, scale=0.05), torch.zeros(1536).cuda(), 1)
eng.op_linear(rnd(130, 96), rnd(32, 96, scale=0.1), None, 2)
# fused GEMM + residual + LayerNorm (cluster pair)
x = rnd(700, 384, dtype=torch.float32)
eng.op_linear_residual_ln(rnd(700, 384), rnd(384, 384, scale=0.05), torch.zeros(384).cuda(), x, torch.ones(384).cuda(), torch.zeros(384).cuda())
# flash attention (persistent, two items per CTA at least on a small grid is not guaranteed: 3 frames x 6 heads x 2 tiles)
eng.op_attention(rnd(3 * 300, 1152, scale=0.5), 3, 300, 6)
# convs: halo kernel (64 -> 64, 64 -> 32) and TMA implicit GEMM (192 -> 64)
eng.op_conv3x3(rnd(2, 20, 27, 64), rnd(64, 576, scale=0.04), torch.zeros(64).cuda(), True)
eng.op_conv3x3(rnd(1, 33, 47, 64), rnd(32, 576, scale=0.04), torch.zeros(32).cuda(), False)
eng.op_conv3x3(rnd(1, 19, 19, 192), rnd(64, 1728, scale=0.03), None, False)
# fused disparity head
eng.op_disp_head(rnd(2, 40, 52, 32), rnd(32, 288, scale=0.06), torch.zeros(32).cuda(), torch.ones(33).cuda() * 0.1, 70, 91)
# temporal attention, norms, resampling
eng.op_temporal_attention(rnd(8 * 50, 192, scale=0.7), 1, 8, 50, 64)
eng.op_layernorm(torch.randn(301, 384, generator=g).cuda(), torch.ones(384).cuda(), torch.zeros(384).cuda(), 1e-6, dt)
eng.op_groupnorm(rnd(3, 77, 64), torch.ones(64).cuda(), torch.zeros(64).cuda(), 1e-6)
eng.op_upsample(rnd(2, 19, 19, 64), 37, 37)
eng.op_resize_f32(torch.rand(2, 30, 40, generator=g).cuda(), 64, 80)
eng.op_cubic_resize_u8(torch.randint(0, 256, (2, 48, 64, 3), generator=g, dtype=torch.uint8).cuda(), 28, 42)
torch.cuda.synchronize()
# a whole forward (ViT-S, 3 frames at 70 x 98) and a two-window video with on-GPU stitching
ctor = dict(encoder="vits", features=64, out_channels=[48, 96, 192, 384], r=4, lora_type="dvlora", image_shape=(70, 98),
            disable_conv_head=True, residual_block_indexes=[])
model = E.endodav(dtype="fp16", **ctor)
synthetic.randomize_(model, 1)
model = model.cuda().eval()
out = model(torch.rand(1, 3, 3, 70, 98, generator=g).cuda())
assert bool(torch.isfinite(out[("disp", 0)]).all())
model2 = E.endodav(dtype="fp16", **dict(ctor, image_shape=(28, 42)))
synthetic.randomize_(model2, 1)
v = np.random.default_rng(0).integers(0, 256, size=(30, 40, 56, 3), dtype=np.uint8)
d = model2.cuda().eval().infer_video_depth(v)
assert d.shape == (30, 40, 56) and np.isfinite(d).all()
torch.cuda.synchronize()
print("sanitize_ops: ok")
