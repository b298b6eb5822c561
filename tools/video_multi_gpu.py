"""BASELINE config 3 on N GPUs: SCARED-shaped synthetic video through ``infer_video_depth`` with the windows
sharded over the ranks, one NCCL gather to rank 0 and the on-GPU stitching there.  Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/video_multi_gpu.py [frames]
Prints one JSON line on rank 0 (time = max over ranks, barrier on both sides) and checks the result against
the single-GPU run of the same video on rank 0 (must be bit-identical: same kernels, same order)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import endodav_b200 as E  # noqa: E402
from endodav_b200 import synthetic, video  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
torch.manual_seed(0)   # the base init draws from the global RNG: every rank must build the same weights
model = E.endodav(encoder="vits", features=64, out_channels=[48, 96, 192, 384], r=4, lora_type="dvlora", image_shape=(224, 280),
                  disable_conv_head=True, residual_block_indexes=[])
synthetic.randomize_(model, 1234)
model = model.cuda().eval()
frames = np.random.default_rng(0).integers(0, 256, size=(n, 256, 320, 3), dtype=np.uint8)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


model.infer_video_depth(frames[:200])      # warm-up: plan, NCCL communicator, pinned buffers
times = []
for _ in range(3):
    barrier()
    t0 = time.perf_counter()
    out = model.infer_video_depth(frames)
    barrier()
    times.append(time.perf_counter() - t0)
t = torch.tensor([min(times)], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    single = video.infer_video_depth(model, frames, distributed=False)
    same = bool(np.array_equal(single, out))
    if "--debug" in sys.argv:
        os.environ["ENDODAV_STITCH"] = "host"
        host = video.infer_video_depth(model, frames, distributed=False)
        os.environ["ENDODAV_STITCH"] = "gpu"
        for name, a, b in (("multi vs single", out, single), ("multi vs host", out, host), ("single vs host", single, host)):
            d = np.abs(a - b).reshape(n, -1).max(1)
            bad = np.nonzero(d > 1e-4)[0]
            print(name, "max", float(d.max()), "first bad frames", bad[:12].tolist(), "count", int(bad.size), flush=True)
    print(json.dumps({"workload": "scared_2000x256x320", "n_gpus": world, "frames": n, "seconds": float(t.item()),
                      "frames_per_s": n / float(t.item()), "equals_single_gpu": same,
                      "max_abs_diff": float(np.abs(single - out).max())}))
else:
    assert out is None
if world > 1:
    dist.destroy_process_group()
