#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_forward.py tests/test_gpu_stitch.py tests/test_gpu_endodac.py -m gpu -q --tb=short -x -k "video or stitch or preprocessing or long" > gpurun_out/r3a_pytest.log 2>&1
echo "video tests exit=$?"; tail -n 4 gpurun_out/r3a_pytest.log
ENDODAV_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/video_multi_gpu.py 2000 > gpurun_out/r3a_video_n8.log 2> gpurun_out/r3a_video_n8.err
echo exit=$?; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r3a_video_n8.log | tail -8; tail -c 300 gpurun_out/r3a_video_n8.err
