#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gemm_ln_timeline.py > gpurun_out/r2t_gemm_ln_timeline.txt 2>&1; cat gpurun_out/r2t_gemm_ln_timeline.txt
timeout 300 python bench.py --impl reference --reference-device cuda --steps 3 --warmup 1 > gpurun_out/r2t_eager_gpu.log 2> gpurun_out/r2t_eager_gpu.err
echo "eager gpu exit=$?"; tail -c 400 gpurun_out/r2t_eager_gpu.err; cat gpurun_out/r2t_eager_gpu.log
