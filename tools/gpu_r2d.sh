#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -k "attention" > gpurun_out/r2d_pytest_attn.log 2>&1
echo "attn pytest exit=$?"; tail -n 8 gpurun_out/r2d_pytest_attn.log
timeout 300 python tools/fa_timeline.py > gpurun_out/r2d_fa_timeline.txt 2>&1
echo "timeline exit=$?"; cat gpurun_out/r2d_fa_timeline.txt
for pm in 4 3; do EDV_FA_POLY=$pm timeout 300 python tools/fa_timeline.py 2>&1 | head -1; done
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_fullsize.py tests/test_metrics.py -m gpu -q --tb=short > gpurun_out/r2d_pytest.log 2>&1
echo "pytest exit=$?"; tail -n 12 gpurun_out/r2d_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernels-out gpurun_out/r2d_bench_kernels.json > gpurun_out/r2d_bench.log 2> gpurun_out/r2d_bench.err
echo "bench exit=$?"; tail -c 1800 gpurun_out/r2d_bench.log
