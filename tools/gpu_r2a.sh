#!/bin/bash
# Round 2, call A: full GPU parity suite, pipe microbenchmarks, fp16 + bf16 bench lines, bf16 per-stage errors.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rm -f gpurun_out/fullsize_parity.json
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit=$?"; tail -n 30 gpurun_out/r2a_pytest.log
timeout 120 tools/microbench > gpurun_out/r2a_microbench.txt 2>&1
echo "microbench exit=$?"; cat gpurun_out/r2a_microbench.txt
timeout 600 python bench.py --steps 10 --warmup 3 --kernels-out gpurun_out/r2a_bench_kernels.json > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err
echo "bench exit=$?"; tail -c 3000 gpurun_out/r2a_bench.log
timeout 600 python bench.py --steps 10 --warmup 3 --dtype bf16 --no-cpu-baseline > gpurun_out/r2a_bench_bf16.log 2> gpurun_out/r2a_bench_bf16.err
echo "bench bf16 exit=$?"; tail -c 1500 gpurun_out/r2a_bench_bf16.log
for c in fwd_vits_dvlora fwd_vits_b2 fwd_vitl; do
  for d in bf16 fp16; do
    echo "== $c $d" >> gpurun_out/r2a_stage_diff.txt
    timeout 300 python tests/stage_diff.py $c $d >> gpurun_out/r2a_stage_diff.txt 2>&1
  done
done
cat gpurun_out/r2a_stage_diff.txt
