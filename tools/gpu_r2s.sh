#!/bin/bash
# r2s: gemm_ln with multicast A (parity first, short timeouts), A/B timing, config 4 / config 5 / eager-GPU sanity lines
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "residual_layernorm or linear" > gpurun_out/r2s_pytest_ln.log 2>&1
rc=$?; echo "gemm_ln op tests exit=$rc"; tail -n 4 gpurun_out/r2s_pytest_ln.log
if [ $rc -ne 0 ]; then echo "STOP: gemm_ln broken"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_forward.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -x > gpurun_out/r2s_pytest.log 2>&1
echo "forward tests exit=$?"; tail -n 4 gpurun_out/r2s_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernels-out gpurun_out/r2s_bench_kernels.json > gpurun_out/r2s_bench.log 2> gpurun_out/r2s_bench.err
echo "bench exit=$?"; tail -c 300 gpurun_out/r2s_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s_bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','cuda_graphs')}, d['e2e']['value'])
for k in d['top_kernels']: print(k)
PY
timeout 300 python bench.py --workload vitl_518_t32_b4 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2s_bench_config4.log 2> gpurun_out/r2s_bench_config4.err
echo "config4 exit=$?"; cut -c1-300 gpurun_out/r2s_bench_config4.log
timeout 300 python tools/sweep_clips.py gpurun_out/r2s_config5_sweep.json > gpurun_out/r2s_config5.log 2>&1
echo "config5 exit=$?"; tail -n 5 gpurun_out/r2s_config5.log
timeout 300 python bench.py --impl reference --reference-device cuda --steps 3 --warmup 1 > gpurun_out/r2s_eager_gpu.log 2> gpurun_out/r2s_eager_gpu.err
echo "eager gpu exit=$?"; tail -c 400 gpurun_out/r2s_eager_gpu.err; cat gpurun_out/r2s_eager_gpu.log
