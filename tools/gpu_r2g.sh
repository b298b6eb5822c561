#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "linear" > gpurun_out/r2g_pytest_lin.log 2>&1
echo "linear pytest exit=$?"; tail -n 8 gpurun_out/r2g_pytest_lin.log
timeout 300 python tools/gemm_timeline.py qkv fc1 > gpurun_out/r2g_gemm_timeline.txt 2>&1
echo "timeline exit=$?"; cat gpurun_out/r2g_gemm_timeline.txt
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_fullsize.py tests/test_gpu_endodac.py -m gpu -q --tb=short > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit=$?"; tail -n 12 gpurun_out/r2g_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernels-out gpurun_out/r2g_bench_kernels.json > gpurun_out/r2g_bench.log 2> gpurun_out/r2g_bench.err
echo "bench exit=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/r2g_bench_kernels.json"))
print("ms_per_step", d["ms_per_step"])
for r in d["kernels"][:16]:
    print("   %-26s n=%3d %7.1f us %7.1f TF/s %7.1f GB/s" % (r["name"], r["count"] // d["steps"], r["avg_us"], r["tflops"], r["gbs"]))
PY
