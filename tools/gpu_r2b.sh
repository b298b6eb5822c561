#!/bin/bash
# Round 2, call B: parity suite after the warp-uniform MMA issue + flash attention v2 + CUDA graphs, bench, FA poly sweep.
mkdir -p gpurun_out
rm -f gpurun_out/fullsize_parity.json
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r2b_pytest.log 2>&1
echo "pytest exit=$?"; tail -n 25 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernels-out gpurun_out/r2b_bench_kernels.json > gpurun_out/r2b_bench.log 2> gpurun_out/r2b_bench.err
echo "bench exit=$?"; tail -c 2500 gpurun_out/r2b_bench.log
for pm in 4 3 2; do
  EDV_FA_POLY=$pm timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_pm$pm.log 2>&1
  echo "pm=$pm exit=$?"; python - <<PY
import json
for line in open("gpurun_out/r2b_bench_pm$pm.log"):
    if line.startswith("{"):
        d = json.loads(line); print("pm$pm", d["value"], d["ms_per_step"], [k for k in d["top_kernels"] if "flash" in k["name"]])
PY
done
EDV_GRAPH=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_nograph.log 2>&1
echo "nograph exit=$?"; tail -c 600 gpurun_out/r2b_bench_nograph.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --workload vits_224x280_t8 > gpurun_out/r2b_bench_cfg1.log 2>&1
echo "cfg1 exit=$?"; tail -c 1500 gpurun_out/r2b_bench_cfg1.log
EDV_GRAPH=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --workload vits_224x280_t8 > gpurun_out/r2b_bench_cfg1_nograph.log 2>&1
echo "cfg1 nograph exit=$?"; tail -c 600 gpurun_out/r2b_bench_cfg1_nograph.log
