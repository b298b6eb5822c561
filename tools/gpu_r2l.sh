#!/bin/bash
# r2l: flash-attention 2-threads-per-row variants + CUDA-graph capture on a private stream
mkdir -p gpurun_out
rm -f gpurun_out/r2l_fa_variants.txt
for cfg in "2 0" "2 8" "2 4"; do
  set -- $cfg
  echo "=== EDV_FA_SPLIT=$1 EDV_FA_POLY=$2" >> gpurun_out/r2l_fa_variants.txt
  EDV_FA_SPLIT=$1 EDV_FA_POLY=$2 timeout 60 python tools/fa_timeline.py >> gpurun_out/r2l_fa_variants.txt 2>&1 || { echo "variant $cfg failed"; }
done
grep -E "===|per launch|iteration 5|rror" gpurun_out/r2l_fa_variants.txt
if grep -q "rror" gpurun_out/r2l_fa_variants.txt; then export EDV_FA_SPLIT=1; echo "falling back to SPLIT=1 for the rest"; fi
timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py -m gpu -q --tb=short -x -k "attention or graph" > gpurun_out/r2l_pytest.log 2>&1
echo "pytest exit=$?"; tail -n 8 gpurun_out/r2l_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --kernels-out gpurun_out/r2l_bench_kernels.json > gpurun_out/r2l_bench.log 2> gpurun_out/r2l_bench.err
echo "bench exit=$?"; tail -c 600 gpurun_out/r2l_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','cuda_graphs')}, d['e2e']['value'])
for k,v in d['extra'].items(): print(k, v)
PY
