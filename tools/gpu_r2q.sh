#!/bin/bash
# r2q: decoder branches on a side stream: parity + A/B timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_forward.py tests/test_gpu_fullsize.py tests/test_gpu_endodac.py -m gpu -q --tb=short -x > gpurun_out/r2q_pytest.log 2>&1
echo "pytest exit=$?"; tail -n 5 gpurun_out/r2q_pytest.log
for b in 0 1; do
  EDV_BRANCH=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2q_bench_b$b.log 2> gpurun_out/r2q_bench_b$b.err
  echo "bench EDV_BRANCH=$b exit=$?"; tail -c 300 gpurun_out/r2q_bench_b$b.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2q_bench_b$b.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','cuda_graphs')}, d['e2e']['value'])
for k,v in d['extra'].items(): print(k, {kk:v.get(kk) for kk in ('ms_per_step','frames_per_s','seconds')})
PY
done
