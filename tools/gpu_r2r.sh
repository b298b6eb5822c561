#!/bin/bash
# r2r: programmatic dependent launch: parity + A/B timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r2r_pytest.log 2>&1
echo "pytest exit=$?"; tail -n 6 gpurun_out/r2r_pytest.log
for b in 0 1; do
  EDV_PDL=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2r_bench_p$b.log 2> gpurun_out/r2r_bench_p$b.err
  echo "bench EDV_PDL=$b exit=$?"; tail -c 300 gpurun_out/r2r_bench_p$b.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2r_bench_p$b.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','cuda_graphs')}, d['e2e']['value'])
for k,v in d['extra'].items(): print(k, {kk:v.get(kk) for kk in ('ms_per_step','frames_per_s','seconds','cuda_graphs')})
PY
done
