#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "head" > gpurun_out/r3h_pytest_ops.log 2>&1
echo "head op tests exit=$?"; tail -n 3 gpurun_out/r3h_pytest_ops.log
timeout 300 python -m pytest tests/test_gpu_forward.py tests/test_gpu_fullsize.py tests/test_gpu_endodac.py -m gpu -q --tb=short -x > gpurun_out/r3h_pytest.log 2>&1
echo "forward tests exit=$?"; tail -n 3 gpurun_out/r3h_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --kernels-out gpurun_out/r3h_bench_kernels.json > gpurun_out/r3h_bench.log 2> gpurun_out/r3h_bench.err
echo "bench exit=$?"; tail -c 300 gpurun_out/r3h_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3h_bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','cuda_graphs')}, d['e2e']['value'], d['clocks'])
print(d['roofline'].get('mufu'))
t=json.load(open('gpurun_out/r3h_bench_kernels.json'))['kernels']
for r in t:
    if r['name'].startswith('head_fused') or r['name'] in ('upsample',): print(r['name'], r['count'], round(1e3*r['ms']/r['count'],1),'us')
PY
