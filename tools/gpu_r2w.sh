#!/bin/bash
# r2w: separable upsample: parity + timing
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "upsample or resize" > gpurun_out/r2w_pytest_ops.log 2>&1
echo "upsample op tests exit=$?"; tail -n 3 gpurun_out/r2w_pytest_ops.log
timeout 400 python -m pytest tests/test_gpu_forward.py tests/test_gpu_fullsize.py tests/test_gpu_endodac.py -m gpu -q --tb=short -x > gpurun_out/r2w_pytest.log 2>&1
echo "forward tests exit=$?"; tail -n 4 gpurun_out/r2w_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --kernels-out gpurun_out/r2w_bench_kernels.json > gpurun_out/r2w_bench.log 2> gpurun_out/r2w_bench.err
echo "bench exit=$?"; tail -c 300 gpurun_out/r2w_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','cuda_graphs')}, d['e2e']['value'])
t=json.load(open('gpurun_out/r2w_bench_kernels.json'))['kernels']
for r in t:
    if r['name'] in ('upsample','layernorm','preprocess','resize_f32'): print(r['name'], r['count'], round(1e3*r['ms']/r['count'],1),'us', round(r.get('gbs',0)),'GB/s')
m=json.load(open('gpurun_out/golden_16bit_margins.json'))
print({k:round(v['rel_max'],5) for k,v in m.items() if k.endswith('fp16')})
PY
