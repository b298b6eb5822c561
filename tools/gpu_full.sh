#!/bin/bash
# Full GPU parity suite + default bench line (+ kernel table).  Usage: tools/gpu_full.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
rm -f gpurun_out/fullsize_parity.json
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit=$?"; tail -n 15 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernels-out gpurun_out/${tag}_bench_kernels.json > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
echo "bench exit=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_kernels.json"))
print("ms_per_step", d["ms_per_step"])
for r in d["kernels"][:22]:
    print("   %-26s n=%3d %7.1f us %7.1f TF/s %7.1f GB/s" % (r["name"], r["count"] // d["steps"], r["avg_us"], r["tflops"], r["gbs"]))
PY
