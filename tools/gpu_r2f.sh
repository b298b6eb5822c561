#!/bin/bash
mkdir -p gpurun_out
EDV_GEMM_2SM=1000 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernels-out gpurun_out/r2f_bench_kernels_2sm.json > gpurun_out/r2f_bench_2sm.log 2>&1
echo "2sm exit=$?"
python - <<PY
import json
for tag in ("r2d_bench_kernels", "r2f_bench_kernels_2sm"):
    try:
        d = json.load(open("gpurun_out/%s.json" % tag))
    except Exception as e:
        print(tag, e); continue
    print(tag, d["ms_per_step"])
    for r in d["kernels"]:
        if "blk." in r["name"] or "geglu" in r["name"] or "mm.a.qkv" in r["name"]:
            print("   %-24s %7.1f us %7.1f TF/s" % (r["name"], r["avg_us"], r["tflops"]))
PY
