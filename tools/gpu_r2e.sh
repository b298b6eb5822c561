#!/bin/bash
# ncu --set full of the hot ops (one warm launch each), after a plain run of the same command
mkdir -p gpurun_out
python tools/prof_ops.py --range fc1 qkv fc2 attn rcu head ln > gpurun_out/r2e_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2e_ops \
    python tools/prof_ops.py --range fc1 qkv fc2 attn rcu head ln > gpurun_out/r2e_ncu.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/r2e_ncu.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --tb=short -k "saturate or headroom" > gpurun_out/r2e_pytest.log 2>&1
echo "pytest exit=$?"; tail -n 5 gpurun_out/r2e_pytest.log
