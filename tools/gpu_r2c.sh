#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/fullsize_parity.json
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2c_pytest.log 2>&1
echo "pytest exit=$?"; tail -n 25 gpurun_out/r2c_pytest.log
timeout 300 python tools/fa_timeline.py > gpurun_out/r2c_fa_timeline.txt 2>&1
echo "timeline exit=$?"; cat gpurun_out/r2c_fa_timeline.txt
EDV_FA_POLY=4 timeout 300 python tools/fa_timeline.py > gpurun_out/r2c_fa_timeline_pm4.txt 2>&1
cat gpurun_out/r2c_fa_timeline_pm4.txt
