#!/bin/bash
# r2n: flash-attention variants: threads per row x MUFU ping-pong x polynomial share
mkdir -p gpurun_out
rm -f gpurun_out/r2n_fa_variants.txt
for cfg in "1 0 0" "2 0 0" "2 1 0" "2 1 4"; do
  set -- $cfg
  echo "=== EDV_FA_SPLIT=$1 EDV_FA_PP=$2 EDV_FA_POLY=$3" >> gpurun_out/r2n_fa_variants.txt
  EDV_FA_SPLIT=$1 EDV_FA_PP=$2 EDV_FA_POLY=$3 timeout 60 python tools/fa_timeline.py >> gpurun_out/r2n_fa_variants.txt 2>&1 || { echo "variant $cfg failed"; }
done
grep -E "===|per launch|iteration 5|rror" gpurun_out/r2n_fa_variants.txt
for cfg in "2 1 0" "2 1 4"; do
  set -- $cfg
  EDV_FA_SPLIT=$1 EDV_FA_PP=$2 EDV_FA_POLY=$3 timeout 100 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "attention" 2>&1 | tail -n 3
done
