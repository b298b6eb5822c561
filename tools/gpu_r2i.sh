#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "residual_layernorm" > gpurun_out/r2i_pytest_ln.log 2>&1
echo "gemm_ln pytest exit=$?"; tail -n 12 gpurun_out/r2i_pytest_ln.log
