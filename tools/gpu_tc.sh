#!/bin/bash
# Tensor-kernel visit: tensor parity tests, then tools/tc_quick.py (variants "$@").
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor.py -x -q > gpurun_out/pytest_tensor.log 2>&1; echo "pytest tensor rc=$?"
tail -5 gpurun_out/pytest_tensor.log
timeout 300 python tools/tc_quick.py "$@" 2>&1 | tee gpurun_out/tc_quick.txt
