set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multirank.py -x -q -s 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -12
