#!/bin/bash
# Quick GPU visit while iterating on the tensor kernel: tensor tests, short bench, optional extras ("$@").
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor.py -x -q -s > gpurun_out/pytest_tensor.log 2>&1; echo "pytest tensor rc=$?"
grep -E "max err|passed|failed|Error|error" gpurun_out/pytest_tensor.log | tail -20
timeout 300 python bench.py --steps 20 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
    print('encode ms/step', d['ms_per_step'], 'Mvec/s', d['value'], 'kmeans', d.get('kmeans'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_quick.err').read()[-2000:])
PY
for c in "$@"; do echo "== $c"; timeout 600 bash -c "$c"; done
