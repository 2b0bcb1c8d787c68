#!/bin/bash
# A/B visit: GPU parity tests, then bench.py under a list of environment variants ("NAME=VAL ..." per argument).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
i=0
for v in "$@"; do
  i=$((i+1))
  echo "== variant $i: $v"
  env $v timeout 600 python bench.py ${BENCH_ARGS:-} > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err; echo "rc=$?"
  python - "$i" <<'PY'
import json,sys
i=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/ab_{i}.json').read().strip().splitlines()[-1])
    k=d.get('kmeans') or {}
    print('  encode ms', round(d['ms_per_step'],4), '| e2e', (d.get('e2e') or {}).get('value'), '| kmeans ms/iter', k.get('ms_per_iter'), 'call ms', k.get('train_call_ms'), 'iter/s', k.get('value'))
except Exception as e:
    print('  parse failed', e); print(open(f'gpurun_out/ab_{i}.err').read()[-1500:])
PY
done
