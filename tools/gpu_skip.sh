#!/bin/bash
# timing experiments on the tensor assignment kernel: VQB_TC_SKIP variants (results are wrong, only the time matters)
for v in 0 3; do
  VQB_TC_SKIP=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-paths --cpu-sample 0 --kmeans-iters 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('skip=$v', 'ms', round(d['ms_per_step'],4))"
done
