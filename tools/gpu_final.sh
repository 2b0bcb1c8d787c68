#!/bin/bash
# Slim closing visit: bench (both arms), probe launch list, ncu --set full of the non-tensor kernels exported as CSV on the box.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/probe_launches.csv \
  python tools/probe_paths.py > gpurun_out/probe.log 2>&1
python tools/launch_summary.py gpurun_out/probe_launches.csv 24
timeout 600 ncu --set full --clock-control none -k regex:"k_f32_to_u8|k_u8_to_f32|k_assign_exact|k_pq_decode|k_colsum_w|k_tsvq_encode|k_radix_scatter|k_chain_sums" -c 14 -f -o /tmp/others \
  python tools/probe_paths.py > gpurun_out/probe_full.log 2>&1
ncu -i /tmp/others.ncu-rep --page raw --csv > gpurun_out/others_raw.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/others_raw.csv > gpurun_out/others_pick.txt; head -40 gpurun_out/others_pick.txt
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
