#!/bin/bash
# Round-2 evidence visit (1 GPU): bench (both arms), ncu launch list, ncu full capture of the tensor kernel and of the
# other kernels, hand-off timeline, stated-scale configs.  Outputs under gpurun_out/; copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -c 1700 gpurun_out/r02_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err; cat gpurun_out/r02_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --kmeans-iters 3 --cpu-sample 0 --no-e2e --no-paths --no-clock-probe > gpurun_out/bench_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv 14 | tee gpurun_out/r02_launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_assign -s 4 -c 1 -f -o gpurun_out/r02_tc_assign \
  python bench.py --steps 2 --warmup 1 --kmeans-iters 0 --cpu-sample 0 --no-e2e --no-paths --no-clock-probe > gpurun_out/bench_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_update_tiles|k_update_ordered|k_assign_l1_tiles" -c 3 -f -o gpurun_out/r02_others \
  python tools/probe_r02.py > gpurun_out/probe_r02.log 2>&1; tail -2 gpurun_out/probe_r02.log
timeout 300 python tools/tc_timeline.py > gpurun_out/r02_tc_timeline.txt 2>&1; tail -2 gpurun_out/r02_tc_timeline.txt
for c in metric100M c2 c5b c5a; do timeout 600 python bench.py --config $c 2>> gpurun_out/r02_bench.err | tee -a gpurun_out/r02_configs_n1.json; done
ls -la gpurun_out | tail -20
