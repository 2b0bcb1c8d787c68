#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch lists, ncu full capture of the tensor assignment kernel,
# hand-off timeline.  Outputs under gpurun_out/; copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --kmeans-iters 2 --cpu-sample 0 --no-e2e --no-paths --no-clock-probe > gpurun_out/bench_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv 14
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/probe_launches.csv \
  python tools/probe_paths.py > gpurun_out/probe.log 2>&1
python tools/launch_summary.py gpurun_out/probe_launches.csv 14
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_assign -s 4 -c 2 -f -o gpurun_out/tc_assign \
  python bench.py --steps 2 --warmup 1 --kmeans-iters 0 --cpu-sample 0 --no-e2e --no-paths --no-clock-probe > gpurun_out/bench_ncu2.log 2>&1
timeout 300 python tools/tc_timeline.py > gpurun_out/tc_timeline.txt 2>&1; tail -3 gpurun_out/tc_timeline.txt
timeout 600 ncu --set full --clock-control none -k regex:"k_f32_to_u8|k_u8_to_f32|k_assign_exact|k_pq_decode|k_colsum|k_tsvq_encode|k_radix_scatter|k_chain_sums" -c 45 -f -o gpurun_out/others python tools/probe_paths.py > gpurun_out/probe_full.log 2>&1; tail -2 gpurun_out/probe_full.log
ls -la gpurun_out
