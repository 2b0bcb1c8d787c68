#!/bin/bash
# 8-GPU visit (charged 8x): concurrent PCIe probe, the driver's bench line at N = 8 (strong-scaled k-means inside), stated-scale configs.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29551 tools/pcie_probe_multi.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tee gpurun_out/r02_pcie_n$N.txt
timeout 400 $TR --master-port 29552 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -c 1600 gpurun_out/r02_bench_n$N.json
for c in metric100M c5a c5b c2; do timeout 300 $TR --master-port 29553 bench.py --gpus $N --config $c 2>> gpurun_out/r02_bench_n$N.err | tee -a gpurun_out/r02_configs_n$N.json; done
tail -3 gpurun_out/r02_bench_n$N.err
