#!/bin/bash
# Diagnostics visit: PCIe probe, traced k-means iterations, bench with per-path numbers.
set -u
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/host.txt
nvidia-smi topo -m >> gpurun_out/host.txt 2>&1
timeout 300 python tools/pcie_probe.py > gpurun_out/pcie.txt 2>&1; cat gpurun_out/pcie.txt
VQB_TRACE=1 timeout 300 python tools/diag_train.py > gpurun_out/diag_train.txt 2>&1; tail -30 gpurun_out/diag_train.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
