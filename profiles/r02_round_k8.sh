#!/bin/bash
# round 2, 8-GPU visit: the N=8 bench line with the funnel sub-object, the reference arm on the same box,
# the PCIe floor with 8 concurrent processes
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo8.txt 2>&1
nproc > $O/nproc8.txt; (numactl -H || lscpu | grep -i numa) > $O/numa8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29521 profiles/pcie_probe_ranks.py > $O/pcie_floor_8.txt 2>&1; tail -2 $O/pcie_floor_8.txt
timeout 120 python profiles/pcie_probe_ranks.py > $O/pcie_floor_1.txt 2>&1; tail -2 $O/pcie_floor_1.txt
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; tail -3 $O/bench_8gpu.err; grep '^{' $O/bench_8gpu.json
timeout 300 $TR --master-port 29523 bench.py --gpus 8 --steps 3 --warmup 3 --impl reference > $O/bench_8gpu_ref.json 2> $O/bench_8gpu_ref.err; grep '^{' $O/bench_8gpu_ref.json
