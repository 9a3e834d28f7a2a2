#!/bin/bash
# round 2, 8-GPU visit: the N=8 bench line with the funnel sub-object, the reference arm on the same box,
# the PCIe floor with 8 concurrent processes, the stream workload at N=8
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo8.txt 2>&1
nproc > $O/nproc8.txt; numactl -H > $O/numa8.txt 2>&1 || lscpu | grep -i numa > $O/numa8.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 profiles/pcie_probe_ranks.py > $O/pcie_floor_8.txt 2>&1; tail -2 $O/pcie_floor_8.txt
timeout 300 python profiles/pcie_probe_ranks.py > $O/pcie_floor_1.txt 2>&1; tail -2 $O/pcie_floor_1.txt
timeout 900 $TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; tail -3 $O/bench_8gpu.err; grep '^{' $O/bench_8gpu.json
timeout 600 $TR --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --impl reference > $O/bench_8gpu_ref.json 2> $O/bench_8gpu_ref.err; grep '^{' $O/bench_8gpu_ref.json
timeout 900 $TR --master-port 29524 bench.py --gpus 8 --steps 3 --warmup 2 --workload stream --no-cpu > $O/bench_8gpu_stream.json 2> $O/bench_8gpu_stream.err; tail -3 $O/bench_8gpu_stream.err; grep '^{' $O/bench_8gpu_stream.json
timeout 900 $TR --master-port 29525 bench.py --gpus 8 --steps 2 --warmup 1 --workload sweep --no-gate > $O/bench_8gpu_sweep.json 2> $O/bench_8gpu_sweep.err; tail -3 $O/bench_8gpu_sweep.err; grep '^{' $O/bench_8gpu_sweep.json
