#!/bin/bash
# round 2, 2-GPU visit: the N>1 bench line (funnel sub-object, e2e), multi-device C-ABI calls on two real devices
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo2.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_host_api.py -x -q > $O/pytest_host_api_2gpu.log 2>&1; echo "rc=$?" >> $O/pytest_host_api_2gpu.log; tail -4 $O/pytest_host_api_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; tail -5 $O/bench_2gpu.err; cat $O/bench_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --impl reference > $O/bench_2gpu_ref.json 2> $O/bench_2gpu_ref.err; tail -3 $O/bench_2gpu_ref.err; cat $O/bench_2gpu_ref.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 2 --workload stream > $O/bench_2gpu_stream.json 2> $O/bench_2gpu_stream.err; tail -3 $O/bench_2gpu_stream.err; cat $O/bench_2gpu_stream.json
