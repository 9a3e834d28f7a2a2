#!/bin/bash
# round 2, GPU visit S2: the final tree on two GPUs (the way the driver launches it)
set -u
O=gpurun_out
mkdir -p $O/final
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/final/bench_2gpu.json 2> $O/final/bench_2gpu.err; tail -2 $O/final/bench_2gpu.err
grep '^{' $O/final/bench_2gpu.json | python -c "
import sys, json
b = json.loads(sys.stdin.read())
print('value', b['value'], 'funnel', (b.get('funnel') or {}).get('value'), 'e2e', {k: b['e2e'].get(k) for k in ('value', 'serial_calls', 'pipelined_calls', 'pipelined_by_jobs_in_flight', 'ms_per_step')})"
