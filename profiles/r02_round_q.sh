#!/bin/bash
# round 2, GPU visit Q: e2e with three jobs in flight (bench.py pipelined leg, ahead = 2)
set -u
O=gpurun_out
mkdir -p $O
show() { grep '^{' $1 | python -c "
import sys, json
b = json.loads(sys.stdin.read())
print('value', b['value'], 'e2e', {k: b['e2e'].get(k) for k in ('value', 'serial_calls', 'pipelined_calls', 'pipelined_by_jobs_in_flight', 'ms_per_step')})"; }
timeout 600 python bench.py --steps 4 --warmup 3 > $O/bench_q.json 2> $O/bench_q.err; tail -2 $O/bench_q.err; show $O/bench_q.json
timeout 600 python bench.py --workload stream --steps 3 --warmup 3 > $O/bench_q_stream.json 2> $O/bench_q_stream.err; tail -2 $O/bench_q_stream.err; show $O/bench_q_stream.json
