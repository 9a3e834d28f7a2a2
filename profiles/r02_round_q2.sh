#!/bin/bash
# round 2, GPU visit Q2: e2e pipelined calls, asymptote (16 batches) with 2 / 3 / 4 jobs in flight
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --steps 3 --warmup 3 --no-gate --pipeline-ahead 1,2,3 --pipeline-iters 16 > $O/bench_q2.json 2> $O/bench_q2.err; tail -2 $O/bench_q2.err
grep '^{' $O/bench_q2.json | python -c "
import sys, json
b = json.loads(sys.stdin.read())
print('value', b['value'], 'e2e', {k: b['e2e'].get(k) for k in ('value', 'serial_calls', 'pipelined_calls', 'pipelined_by_jobs_in_flight', 'ms_per_step')})"
