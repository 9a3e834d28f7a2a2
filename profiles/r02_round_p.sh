#!/bin/bash
# round 2, GPU visit P: last sanity run of the final tree (host API tests, smoke, default bench)
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_host_api.py tests/test_stream.py -m gpu -x -q > $O/pytest_p.log 2>&1; echo "pytest rc=$?" >> $O/pytest_p.log; tail -3 $O/pytest_p.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 4 --warmup 3 > $O/bench_p.json 2> $O/bench_p.err; tail -2 $O/bench_p.err; grep '^{' $O/bench_p.json | python -c "
import sys, json
b = json.loads(sys.stdin.read())
print('value', b['value'], b['ms'], 'e2e', {k: b['e2e'][k] for k in ('value', 'serial_calls', 'pipelined_calls', 'ms_per_step', 'serial_ms_per_step')}, b['parity'][:90])"
