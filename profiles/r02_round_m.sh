#!/bin/bash
# round 2, GPU visit M: encode upload yields to pending decode uploads (pipelined e2e)
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 900 python bench.py --steps 6 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err; grep '^{' $O/bench_default.json | python -c "
import sys, json
b = json.loads(sys.stdin.read())
print('value', b['value'], b['ms'], 'e2e', {k: b['e2e'][k] for k in ('value', 'serial_calls', 'pipelined_calls', 'ms_per_step', 'serial_ms_per_step')})"
