#!/bin/bash
# round 2, GPU visit S: final tree (bench.py with three jobs in flight in the e2e leg): tests, smoke,
# default bench, stream benches; outputs go to profiles/r02_record/*_final.*
set -u
O=gpurun_out
mkdir -p $O/final
timeout 600 python -m pytest tests -m gpu -x -q > $O/final/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/final/pytest_gpu.log; tail -3 $O/final/pytest_gpu.log
python __graft_entry__.py --smoke > $O/final/smoke.log 2>&1; tail -1 $O/final/smoke.log
B="timeout 600 python bench.py"
show() { grep '^{' $1 | python -c "
import sys, json
b = json.loads(sys.stdin.read())
print('value', b['value'], 'e2e', {k: b['e2e'].get(k) for k in ('value', 'serial_calls', 'pipelined_calls', 'pipelined_by_jobs_in_flight', 'ms_per_step')})"; }
$B --steps 10 --warmup 5 > $O/final/bench_default.json 2> $O/final/bench_default.err; tail -2 $O/final/bench_default.err; show $O/final/bench_default.json
$B --workload stream > $O/final/bench_stream.json 2> $O/final/bench_stream.err; tail -2 $O/final/bench_stream.err; show $O/final/bench_stream.json
$B --workload stream --level 2 > $O/final/bench_stream_l2.json 2> $O/final/bench_stream_l2.err; tail -2 $O/final/bench_stream_l2.err; show $O/final/bench_stream_l2.json
