#!/bin/bash
# round 2, GPU visit W: final tree with one-warp CTAs: full GPU test suite, default bench, ncu --set full of encode_l1
set -u
O=gpurun_out
mkdir -p $O/final2
timeout 120 python -m pytest tests -m gpu -x -q > $O/final2/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/final2/pytest_gpu.log; tail -3 $O/final2/pytest_gpu.log
timeout 100 python bench.py --steps 5 --warmup 3 > $O/final2/bench_default.json 2> $O/final2/bench_default.err; tail -2 $O/final2/bench_default.err
grep '^{' $O/final2/bench_default.json | python -c "
import sys, json
b = json.loads(sys.stdin.read())
print('value', b['value'], b['ms'], 'roofline', b['roofline'], 'e2e', {k: b['e2e'].get(k) for k in ('value', 'serial_calls', 'pipelined_by_jobs_in_flight', 'ms_per_step')}, b['parity'][:80])"
timeout 70 ncu --set full --clock-control none --import-source on -k regex:encode_l1 -c 1 -o $O/final2/enc_l1_w1 -f python profiles/prof_run.py 4096 > $O/final2/ncu_enc.log 2>&1; tail -1 $O/final2/ncu_enc.log
