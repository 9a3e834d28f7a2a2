#!/bin/bash
# round 2, GPU visit J: L2 walk with snapshot-decided lengths + fused extension; pinned CRC staging (stream e2e);
# decode lexer frontier fixes under racecheck / synccheck
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
{
for cfg in "4096 1048576 json" "2048 2097152 log" "512 8388608 text" "512 8388608 binary"; do
  echo "== L2 probe-window walk v2: $cfg"; timeout 300 python profiles/ab_encode.py 2 $cfg 2 2>&1 | tail -2
done
echo "== decode"; timeout 300 python profiles/time_decode.py 2>&1 | tail -2
} | tee $O/l2_walk2.log
timeout 900 python bench.py --workload stream --no-cpu > $O/bench_stream.json 2> $O/bench_stream.err; tail -2 $O/bench_stream.err; cat $O/bench_stream.json
timeout 900 python bench.py --workload stream --level 2 --no-cpu > $O/bench_stream_l2.json 2> $O/bench_stream_l2.err; tail -2 $O/bench_stream_l2.err; cat $O/bench_stream_l2.json
timeout 900 compute-sanitizer --tool racecheck --print-limit 40 python profiles/sanitize_run.py quick > $O/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitize_racecheck.log; tail -4 $O/sanitize_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python profiles/sanitize_run.py quick > $O/sanitize_synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/sanitize_synccheck.log; tail -4 $O/sanitize_synccheck.log
