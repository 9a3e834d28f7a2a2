#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
{
for cfg in "4096 1048576 json" "2048 2097152 log" "512 8388608 text" "512 8388608 binary"; do
  echo "== L2 probe-window walk v2: $cfg"; timeout 300 python profiles/ab_encode.py 2 $cfg 2 2>&1 | tail -2
done
} | tee $O/l2_walk2.log
MINLZ_LEVEL=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_l2 -c 1 -o $O/enc_l2_walk_full -f python profiles/prof_run.py 2048 > $O/ncu_l2.log 2>&1; tail -1 $O/ncu_l2.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 40 python profiles/sanitize_run.py quick > $O/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitize_racecheck.log; tail -4 $O/sanitize_racecheck.log
