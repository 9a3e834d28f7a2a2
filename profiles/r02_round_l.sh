#!/bin/bash
# round 2, GPU visit L: L2 walk v3 (source prefetch, lazy repeat check), whole-call launch gate (pipelined e2e)
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
{
for cfg in "4096 1048576 json" "2048 2097152 log" "512 8388608 text" "512 8388608 binary"; do
  echo "== L2 probe-window walk v3: $cfg"; timeout 300 python profiles/ab_encode.py 2 $cfg 2 2>&1 | tail -2
done
} | tee $O/l2_walk3.log
timeout 900 python bench.py --steps 6 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err; cat $O/bench_default.json
# racecheck with the lexers' ring refill as bulk copies (product) and as per-lane cp.async (variant build)
timeout 900 compute-sanitizer --tool racecheck --print-limit 40 python profiles/sanitize_run.py quick > $O/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitize_racecheck.log; tail -3 $O/sanitize_racecheck.log
MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda_decnobulk.so timeout 900 compute-sanitizer --tool racecheck --print-limit 40 python profiles/sanitize_run.py quick > $O/sanitize_racecheck_cpasync.log 2>&1; echo "racecheck rc=$?" >> $O/sanitize_racecheck_cpasync.log; tail -3 $O/sanitize_racecheck_cpasync.log
MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda_decnobulk.so timeout 600 compute-sanitizer --tool synccheck --print-limit 10 python profiles/sanitize_run.py quick > $O/sanitize_synccheck_cpasync.log 2>&1; echo "synccheck rc=$?" >> $O/sanitize_synccheck_cpasync.log; tail -3 $O/sanitize_synccheck_cpasync.log
echo "== decode, cp.async lexers"; MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda_decnobulk.so timeout 300 python profiles/time_decode.py 2>&1 | tail -2
echo "== decode, bulk-copy lexers"; timeout 300 python profiles/time_decode.py 2>&1 | tail -2
