#!/bin/bash
# round 2, GPU visit A: parity with the tag filter, A/B against the unfiltered build, ncu, sanitizers
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
for lv in 1 -1; do
  echo "== tags level $lv"; timeout 300 python profiles/ab_encode.py $lv 4096 1048576 json 3 2>&1 | tail -2
  echo "== notags level $lv"; MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda_notags.so timeout 300 python profiles/ab_encode.py $lv 4096 1048576 json 3 2>&1 | tail -2
done | tee $O/ab_tags.log
for k in log text; do
  echo "== tags $k"; timeout 300 python profiles/ab_encode.py 1 2048 2097152 $k 3 2>&1 | tail -2
  echo "== notags $k"; MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda_notags.so timeout 300 python profiles/ab_encode.py 1 2048 2097152 $k 3 2>&1 | tail -2
done | tee -a $O/ab_tags.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_l1 -c 1 -o $O/enc_l1_tags_full -f python profiles/prof_run.py 4096 > $O/ncu_enc.log 2>&1
tail -2 $O/ncu_enc.log
timeout 700 compute-sanitizer --tool memcheck --print-limit 20 python profiles/sanitize_run.py quick > $O/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitize_memcheck.log
tail -8 $O/sanitize_memcheck.log
timeout 700 compute-sanitizer --tool racecheck --print-limit 20 python profiles/sanitize_run.py quick > $O/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitize_racecheck.log
tail -8 $O/sanitize_racecheck.log
