#!/bin/bash
# round 2, GPU visit I: LevelBalanced probe-window walk (parity, A/B against the step-at-a-time walk)
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
{
for cfg in "4096 1048576 json" "2048 2097152 log" "512 8388608 text" "512 8388608 binary"; do
  echo "== L2 probe-window walk: $cfg"; timeout 300 python profiles/ab_encode.py 2 $cfg 2 2>&1 | tail -2
  echo "== L2 step walk: $cfg"; MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda_l2step.so timeout 300 python profiles/ab_encode.py 2 $cfg 2 2>&1 | tail -2
done
} | tee $O/l2_walk.log
