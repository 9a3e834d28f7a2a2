#!/bin/bash
# round 2, GPU visit O: short-table tag filter in the L2 walk (variant build): parity + A/B
set -u
O=gpurun_out
mkdir -p $O
V=$PWD/minlz_b200/libminlz_cuda_l2stags.so
MINLZ_CUDA_SO=$V timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_l2stags.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_l2stags.log
tail -5 $O/pytest_gpu_l2stags.log
MINLZ_CUDA_SO=$V timeout 600 python profiles/fuzz_gpu.py 2000 11 > $O/fuzz_l2stags.log 2>&1; echo "fuzz rc=$?" >> $O/fuzz_l2stags.log; tail -3 $O/fuzz_l2stags.log
{
for cfg in "4096 1048576 json" "2048 2097152 log" "512 8388608 text"; do
  echo "== L2 walk + short-table tags: $cfg"; MINLZ_CUDA_SO=$V timeout 300 python profiles/ab_encode.py 2 $cfg 2 2>&1 | tail -2
  echo "== L2 walk (product): $cfg"; timeout 300 python profiles/ab_encode.py 2 $cfg 2 2>&1 | tail -2
done
} | tee $O/l2_stags.log
