#!/bin/bash
# round 2, GPU visit F: LevelBalanced with lane-parallel match indexing; decode lexers with bulk copies
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
{
echo "== L2 4096 x 1 MiB json"; timeout 300 python profiles/ab_encode.py 2 4096 1048576 json 3 2>&1 | tail -2
echo "== L2 2048 x 2 MiB log"; timeout 300 python profiles/ab_encode.py 2 2048 2097152 log 3 2>&1 | tail -2
echo "== L2 512 x 8 MiB text"; timeout 300 python profiles/ab_encode.py 2 512 8388608 text 2 2>&1 | tail -2
echo "== decode (bulk-copy lexers)"; timeout 300 python profiles/time_decode.py 2>&1 | tail -2
} | tee $O/l2_parallel_index.log
