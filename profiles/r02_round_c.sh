#!/bin/bash
# round 2, GPU visit C: decoder with lexer warps (parity, timing, parser-alone experiment, ncu)
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
{
echo "== decode (lexer warps)"; timeout 300 python profiles/time_decode.py 2>&1 | tail -2
echo "== decode, copiers disabled (parser + lexers alone)"; MINLZ_NO_CHECK=1 MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda_nocopy.so timeout 300 python profiles/time_decode.py 2>&1 | tail -1
echo "== encode default (1-bit smem tags)"; timeout 300 python profiles/ab_encode.py 1 4096 1048576 json 3 2>&1 | tail -2
} | tee $O/decode_lexer.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_pc -c 1 -o $O/dec_pc_lexer_full -f python profiles/prof_run.py 4096 > $O/ncu_dec.log 2>&1
tail -2 $O/ncu_dec.log
