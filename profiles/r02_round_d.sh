#!/bin/bash
# round 2, GPU visit D: decoder with lexer warps v2 (16-bit descriptors, headers from the ring), pipelined e2e
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
{
echo "== decode (lexer warps v2)"; timeout 300 python profiles/time_decode.py 2>&1 | tail -2
echo "== decode, copiers disabled (parser + lexers alone)"; MINLZ_NO_CHECK=1 MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda_nocopy.so timeout 300 python profiles/time_decode.py 2>&1 | tail -1
} | tee $O/decode_lexer2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_pc -c 1 -o $O/dec_pc_lexer2_full -f python profiles/prof_run.py 4096 > $O/ncu_dec.log 2>&1
tail -2 $O/ncu_dec.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -3 $O/bench_default.err; cat $O/bench_default.json
