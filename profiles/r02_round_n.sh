#!/bin/bash
# round 2, GPU visit N: larger fuzz of every encoder (stress inputs for the L2 walk's patching), memcheck of the host API tests
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python profiles/fuzz_gpu.py 5000 > $O/fuzz_gpu.log 2>&1; echo "fuzz rc=$?" >> $O/fuzz_gpu.log; tail -9 $O/fuzz_gpu.log
timeout 900 python profiles/fuzz_gpu.py 3000 7 > $O/fuzz_gpu2.log 2>&1; echo "fuzz rc=$?" >> $O/fuzz_gpu2.log; tail -3 $O/fuzz_gpu2.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_host_api.py -x -q > $O/sanitize_memcheck_host_api.log 2>&1; echo "memcheck rc=$?" >> $O/sanitize_memcheck_host_api.log; tail -5 $O/sanitize_memcheck_host_api.log
