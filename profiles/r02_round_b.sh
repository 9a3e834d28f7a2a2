#!/bin/bash
# round 2, GPU visit B: tag-filter variants (L2 hints, 2-bit, shared-memory 1-bit), new host API tests, new bench legs
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__issue_active.avg.pct,sm__inst_executed.sum
run_variant() { # name, so-suffix, extra env
  local so=$PWD/minlz_b200/libminlz_cuda$2.so
  echo "== $1"
  env $3 MINLZ_CUDA_SO=$so timeout 300 python profiles/ab_encode.py 1 4096 1048576 json 3 2>&1 | tail -2
  env $3 MINLZ_CUDA_SO=$so timeout 300 ncu --metrics $M --clock-control none -k regex:encode_l1 -c 1 python profiles/prof_run.py 4096 2>&1 | grep -E "dram__|lts__|gpu__time|issue_active|inst_executed"
}
{
run_variant "tags4 L2 + hints (default)" "" "X=1"
run_variant "tags4 L2 + hints + 96 MiB persisting set-aside" "" "MINLZ_CUDA_L2_PERSIST_MB=96"
run_variant "tags2 L2 + hints" "_tags2" "X=1"
run_variant "tags 1-bit smem + slot hints" "_tagsmem" "X=1"
run_variant "tags 1-bit smem, no hints" "_tagsmem_nohint" "X=1"
run_variant "no tags + slot hints" "_notags_hint" "X=1"
run_variant "no tags (round-1 kernel)" "_notags" "X=1"
} 2>&1 | tee $O/ab_variants.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -3 $O/bench_default.err; cat $O/bench_default.json
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -3 $O/bench_reference.err; cat $O/bench_reference.json
timeout 900 python bench.py --workload stream > $O/bench_stream.json 2> $O/bench_stream.err; tail -3 $O/bench_stream.err; cat $O/bench_stream.json
timeout 900 python bench.py --workload sweep --steps 2 --warmup 1 > $O/bench_sweep.json 2> $O/bench_sweep.err; tail -3 $O/bench_sweep.err; cat $O/bench_sweep.json
timeout 600 python bench.py --workload sweep --impl reference --steps 2 --warmup 1 > $O/bench_sweep_ref.json 2> $O/bench_sweep_ref.err; tail -3 $O/bench_sweep_ref.err; cat $O/bench_sweep_ref.json
