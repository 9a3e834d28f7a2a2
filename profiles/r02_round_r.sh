#!/bin/bash
# round 2, final record run on one GPU: tests, every bench workload with its reference arm, launch list,
# full ncu captures of the three kernels, memcheck
set -u
O=gpurun_out
mkdir -p $O/record
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/record/smi.txt 2>&1
nproc > $O/record/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/record/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/record/pytest_gpu.log; tail -3 $O/record/pytest_gpu.log
python __graft_entry__.py --smoke > $O/record/smoke.log 2>&1; tail -1 $O/record/smoke.log
B="timeout 900 python bench.py"
$B --steps 10 --warmup 5 > $O/record/bench_default.json 2> $O/record/bench_default.err; tail -2 $O/record/bench_default.err; grep '^{' $O/record/bench_default.json | cut -c1-400
$B --impl reference --steps 10 --warmup 5 > $O/record/bench_reference.json 2>/dev/null; grep '^{' $O/record/bench_reference.json | cut -c1-300
$B --workload stream > $O/record/bench_stream.json 2> $O/record/bench_stream.err; tail -2 $O/record/bench_stream.err
$B --workload stream --impl reference > $O/record/bench_stream_ref.json 2>/dev/null
$B --workload stream --level 2 > $O/record/bench_stream_l2.json 2> $O/record/bench_stream_l2.err; tail -2 $O/record/bench_stream_l2.err
$B --workload stream --level 2 --impl reference > $O/record/bench_stream_l2_ref.json 2>/dev/null
$B --workload sweep --steps 2 --warmup 1 > $O/record/bench_sweep.json 2> $O/record/bench_sweep.err; tail -2 $O/record/bench_sweep.err
$B --workload sweep --impl reference --steps 2 --warmup 1 > $O/record/bench_sweep_ref.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_l|decode_pc|scan_lengths|pack_blocks|crc32c" --csv --log-file $O/record/launch_list.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $O/record/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_l1 -c 1 -o $O/record/enc_l1 -f python profiles/prof_run.py 4096 > $O/record/ncu_enc.log 2>&1; tail -1 $O/record/ncu_enc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_pc -c 1 -o $O/record/dec_pc -f python profiles/prof_run.py 4096 > $O/record/ncu_dec.log 2>&1; tail -1 $O/record/ncu_dec.log
MINLZ_LEVEL=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_l2 -c 1 -o $O/record/enc_l2 -f python profiles/prof_run.py 4096 > $O/record/ncu_l2.log 2>&1; tail -1 $O/record/ncu_l2.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/sanitize_run.py > $O/record/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/record/sanitize_memcheck.log; tail -3 $O/record/sanitize_memcheck.log
