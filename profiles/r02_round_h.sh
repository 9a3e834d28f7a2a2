#!/bin/bash
# round 2, GPU visit H: the record run -- tests, every bench workload with its reference arm, launch list,
# full ncu captures of the two headline kernels, sanitizers
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py --steps 10 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err; cat $O/bench_default.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; cat $O/bench_reference.json
timeout 900 python bench.py --workload stream > $O/bench_stream.json 2> $O/bench_stream.err; tail -2 $O/bench_stream.err; cat $O/bench_stream.json
timeout 600 python bench.py --workload stream --impl reference > $O/bench_stream_ref.json 2> $O/bench_stream_ref.err; cat $O/bench_stream_ref.json
timeout 900 python bench.py --workload stream --level 2 > $O/bench_stream_l2.json 2> $O/bench_stream_l2.err; tail -2 $O/bench_stream_l2.err; cat $O/bench_stream_l2.json
timeout 600 python bench.py --workload stream --level 2 --impl reference > $O/bench_stream_l2_ref.json 2> $O/bench_stream_l2_ref.err; cat $O/bench_stream_l2_ref.json
timeout 900 python bench.py --workload sweep --steps 2 --warmup 1 > $O/bench_sweep.json 2> $O/bench_sweep.err; tail -2 $O/bench_sweep.err; cat $O/bench_sweep.json
timeout 600 python bench.py --workload sweep --impl reference --steps 2 --warmup 1 > $O/bench_sweep_ref.json 2> $O/bench_sweep_ref.err; cat $O/bench_sweep_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_l|decode_pc|scan_lengths|pack_blocks|crc32c" --csv --log-file $O/launch_list.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_l1 -c 1 -o $O/enc_l1_final -f python profiles/prof_run.py 4096 > $O/ncu_enc.log 2>&1; tail -1 $O/ncu_enc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_pc -c 1 -o $O/dec_pc_final -f python profiles/prof_run.py 4096 > $O/ncu_dec.log 2>&1; tail -1 $O/ncu_dec.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/sanitize_run.py > $O/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitize_memcheck.log; tail -4 $O/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 40 python profiles/sanitize_run.py quick > $O/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitize_racecheck.log; tail -4 $O/sanitize_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python profiles/sanitize_run.py quick > $O/sanitize_synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/sanitize_synccheck.log; tail -4 $O/sanitize_synccheck.log
