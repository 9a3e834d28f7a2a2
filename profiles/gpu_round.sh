#!/bin/bash
# One GPU-box visit: parity tests, bench lines (both arms), ncu launch list, ncu full captures.
# Outputs land in gpurun_out/ (scratch); the summaries worth keeping are copied to profiles/ by hand.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_l|decode_pc|scan_lengths|pack_blocks|crc32c" --csv --log-file $O/launch_list.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_l1 -c 1 -o $O/enc_l1_full -f python profiles/prof_run.py 4096 > $O/ncu_enc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_pc -c 1 -o $O/dec_pc_full -f python profiles/prof_run.py 4096 > $O/ncu_dec.log 2>&1
tail -3 $O/pytest_gpu.log; cat $O/bench_default.json $O/bench_reference.json
for lv in -1 1 2; do timeout 300 python profiles/ab_encode.py $lv 4096 1048576 json 3 2>&1 | tail -2; done | tee $O/ab_levels.log
MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda_stats.so timeout 300 python profiles/enc_stats.py 1 512 json > $O/enc_stats.log 2>&1
