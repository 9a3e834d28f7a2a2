#!/bin/bash
# round 2, GPU visit X (last 40 s of the budget): one-pass ncu metrics of the one-warp-per-CTA encode_l1 kernel
O=gpurun_out
mkdir -p $O/final2
timeout 36 ncu --metrics sm__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct --clock-control none -k regex:encode_l1 -c 1 python profiles/prof_run.py 4096 > $O/final2/ncu_metrics_enc_l1_w1.txt 2>&1
grep -E "inst_executed|time_duration|dram__|issue_active|ok " $O/final2/ncu_metrics_enc_l1_w1.txt
