O=gpurun_out; mkdir -p $O
MINLZ_LEVEL=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_l2 -c 1 -o $O/enc_l2_full -f python profiles/prof_run.py 2048 > $O/ncu_enc_l2.log 2>&1; tail -2 $O/ncu_enc_l2.log
