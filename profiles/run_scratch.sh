O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/t8.log 2>&1; tail -15 $O/t8.log
for v in "" _serial; do echo "== variant '$v'"; MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda$v.so timeout 300 python profiles/ab_encode.py 1 4096 1048576 json 3 2>&1 | tail -2; done | tee $O/group.log
MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda.so timeout 300 python profiles/ab_encode.py -1 4096 1048576 json 3 2>&1 | tail -2 | tee -a $O/group.log
