O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flavor_amd64.py -m gpu -x -q > $O/t12.log 2>&1; tail -3 $O/t12.log
for v in "" _pf1 _pf2; do echo "== variant '$v'"; MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda$v.so timeout 300 python profiles/ab_encode.py 1 4096 1048576 json 3 2>&1 | tail -2; done | tee $O/pf.log
