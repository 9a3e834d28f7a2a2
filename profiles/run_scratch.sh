O=gpurun_out; mkdir -p $O
for v in "" _el _els; do echo "== variant '$v'"; MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda$v.so timeout 300 python profiles/ab_encode.py 1 4096 1048576 json 3 2>&1 | tail -1; done | tee $O/evict.log
