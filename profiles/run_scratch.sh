O=gpurun_out; mkdir -p $O
echo "full:"; timeout 300 python profiles/time_decode.py | tee $O/dec_time.log
echo "parser only (copiers idle):"; MINLZ_CUDA_SO=$PWD/minlz_b200/libminlz_cuda_nocopy.so timeout 300 python profiles/time_decode.py | tee -a $O/dec_time.log
