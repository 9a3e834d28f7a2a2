O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flavor_amd64.py tests/test_stream.py -m gpu -x -q > $O/t17.log 2>&1; tail -12 $O/t17.log
timeout 300 python profiles/ab_encode.py 2 4096 1048576 json 3 2>&1 | tail -2 | tee $O/ab17.log
