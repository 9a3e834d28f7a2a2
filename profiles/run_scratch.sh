O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_cpp_host.py -m gpu -x -q > $O/t21.log 2>&1; tail -25 $O/t21.log
