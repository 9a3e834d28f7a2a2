O=gpurun_out; mkdir -p $O
for cfg in "--blocks 2048 --block-size 2097152 --kind log --level 1" "--blocks 2048 --block-size 2097152 --kind log --level 2" "--blocks 512 --block-size 8388608 --kind text --level 2" "--blocks 512 --block-size 8388608 --kind binary --level 2" "--blocks 512 --block-size 8388608 --kind text --level 1" "--blocks 4096 --block-size 1048576 --kind json --level -1"; do
  echo "== $cfg"; timeout 600 python bench.py $cfg --steps 2 --warmup 3 --no-cpu --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms','ratio','encode_gbps','decode_gbps')}), d['config']['encoder_flavor'])"
done 2>&1 | tee $O/shapes.log
