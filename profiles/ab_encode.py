"""A/B timing of the encode kernels on one batch (CUDA events, interleaved runs):
Go flavour vs amd64 flavour at a level.  usage: python profiles/ab_encode.py [level] [blocks] [block_size] [kind] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import minlz_b200 as mz
import synth

level = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
kind = sys.argv[4] if len(sys.argv) > 4 else "json"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dev = torch.device("cuda:0")
src = synth.make_blocks(kind, nblk, bs, device=dev).reshape(-1)
soff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * bs
cap = bs + 16
eoff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * cap
enc = torch.empty(nblk * cap, dtype=torch.uint8, device=dev)
out_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {0: [], 1: []}
for r in range(reps + 2):
    for fl in (mz.FlavorGo, mz.FlavorAMD64):
        mz.set_encoder_flavor(fl)
        ev0.record()
        mz.encode_blocks_dev(src, soff, enc, eoff, out_len, level)
        ev1.record()
        torch.cuda.synchronize()
        if r >= 2:
            res[fl].append(ev0.elapsed_time(ev1))
for fl, name in ((0, "go"), (1, "amd64")):
    v = sorted(res[fl])
    print("level %d %s %d x %d %s: median %.2f ms  min %.2f  max %.2f  -> %.1f GB/s" %
          (level, name, nblk, bs, kind, v[len(v) // 2], v[0], v[-1], nblk * bs / v[len(v) // 2] / 1e6))
