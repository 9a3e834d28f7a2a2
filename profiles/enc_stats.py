"""Walk counters of the L1 encoder (profiling build with -DMZ_ENC_STATS):
batches (DRAM round trips), steps replayed per batch, why a batch ended.
usage: MINLZ_CUDA_SO=minlz_b200/libminlz_cuda_stats.so python profiles/enc_stats.py [flavor] [blocks] [kind]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import minlz_b200 as mz
from minlz_b200 import _lib
import synth

flavor = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 512
kind = sys.argv[3] if len(sys.argv) > 3 else "json"
bs = 1 << 20
dev = torch.device("cuda:0")
src = synth.make_blocks(kind, nblk, bs, device=dev).reshape(-1)
soff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * bs
cap = bs + 16
eoff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * cap
enc = torch.empty(nblk * cap, dtype=torch.uint8, device=dev)
out_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
mz.set_encoder_flavor(flavor)
lib = _lib.load()
out = (C.c_ulonglong * 16)()
lib.mzcu_debug_enc_stats(out)
mz.encode_blocks_dev(src, soff, enc, eoff, out_len, 1)
torch.cuda.synchronize()
lib.mzcu_debug_enc_stats(out)
names = ["batches", "rematch steps", "search steps", "end: chain left window", "end: search at window edge",
         "end: repeat check beyond snapshot", "forwarded probes", "repeats", "fwd extension loads", "back extension loads"]
for i, nm in enumerate(names):
    print("%-36s %12d  per block %10.1f  per batch %.3f" % (nm, out[i], out[i] / nblk, out[i] / max(1, out[0])))
