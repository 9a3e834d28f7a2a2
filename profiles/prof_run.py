"""Small driver for ncu captures: runs one encode / pack / decode pass over a
synthetic batch through the C ABI (no timing; numbers under ncu are not bench
values).  usage: python profiles/prof_run.py [blocks] [block_size] [kind] [passes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import minlz_b200 as mz
import synth

nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
kind = sys.argv[3] if len(sys.argv) > 3 else "json"
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 1
# the flavour bench.py runs by default: the reference's amd64 assembly (MINLZ_FLAVOR=go for the other one)
mz.set_encoder_flavor(mz.FlavorGo if os.environ.get("MINLZ_FLAVOR") == "go" else mz.FlavorAMD64)
dev = torch.device("cuda:0")
src = synth.make_blocks(kind, nblk, bs, device=dev).reshape(-1)
soff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * bs
cap = bs + 16
eoff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * cap
enc = torch.empty(nblk * cap, dtype=torch.uint8, device=dev)
out_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
comp = torch.empty(nblk * bs, dtype=torch.uint8, device=dev)
coff = torch.zeros(nblk + 1, dtype=torch.int64, device=dev)
dec = torch.empty(nblk * bs, dtype=torch.uint8, device=dev)
status = torch.zeros(nblk, dtype=torch.int32, device=dev)
for _ in range(passes):
    mz.encode_blocks_dev(src, soff, enc, eoff, out_len, int(os.environ.get("MINLZ_LEVEL", "1")))
    mz.pack_blocks_dev(enc, eoff, out_len, comp, coff)
    mz.decode_blocks_dev(comp, coff, dec, soff, status)
torch.cuda.synchronize()
assert torch.equal(dec, src) and int(status.abs().sum()) == 0
print("ok", nblk, bs, kind, "ratio %.3f" % (nblk * bs / int(coff[-1])))
