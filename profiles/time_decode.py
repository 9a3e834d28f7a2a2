"""Times only the decode kernel (MINLZ_NO_CHECK=1: experiment builds whose output is not meant to be right)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minlz_b200 as mz
import synth
nblk, bs = 4096, 1 << 20
dev = torch.device("cuda:0")
src = synth.make_blocks("json", nblk, bs, device=dev).reshape(-1)
soff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * bs
cap = bs + 16
eoff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * cap
enc = torch.empty(nblk * cap, dtype=torch.uint8, device=dev)
out_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
comp = torch.empty(nblk * bs, dtype=torch.uint8, device=dev)
coff = torch.zeros(nblk + 1, dtype=torch.int64, device=dev)
dec = torch.empty(nblk * bs, dtype=torch.uint8, device=dev)
status = torch.zeros(nblk, dtype=torch.int32, device=dev)
mz.encode_blocks_dev(src, soff, enc, eoff, out_len, 1)
mz.pack_blocks_dev(enc, eoff, out_len, comp, coff)
for _ in range(3):
    mz.decode_blocks_dev(comp, coff, dec, soff, status)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    mz.decode_blocks_dev(comp, coff, dec, soff, status)
e1.record()
torch.cuda.synchronize()
print("decode ms", e0.elapsed_time(e1) / 5)
if not os.environ.get("MINLZ_NO_CHECK"):
    assert int(status.abs().sum()) == 0 and torch.equal(dec, src), "decode mismatch"
    print("decode output verified")
