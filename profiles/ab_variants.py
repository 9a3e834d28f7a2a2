"""A/B of experiment builds of the library in ONE process: every .so named on the command line is loaded
with ctypes, times encode_l1 (amd64 and Go flavour) on the bench batch with CUDA events, and prints a
checksum of everything it produced (the bench batch, plus a ragged batch whose blocks start at odd
addresses, at LevelFastest and LevelSuperFast) -- variants must agree with the first library named.
usage: python profiles/ab_variants.py lib0.so lib1.so ...   [env AB_BLOCKS=4096 AB_REPS=4]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import synth

nblk = int(os.environ.get("AB_BLOCKS", "4096"))
reps = int(os.environ.get("AB_REPS", "4"))
l2_reps = int(os.environ.get("AB_L2_REPS", "0"))  # > 0: time LevelBalanced too
bs = 1 << 20
dev = torch.device("cuda:0")
P = C.c_void_p


def checksum(enc, cap, out_len):
    tot = 0
    w = torch.arange(1, cap + 1, dtype=torch.int64, device=dev)
    rows = enc.view(-1, cap)
    for i in range(0, rows.shape[0], 128):
        r = rows[i:i + 128].to(torch.int64)
        m = torch.arange(cap, device=dev)[None, :] < out_len[i:i + 128, None].to(torch.int64)
        tot = (tot * 1000003 + int((r * w * m).sum().item())) % (1 << 61)
    return "%d/%016x" % (int(out_len.to(torch.int64).sum().item()), tot)


src = synth.make_blocks("json", nblk, bs, device=dev).reshape(-1)
soff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * bs
cap = bs + 16
eoff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * cap
enc = torch.empty(nblk * cap, dtype=torch.uint8, device=dev)
out_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
# ragged batch: 150 blocks of 700 001 .. 700 150 bytes packed back to back from byte 5 (every alignment)
rn = 150
rl = torch.arange(rn, dtype=torch.int64) + 700001
roff = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(rl, 0)]) + 5
r_src = torch.cat([torch.zeros(5, dtype=torch.uint8, device=dev),
                   synth.make_blocks("log", 128, 1 << 20, device=dev, seed=7).reshape(-1)[: int(roff[-1]) - 5]])
r_soff = roff.to(dev)
rcap = 700150 + 64
r_eoff = torch.arange(rn + 1, dtype=torch.int64, device=dev) * rcap
r_enc = torch.empty(rn * rcap, dtype=torch.uint8, device=dev)
r_len = torch.zeros(rn, dtype=torch.int32, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
first = None
for path in sys.argv[1:]:
    lib = C.CDLL(os.path.abspath(path))
    lib.mzcu_encode_blocks_dev.restype = C.c_int
    lib.mzcu_encode_blocks_dev.argtypes = [C.c_int, C.c_int, C.c_int, P, P, P, P, P, P]
    lib.mzcu_last_error.restype = C.c_char_p
    sp = P(torch.cuda.current_stream().cuda_stream)
    sums = []
    line = []
    for fl, name in ((1, "amd64"), (0, "go")):
        lib.mzcu_set_encoder_flavor(fl)
        ts = []
        for r in range(reps + 2):
            out_len.zero_()
            ev0.record()
            rc = lib.mzcu_encode_blocks_dev(0, 1, nblk, src.data_ptr(), soff.data_ptr(), enc.data_ptr(), eoff.data_ptr(),
                                            out_len.data_ptr(), sp)
            ev1.record()
            torch.cuda.synchronize()
            assert rc == 0, lib.mzcu_last_error()
            if r >= 2:
                ts.append(ev0.elapsed_time(ev1))
        ts.sort()
        line.append("%s median %.2f min %.2f ms" % (name, ts[len(ts) // 2], ts[0]))
        sums.append(checksum(enc, cap, out_len))
        if l2_reps:
            ts = []
            for r in range(l2_reps + 1):
                out_len.zero_()
                ev0.record()
                rc = lib.mzcu_encode_blocks_dev(0, 2, nblk, src.data_ptr(), soff.data_ptr(), enc.data_ptr(), eoff.data_ptr(),
                                                out_len.data_ptr(), sp)
                ev1.record()
                torch.cuda.synchronize()
                assert rc == 0, lib.mzcu_last_error()
                if r >= 1:
                    ts.append(ev0.elapsed_time(ev1))
            ts.sort()
            line.append("L2 %s median %.2f ms" % (name, ts[len(ts) // 2]))
            sums.append(checksum(enc, cap, out_len))
        for level in (1, -1) + ((2,) if l2_reps else ()):
            r_len.zero_()
            rc = lib.mzcu_encode_blocks_dev(0, level, rn, r_src.data_ptr(), r_soff.data_ptr(), r_enc.data_ptr(), r_eoff.data_ptr(),
                                            r_len.data_ptr(), sp)
            torch.cuda.synchronize()
            assert rc == 0, lib.mzcu_last_error()
            sums.append(checksum(r_enc, rcap, r_len))
    if first is None:
        first = sums
    print("%-34s %s | outputs %s" % (os.path.basename(path), "; ".join(line), "SAME" if sums == first else "DIFFERENT " + str(sums)),
          flush=True)
    if first is sums:
        print("   reference checksums:", sums, flush=True)
