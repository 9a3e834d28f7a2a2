"""A larger GPU fuzz than the suite's slices: random structures (tests/patterns.random_structure) plus
inputs built to stress the LevelBalanced probe-window walk's patching (runs, short periods, hash-group
collisions inside a window, matches that back-extend across a window base), every level, both encoder
flavours, encoder bytes against the oracle and decode of the result against the input.
usage: python profiles/fuzz_gpu.py [count] [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import minlz_b200 as mz
from oracle import binding as oracle
import patterns

count = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 20261017
rng = np.random.default_rng(seed)


def stress_inputs():
    out = []
    for n in (17, 40, 100, 1000, 5000, 70000, 300000):
        for per in (1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 33):
            base = rng.integers(0, 256, per, dtype=np.uint8)
            d = np.tile(base, n // per + 1)[:n].copy()
            k = int(rng.integers(0, 4))
            if k and n > 64:  # a few disturbances: runs of different lengths end inside / outside a window
                for p in rng.integers(20, n - 20, k * 3):
                    d[p] ^= 0x55
            out.append(d.tobytes())
    for n in (3000, 70000, 600000):
        # the same 7-gram at positions 1..30 apart: hash-group collisions inside one window
        d = rng.integers(0, 256, n, dtype=np.uint8)
        g = rng.integers(0, 256, 12, dtype=np.uint8)
        p = 50
        while p + 64 < n:
            gap = int(rng.integers(1, 31))
            d[p:p + 12] = g
            d[p + gap + 12:p + gap + 24] = g
            p += int(rng.integers(40, 400))
        out.append(d.tobytes())
        # literals then a match whose backward extension crosses many positions
        d = rng.integers(0, 256, n, dtype=np.uint8)
        p = 200
        while p + 200 < n:
            ln = int(rng.integers(20, 120))
            q = int(rng.integers(0, p - ln))
            d[p:p + ln] = d[q:q + ln]
            d[p + ln // 2] ^= 1  # the hash hit lands in the second half: extension runs backwards
            p += int(rng.integers(ln + 8, 600))
        out.append(d.tobytes())
    return out


items = stress_inputs()
while len(items) < count:
    items.append(patterns.random_structure(rng, max_n=700000).tobytes())
sizes = [len(b) for b in items]
print("inputs:", len(items), "bytes:", sum(sizes))
off = np.zeros(len(items) + 1, dtype=np.uint64)
np.cumsum(sizes, out=off[1:])
flat = np.frombuffer(b"".join(items), dtype=np.uint8)
bad = 0
for flavor, fname in ((mz.FlavorGo, "go"), (mz.FlavorAMD64, "asm")):
    mz.set_encoder_flavor(flavor)
    for level in (2, 1, -1):
        dst, doff, out_len = mz.encode_blocks(flat, off, level)
        streams, raws = [], []
        for i, d in enumerate(items):
            want = oracle.encode_block(d, level, flavor=fname)
            got = dst[int(doff[i]):int(doff[i]) + int(out_len[i])].tobytes()
            if got != want:
                bad += 1
                if bad < 10:
                    print("MISMATCH", fname, level, i, len(d), len(got), len(want))
            elif got:
                streams.append(got)
                raws.append(d)
        so = np.zeros(len(streams) + 1, dtype=np.uint64)
        np.cumsum([len(s) for s in streams], out=so[1:])
        do = np.zeros(len(raws) + 1, dtype=np.uint64)
        np.cumsum([len(r) for r in raws], out=do[1:])
        out, status = mz.decode_blocks(np.frombuffer(b"".join(streams), dtype=np.uint8), so, do)
        ok = (not status.any()) and out.tobytes() == b"".join(raws)
        print(fname, "level", level, "encoder mismatches so far:", bad, "decode ok:", ok)
        if not ok:
            bad += 1
mz.set_encoder_flavor(mz.FlavorGo)
print("FUZZ", "FAILED" if bad else "OK", "mismatches:", bad)
sys.exit(1 if bad else 0)
