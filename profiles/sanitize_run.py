"""Reduced parity subset for `compute-sanitizer` (memcheck / racecheck / synccheck) under gpurun.

The sanitizer slows kernels 10-100x, so this is NOT the parity suite: it is one pass
over every code path of the kernels at small sizes, each result still compared with
the oracle (a sanitizer-clean wrong answer is no use):

  * encoders: every size class boundary (both flavours, levels -1 / 1 / 2), compressible,
    barely compressible (bail-outs) and incompressible inputs;
  * decoder: valid streams of the above, a sample of the adversarial corpus, mutated
    streams (corrupt inputs must not write outside their range);
  * the sliced-upload path of the host encode call (>= 64 equal blocks of >= 256 KiB,
    arrival gate + CRC + pack), compressible and expanding;
  * stream encode / decode with CRC-32C, pack, validate mode.

usage: compute-sanitizer --tool memcheck python profiles/sanitize_run.py [quick]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import minlz_b200 as mz
from oracle import binding as oracle

import corpus
import patterns

quick = "quick" in sys.argv[1:]
rng = np.random.default_rng(7)


def cat(blobs):
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    np.cumsum([len(b) for b in blobs], out=off[1:])
    flat = np.frombuffer(b"".join(blobs), dtype=np.uint8) if off[-1] else np.zeros(0, dtype=np.uint8)
    return flat, off


def texty(n, seed):
    r = np.random.default_rng(seed)
    words = [bytes(r.integers(97, 123, int(r.integers(2, 9)), dtype=np.uint8)) for _ in range(300)]
    out = bytearray()
    while len(out) < n:
        out += words[int(r.integers(0, 300)) if r.random() < 0.8 else int(r.integers(0, 20))] + b" "
    return bytes(out[:n])


def inputs():
    sizes = [16, 17, 33, 100, 1024, 1025, 4096, 4097, 16384, 16385, 65536, 65537, 200000]
    if not quick:
        sizes += [512 << 10, (512 << 10) + 1, (2 << 20) + 1]
    out = []
    for i, n in enumerate(sizes):
        out.append(texty(n, i))
        noisy = bytearray(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
        for k in range(0, n - 40, 97):  # barely compressible: short repeats in noise
            noisy[k + 20:k + 28] = noisy[k:k + 8]
        out.append(bytes(noisy))
    out.append(rng.integers(0, 256, 70000, dtype=np.uint8).tobytes())
    out.append(bytes(300000))
    out += [d for _, d in patterns.reference_patterns()[:8]]
    return out


def check_encoders(raws):
    src, soff = cat(raws)
    for flavor, fname in ((mz.FlavorGo, "go"), (mz.FlavorAMD64, "asm")):
        mz.set_encoder_flavor(flavor)
        try:
            for level in (-1, 1, 2):
                dst, doff, out_len = mz.encode_blocks(src, soff, level)
                for i, d in enumerate(raws):
                    want = oracle.encode_block(d, level, flavor=fname)
                    got = dst[int(doff[i]):int(doff[i]) + int(out_len[i])].tobytes()
                    assert got == want, ("encode", fname, level, i, len(d))
        finally:
            mz.set_encoder_flavor(mz.FlavorGo)
    print("encoders ok:", len(raws), "inputs x 2 flavours x 3 levels")


def check_decoder(raws):
    streams, plain = [], []
    for level in (-1, 1, 2):
        for d in raws:
            t = oracle.encode_block(d, level)
            if t:
                streams.append(t)
                plain.append(d)
    src, soff = cat(streams)
    _, doff = cat(plain)
    dst, status = mz.decode_blocks(src, soff, doff)
    assert not status.any() and dst.tobytes() == b"".join(plain)
    # adversarial + mutated: accept / reject and bytes as the oracle, guard ranges intact
    blobs = [b for _, b in corpus.load_zip(corpus.golden_path("block-corpus-dec.zip"))][:60 if quick else 250]
    res = mz.DecodeBatch(blobs)
    for blob, r in zip(blobs, res):
        want = oracle.decode(blob)
        if isinstance(want, bytes):
            assert r == want
        else:
            assert isinstance(r, mz.MinLZError), (want, r)
    muts, lens = [], []
    base = [(s, len(p)) for s, p in zip(streams, plain) if 200 < len(p) <= 70000][:12]
    for s, n in base:
        for _ in range(6):
            t = bytearray(s)
            k = int(rng.integers(0, 3))
            if k == 0:
                t[int(rng.integers(0, len(t)))] = int(rng.integers(0, 256))
            elif k == 1:
                t = t[:int(rng.integers(1, len(t)))]
            else:
                p = int(rng.integers(0, len(t)))
                t[p:p] = bytes(rng.integers(0, 256, 3, dtype=np.uint8))
            muts.append(bytes(t))
            lens.append(n)
    # decode with exact dst lengths inside guarded ranges
    for i, (t, n) in enumerate(zip(muts, lens)):
        st, o = oracle.decode_block(t, n)
        o1, s1 = mz.decode_blocks(np.frombuffer(t, dtype=np.uint8), np.array([0, len(t)], dtype=np.uint64),
                                  np.array([0, n], dtype=np.uint64))
        assert int(s1[0]) == st, ("mutated status", i)
        if st == 0:
            assert o1.tobytes() == o
    print("decoder ok:", len(streams), "streams,", len(blobs), "adversarial,", len(muts), "mutated")


def check_sliced_and_stream():
    import synth
    from minlz_b200 import _lib
    lib = _lib.load()
    nblk, bs = 64, 256 << 10
    blocks = synth.make_blocks("json", nblk, bs, device="cpu").numpy()
    noise = rng.integers(0, 256, (nblk, bs), dtype=np.uint8)
    for name, data in (("json", blocks), ("noise", noise)):
        flat = np.ascontiguousarray(data).reshape(-1)
        soff = np.arange(nblk + 1, dtype=np.uint64) * bs
        for level in (1, -1, 2):
            cap = flat.size + 2 * nblk + 64
            dst = np.zeros(cap, dtype=np.uint8)
            poff = np.zeros(nblk + 1, dtype=np.uint64)
            crc = np.zeros(nblk, dtype=np.uint32)
            r = lib.mzcu_stream_encode_blocks(-1, level, nblk, flat.ctypes.data, soff.ctypes.data, dst.ctypes.data, dst.size,
                                              poff.ctypes.data, crc.ctypes.data)
            assert r == 0, lib.mzcu_last_error()
            for i in range(nblk):
                want = oracle.encode_block(data[i], level)
                got = dst[int(poff[i]):int(poff[i + 1])].tobytes()
                assert got == want, ("sliced", name, level, i)
                assert int(crc[i]) == oracle.crc(data[i].tobytes()), ("crc", name, level, i)
            if name == "json":
                out = np.zeros(flat.size, dtype=np.uint8)
                status = np.zeros(nblk, dtype=np.int32)
                dcrc = np.zeros(nblk, dtype=np.uint32)
                r = lib.mzcu_stream_decode_blocks(-1, nblk, dst.ctypes.data, poff.ctypes.data, out.ctypes.data,
                                                  soff.ctypes.data, status.ctypes.data, dcrc.ctypes.data)
                assert r == 0, lib.mzcu_last_error()
                assert not status.any() and np.array_equal(out, flat) and np.array_equal(dcrc, crc)
    print("sliced upload + stream crc ok")


if __name__ == "__main__":
    raws = inputs()
    check_encoders(raws)
    check_decoder(raws)
    check_sliced_and_stream()
    print("sanitize subset ok")
