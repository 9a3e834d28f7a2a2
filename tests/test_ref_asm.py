"""Pins the oracle to the reference ITSELF: the reference's own AMD64 assembly
(asm_amd64.s), transliterated to GNU as and run here through oracle/_ref
(oracle/p9_to_gas.py, oracle/ref_shim.c, oracle/refasm.py).

What is pinned against real reference code, not against a restatement:
  * decoder: status + bytes of mzo_decode_block == decodeBlockAsm on the golden
    block, on every encoder output below and on the adversarial corpora
    (the reference's own TestCompareDecoders idea, decode_asm_test.go);
  * emitters and matchLen: == the assembly's emitLiteral/emitRepeat/emitCopy/
    emitCopyLits2/emitCopyLits3/matchLen on value grids;
  * all three encoders, amd64 flavour (mzo_encode_block_l{0,1,2}_asm):
    BYTE-IDENTICAL to encodeFastBlockAsm* / encodeBlockAsm* / encodeBetterBlockAsm*
    over every size class, corpus and bail-out case;
  * the Go-flavour encoders (the noasm build, what the default GPU mode mirrors):
    their streams decode to the input with the real decoder, and for blocks
    > 512 KiB the L1 Go flavour differs from the real assembly only in the tail
    (same table, hash, skip: SURVEY 8a) -- checked as a common-prefix bound.

Skipped when neither oracle/_ref/libminlz_ref.so nor /root/reference exists.
"""
import numpy as np
import pytest

import corpus
import patterns
import synth
from oracle import refasm

pytestmark = pytest.mark.skipif(not refasm.available(), reason="oracle/_ref not built and /root/reference absent")

SIZES = (17, 18, 31, 32, 33, 40, 64, 100, 500, 1024, 1025, 2000, 4096, 4097, 10000, 16384, 16385, 40000,
         65536, 65537, 100000, 300000, 524288, 524289, 1 << 20, 2 << 20, (2 << 20) + 1, 3 << 20)


def _split_block(blob):
    """0x00 + uvarint(size) + tokens -> (size, tokens) or None (not a compressed MinLZ block)."""
    if len(blob) < 2 or blob[0] != 0:
        return None
    v, sh, i = 0, 0, 1
    while True:
        if i >= len(blob) or sh > 63:
            return None
        b = blob[i]
        i += 1
        v |= (b & 0x7F) << sh
        sh += 7
        if b < 0x80:
            break
    if v == 0 or v > (8 << 20) or len(blob) - i > v:
        return None
    return v, blob[i:]


def _same_encoders(oracle, data, tag):
    for level in (-1, 1, 2):
        want = refasm.encode_block(data, level)
        got = oracle.encode_block(data, level, flavor="asm")
        assert got == want, "level %d %s: restated %d B vs assembly %d B" % (level, tag, len(got), len(want))
        if want:
            st, out = refasm.decode_block(want, len(data))
            assert st == 0 and out == bytes(data)
            st, out = oracle.decode_block(want, len(data))
            assert st == 0 and out == bytes(data)


def test_decoder_golden(oracle):
    want = open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt"), "rb").read()
    size, body = _split_block(open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt.mzb"), "rb").read())
    st, out = refasm.decode_block(body, size)
    assert st == 0 and out == want
    assert oracle.decode_block(body, size) == (0, want)


def test_decoder_adversarial_agrees(oracle):
    n = bad = 0
    for name in ("dec-block-regressions.zip", "block-corpus-dec.zip"):
        for tag, blob in corpus.load_zip(corpus.golden_path(name)):
            # as a block (header parsed) and, like the fuzzers, as a bare token stream with guessed sizes
            cases = [(blob, len(blob) * 2 + 8), (blob[1:], 1000)] if len(blob) < 50000 else []
            sp = _split_block(blob)
            if sp is not None:
                cases.append((sp[1], sp[0]))
            for body, size in cases:
                st_r, out_r = refasm.decode_block(body, size)
                st_o, out_o = oracle.decode_block(body, size)
                assert (st_r != 0) == (st_o != 0), "%s: assembly %d oracle %d" % (tag, st_r, st_o)
                if st_r == 0:
                    assert out_r == out_o, tag
                else:
                    bad += 1
                n += 1
    assert n > 1000 and bad > 500


def test_decoder_mutations_agree(oracle):
    """Valid streams with random byte flips / truncations: same accept/reject and same bytes."""
    rng = np.random.default_rng(23)
    n = ok = 0
    for kind, size, level in (("json", 40000, 1), ("text", 70000, 2), ("binary", 20000, 1), ("log", 300000, 2)):
        blk = synth.make_blocks(kind, 1, size).numpy()[0]
        enc = np.frombuffer(oracle.encode_block(blk, level), dtype=np.uint8)
        for k in range(150):
            m = enc.copy()
            if k % 3 == 2:
                m = m[:int(rng.integers(1, m.size))]
            else:
                for p in rng.integers(0, m.size, 1 + k % 2):
                    m[p] = rng.integers(0, 256)
            st_r, out_r = refasm.decode_block(m, size)
            st_o, out_o = oracle.decode_block(m, size)
            assert (st_r != 0) == (st_o != 0), (kind, k, st_r, st_o)
            if st_r == 0:
                assert out_r == out_o
                ok += 1
            n += 1
    assert n == 600 and 20 < ok < 580


def test_emitters_match_assembly(oracle):
    L, R = oracle.lib(), refasm.lib()
    a = np.zeros(1 << 17, dtype=np.uint8)
    b = np.zeros(1 << 17, dtype=np.uint8)
    lit = (np.arange(1 << 17) * 7 % 251).astype(np.uint8)
    for n in list(range(1, 300)) + [65535 + 29, 65536 + 29, 65536 + 30, 65536 + 31, 70000, 100000]:
        a[:] = 0
        b[:] = 0
        x = L.mzo_emit_literal(a.ctypes.data, lit.ctypes.data, n)
        y = R.mzr_emit_literal(b.ctypes.data, b.size, lit.ctypes.data, n)
        assert x == y and bytes(a[:x]) == bytes(b[:y]), n
    lens = list(range(1, 320)) + [65535 + 29, 65536 + 29, 65536 + 30, 65536 + 31, 1 << 20, (8 << 20) - 1]
    for n in lens:
        x = L.mzo_emit_repeat(a.ctypes.data, n)
        y = R.mzr_emit_repeat(b.ctypes.data, b.size, n)
        assert x == y and bytes(a[:x]) == bytes(b[:y]), n
    offs = [1, 2, 63, 64, 65, 1023, 1024, 1025, 65535, 65536, 65598, 65599, 65600, 70000, 1 << 20, 2162687]
    clens = [4, 5, 11, 12, 18, 19, 63, 64, 65, 273, 274, 275, 300, 319, 320, 321, 65599, 65600, 65601, 70000, 1 << 20]
    for off in offs:
        for n in clens:
            x = L.mzo_emit_copy(a.ctypes.data, off, n)
            y = R.mzr_emit_copy(b.ctypes.data, b.size, off, n)
            assert x == y and bytes(a[:x]) == bytes(b[:y]), (off, n)
            for nl in (1, 2, 3, 4):
                if 64 <= off <= 65599:
                    x = L.mzo_emit_copy_lits2(a.ctypes.data, lit.ctypes.data, nl, off, n)
                    y = R.mzr_emit_copy_lits2(b.ctypes.data, b.size, lit.ctypes.data, nl, off, n)
                    assert x == y and bytes(a[:x]) == bytes(b[:y]), (off, n, nl)
                if off >= 65536 and nl <= 3:
                    x = L.mzo_emit_copy_lits3(a.ctypes.data, lit.ctypes.data, nl, off, n)
                    y = R.mzr_emit_copy_lits3(b.ctypes.data, b.size, lit.ctypes.data, nl, off, n)
                    assert x == y and bytes(a[:x]) == bytes(b[:y]), (off, n, nl)


def test_match_len_assembly():
    R = refasm.lib()
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, 5000, dtype=np.uint8)
    for n in list(range(0, 80)) + [127, 128, 129, 1000, 4999]:
        for k in sorted({0, 1, 7, 8, 9, 15, 16, 17, 31, 33, n - 1, n}):
            if k < 0 or k > n:
                continue
            x = base[:n].copy()
            y = base[:n].copy()
            if k < n:
                y[k] ^= 0x40
            assert R.mzr_match_len(x.ctypes.data, n, y.ctypes.data, n) == k, (n, k)


def test_encoders_identical_synthetic(oracle):
    for kind in ("json", "log", "text", "binary", "random"):
        big = synth.make_blocks(kind, 1, 3 << 20).numpy()[0]
        for n in SIZES:
            _same_encoders(oracle, big[:n].copy(), (kind, n))


def test_encoders_identical_8mb(oracle):
    # the 8 MiB variant (encodeBlockAsm / encodeFastBlockAsm): far-candidate clamp, copy3 offsets
    for kind in ("text", "binary"):
        blk = synth.make_blocks(kind, 1, 8 << 20).numpy()[0]
        _same_encoders(oracle, blk, (kind, "8MiB"))
    far = np.frombuffer(patterns.large_offset(8 << 20, (2 << 20) + 70000), dtype=np.uint8)
    _same_encoders(oracle, far, "large_offset")
    _same_encoders(oracle, np.zeros(8 << 20, dtype=np.uint8), "zeros")


def test_encoders_identical_corpora(oracle):
    n = 0
    for name in ("enc_regressions.zip", "block-corpus-raw-sample.zip", "block-corpus-enc-sample.zip"):
        for tag, data in corpus.load_zip(corpus.golden_path(name)):
            if len(data) < 17:
                continue
            _same_encoders(oracle, np.frombuffer(data, dtype=np.uint8), tag)
            n += 1
    assert n > 100
    tw = np.frombuffer(open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt"), "rb").read(), dtype=np.uint8)
    _same_encoders(oracle, tw, "twain")
    assert len(refasm.encode_block(tw, 1)) == 10288    # sizes of the real assembly, recorded 2026-10
    assert len(refasm.encode_block(tw, 2)) == 9475


def test_encoders_identical_patterns(oracle):
    gens = [patterns.fused_lits(70000), patterns.long_literals(70000), patterns.offset2(70000),
            patterns.pattern_35_7a(100000), patterns.short_repeat(3, 400), patterns.large_offset(300000, 70000),
            b"abcdefgh" * 100, b"a" * 17, b"ab" * 3000, bytes(range(256)) * 300]
    for i, g in enumerate(gens):
        _same_encoders(oracle, np.frombuffer(g, dtype=np.uint8), "pattern%d" % i)


def test_encoders_identical_bailouts(oracle):
    """Barely compressible inputs: random bytes with repeats sprinkled at a density that
    puts the output next to dstLimit, so every bail test (gen.go:395-417) fires somewhere;
    plus literal runs of 286+ bytes in the 16K-64K class (the 3-byte-length quirk)."""
    rng = np.random.default_rng(11)
    hits = [0, 0, 0, 0]   # [cases, level 1, level 2, level -1]
    for n in (600, 3000, 12000, 50000, 200000, 600000, 1 << 20):
        for dens in (0.0, 0.05, 0.08, 0.1, 0.12, 0.15, 0.2, 0.3, 0.4, 0.5, 0.6, 0.8):
            d = rng.integers(0, 256, n, dtype=np.uint8)
            k = int(n * dens / 48)
            for p in rng.integers(64, n - 64, k):
                q = int(rng.integers(0, p - 48))
                d[p:p + 48] = d[q:q + 48]
            _same_encoders(oracle, d, ("bail", n, dens))
            for level in (-1, 1, 2):
                hits[level] += refasm.encode_block(d, level) == b""
            hits[0] += 1
    assert all(5 < hits[lv] < hits[0] - 5 for lv in (-1, 1, 2)), hits


def test_l2_assembly_round_trips(oracle):
    # LevelBalanced assembly streams decode with both decoders
    for kind, n in (("json", 1 << 20), ("text", 300000), ("binary", 40000), ("log", 3 << 20)):
        blk = synth.make_blocks(kind, 1, n).numpy()[0]
        enc = refasm.encode_block(blk, 2)
        assert enc
        assert oracle.decode_block(enc, n) == (0, blk.tobytes())
        assert refasm.decode_block(enc, n) == (0, blk.tobytes())


def test_go_flavour_against_real_decoder_and_prefix(oracle):
    for kind, n in (("json", 1 << 20), ("log", 2 << 20), ("text", 3 << 20), ("json", 100000), ("binary", 5000)):
        blk = synth.make_blocks(kind, 1, n).numpy()[0]
        for level in (-1, 1, 2):
            enc = oracle.encode_block(blk, level)
            assert enc
            assert refasm.decode_block(enc, n) == (0, blk.tobytes())
        if n > (512 << 10):
            a, g = refasm.encode_block(blk, 1), oracle.encode_block(blk, 1)
            same = next((i for i in range(min(len(a), len(g))) if a[i] != g[i]), min(len(a), len(g)))
            assert same >= len(a) - 64, "Go and asm L1 flavours should only differ in the block tail"


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/testdata/fuzz"), reason="reference checkout not present")
def test_encoders_identical_full_fuzz_corpora(oracle):
    """The reference's complete block corpora (fuzz_test.go:31 FuzzEncodingBlocks seeds), not the
    committed samples: every input, all three levels, byte-identical to the real assembly."""
    n = 0
    for name in ("block-corpus-raw.zip", "block-corpus-enc.zip"):
        for tag, data in corpus.load_zip("/root/reference/testdata/fuzz/" + name):
            if len(data) < 17 or len(data) > (1 << 20):
                continue
            _same_encoders(oracle, np.frombuffer(data, dtype=np.uint8), tag)
            n += 1
    assert n > 900


def test_encoders_identical_random_structures(oracle):
    """The same generator as the GPU fuzz batch (patterns.random_structure), on the CPU: restated
    amd64 flavour == real assembly, all levels; the streams decode with both decoders."""
    rng = np.random.default_rng(77)
    for i in range(250):
        _same_encoders(oracle, patterns.random_structure(rng, max_n=800000), ("fuzz", i))
