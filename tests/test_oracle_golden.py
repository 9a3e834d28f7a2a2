"""Pins the CPU oracle against the reference's golden vectors and KATs.

Sources (reference tree):
  minlz_test.go:632-660   TestDecodeGoldenInput  (.mzb -> .txt)
  minlz_test.go:871-911   TestEmitLiteral
  minlz_test.go:913-1026  TestEmitCopy
  minlz_test.go:42-69     TestMaxEncodedLen
  minlz_test.go:1092-1094 masked CRC32C("abcd")
  encode_test.go:102-534  TestEmitters (emit -> decode field round trip, subsampled)
"""
import ctypes as C
import random

import numpy as np
import pytest

import corpus
import spec_decoder


def test_decode_golden(oracle):
    want = open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt"), "rb").read()
    mzb = open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt.mzb"), "rb").read()
    assert oracle.decoded_len(mzb) == len(want)
    assert oracle.decode(mzb) == want
    assert spec_decoder.decode_block(mzb) == want


EMIT_LITERAL = [(1, b"\x00"), (2, b"\b"), (27, b"\xd0"), (28, b"\xd8"), (29, b"\xe0"), (30, b"\xe8\x00"),
                (59, b"\xe8\x1d"), (60, b"\xe8\x1e"), (61, b"\xe8\x1f"), (62, b"\xe8 "), (254, b"\xe8\xe0"),
                (255, b"\xe8\xe1"), (256, b"\xe8\xe2"), (257, b"\xe8\xe3"), (65534, b"\xf0\xe0\xff"),
                (65535, b"\xf0\xe1\xff"), (65536, b"\xf0\xe2\xff"), (165536, b"\xf8\x82\x86\x02")]


def test_emit_literal(oracle):
    L = oracle.lib()
    nines = np.full(1 << 18, 0x99, dtype=np.uint8)
    dst = np.zeros((1 << 18) + 8, dtype=np.uint8)
    for length, want in EMIT_LITERAL:
        n = L.mzo_emit_literal(dst.ctypes.data, nines.ctypes.data, length)
        got = dst[:n].tobytes()
        assert got[-length:] == nines[:length].tobytes()
        assert got[:n - length] == want, length


EMIT_COPY = [
    (8, 4, [0xc1, 0x1]), (8, 11, [0xdd, 0x1]), (8, 12, [0xe1, 0x1]), (8, 13, [0xe5, 0x1]), (8, 17, [0xf5, 0x1]),
    (8, 18, [0xf9, 0x1]), (8, 19, [0xfd, 0x1, 0x1]), (8, 59, [0xfd, 0x1, 0x29]), (8, 60, [0xfd, 0x1, 0x2a]),
    (8, 61, [0xfd, 0x1, 0x2b]), (8, 62, [0xfd, 0x1, 0x2c]), (8, 63, [0xfd, 0x1, 0x2d]), (8, 64, [0xfd, 0x1, 0x2e]),
    (8, 65, [0xfd, 0x1, 0x2f]), (8, 66, [0xfd, 0x1, 0x30]), (8, 67, [0xfd, 0x1, 0x31]), (8, 68, [0xfd, 0x1, 0x32]),
    (8, 69, [0xfd, 0x1, 0x33]), (8, 80, [0xfd, 0x1, 0x3e]), (8, 800, [0xf9, 0x1, 0xf4, 0xf0, 0x2]),
    (8, 800000, [0xf9, 0x1, 0xfc, 0xd0, 0x34, 0xc]),
    (256, 4, [0xc1, 0x3f]), (256, 11, [0xdd, 0x3f]), (256, 12, [0xe1, 0x3f]), (256, 13, [0xe5, 0x3f]),
    (256, 18, [0xf9, 0x3f]), (256, 19, [0xfd, 0x3f, 0x1]), (256, 59, [0xfd, 0x3f, 0x29]),
    (256, 60, [0xfd, 0x3f, 0x2a]), (256, 61, [0xfd, 0x3f, 0x2b]), (256, 62, [0xfd, 0x3f, 0x2c]),
    (256, 63, [0xfd, 0x3f, 0x2d]), (256, 64, [0xfd, 0x3f, 0x2e]), (256, 65, [0xfd, 0x3f, 0x2f]),
    (256, 66, [0xfd, 0x3f, 0x30]), (256, 67, [0xfd, 0x3f, 0x31]), (256, 68, [0xfd, 0x3f, 0x32]),
    (256, 69, [0xfd, 0x3f, 0x33]), (256, 80, [0xfd, 0x3f, 0x3e]), (256, 800, [0xf9, 0x3f, 0xf4, 0xf0, 0x2]),
    (256, 80000, [0xf9, 0x3f, 0xfc, 0x50, 0x38, 0x1]),
    (2048, 4, [0x2, 0xc0, 0x7]), (2048, 11, [0x1e, 0xc0, 0x7]), (2048, 12, [0x22, 0xc0, 0x7]),
    (2048, 13, [0x26, 0xc0, 0x7]), (2048, 59, [0xde, 0xc0, 0x7]), (2048, 60, [0xe2, 0xc0, 0x7]),
    (2048, 61, [0xe6, 0xc0, 0x7]), (2048, 62, [0xea, 0xc0, 0x7]), (2048, 63, [0xee, 0xc0, 0x7]),
    (2048, 64, [0xf2, 0xc0, 0x7]), (2048, 65, [0xf6, 0xc0, 0x7, 0x1]), (2048, 66, [0xf6, 0xc0, 0x7, 0x2]),
    (2048, 67, [0xf6, 0xc0, 0x7, 0x3]), (2048, 68, [0xf6, 0xc0, 0x7, 0x4]), (2048, 69, [0xf6, 0xc0, 0x7, 0x5]),
    (2048, 80, [0xf6, 0xc0, 0x7, 0x10]), (2048, 800, [0xfa, 0xc0, 0x7, 0xe0, 0x2]),
    (2048, 80000, [0xfe, 0xc0, 0x7, 0x40, 0x38, 0x1]),
    (204800, 4, [0x7, 0x0, 0x0, 0x11]), (204800, 28, [0x7, 0x3, 0x0, 0x11]), (204800, 32, [0x87, 0x3, 0x0, 0x11]),
    (204800, 33, [0xa7, 0x3, 0x0, 0x11]), (204800, 40, [0x87, 0x4, 0x0, 0x11]),
    (204800, 65, [0xa7, 0x7, 0x0, 0x11, 0x1]), (204800, 69, [0xa7, 0x7, 0x0, 0x11, 0x5]),
    (204800, 800, [0xc7, 0x7, 0x0, 0x11, 0xe0, 0x2]), (204800, 80000, [0xe7, 0x7, 0x0, 0x11, 0x40, 0x38, 0x1]),
]


def test_emit_copy(oracle):
    L = oracle.lib()
    dst = np.zeros(1024, dtype=np.uint8)
    assert len(EMIT_COPY) == 68
    for off, length, want in EMIT_COPY:
        n = L.mzo_emit_copy(dst.ctypes.data, off, length)
        assert dst[:n].tolist() == want, (off, length)


def test_max_encoded_len(oracle):
    # minlz_test.go:42-69 / encode.go:234-244
    assert oracle.max_encoded_len(0) == 1
    for n in (1, 15, 16, 100, 65536, 1 << 20, 8 << 20):
        assert oracle.max_encoded_len(n) == n + 2
    assert oracle.max_encoded_len((8 << 20) + 1) == -1


def test_crc_kat(oracle):
    assert oracle.crc(b"abcd").to_bytes(4, "little") == bytes([0x68, 0x10, 0xe6, 0xb6])


def _decode_tokens(oracle, prefix_len, tokens, total):
    """Decode `tokens` after a literal run of prefix_len bytes; returns output."""
    L = oracle.lib()
    pre = np.arange(prefix_len, dtype=np.uint32).astype(np.uint8)
    hdr = np.zeros(8, dtype=np.uint8)
    n = L.mzo_emit_literal(hdr.ctypes.data, pre.ctypes.data, 0)  # header only when len 0
    buf = np.zeros(prefix_len + 8, dtype=np.uint8)
    n = L.mzo_emit_literal(buf.ctypes.data, pre.ctypes.data, prefix_len)
    stream = buf[:n].tobytes() + tokens
    return oracle.decode_block(stream, total)


def test_emitters_roundtrip(oracle):
    """encode_test.go:102-534 TestEmitters, subsampled: every emitter's output
    decodes to the (offset, length, literals) it was given."""
    L = oracle.lib()
    rng = random.Random(1)
    dst = np.zeros(64, dtype=np.uint8)
    lits = np.array([0xa1, 0xa2, 0xa3, 0xa4], dtype=np.uint8)
    cases = []
    for off in list(range(1, 70)) + [1023, 1024, 1025, 65535, 65536, 65599, 65600, 100000, 2162687]:
        for length in list(range(4, 80)) + [273, 274, 275, 1000, 65599, 65600, 70000]:
            cases.append((off, length))
    rng.shuffle(cases)
    for off, length in cases[:1500]:
        prefix = off + rng.randrange(0, 8)
        pre = (np.arange(prefix, dtype=np.uint32) * 7 + 3).astype(np.uint8)
        want = bytearray(pre.tobytes())
        n = L.mzo_emit_copy(dst.ctypes.data, off, length)
        tok = dst[:n].tobytes()
        nl = 0
        if off >= 64 and rng.random() < 0.5:
            nl = rng.randrange(1, 5 if off <= 65599 else 4)
            if off <= 65599:
                n = L.mzo_emit_copy_lits2(dst.ctypes.data, lits.ctypes.data, nl, off, length)
            else:
                n = L.mzo_emit_copy_lits3(dst.ctypes.data, lits.ctypes.data, nl, off, length)
            tok = dst[:n].tobytes()
            # offset is relative to the position after the fused literals
            want += lits[:nl].tobytes()
        if off > len(want):
            continue
        for i in range(length):
            want.append(want[len(want) - off])
        buf = np.zeros(prefix + 8, dtype=np.uint8)
        m = L.mzo_emit_literal(buf.ctypes.data, pre.ctypes.data, prefix)
        st, out = oracle.decode_block(buf[:m].tobytes() + tok, len(want))
        assert st == 0, (off, length, nl)
        assert out == bytes(want), (off, length, nl)
        full = b"\x00" + _uvarint(len(want)) + buf[:m].tobytes() + tok
        if len(full) - 1 - len(_uvarint(len(want))) <= len(want):
            assert spec_decoder.decode_block(full) == bytes(want)


def _uvarint(x):
    out = bytearray()
    while x >= 0x80:
        out.append((x & 0x7f) | 0x80)
        x >>= 7
    out.append(x)
    return bytes(out)


def test_repeat_lengths(oracle):
    L = oracle.lib()
    dst = np.zeros(16, dtype=np.uint8)
    for length in [1, 2, 29, 30, 31, 285, 286, 65565, 65566, 100000]:
        n = L.mzo_emit_repeat(dst.ctypes.data, length)
        pre = b"\x00z"  # literal 'z', then repeat with offset 1
        st, out = oracle.decode_block(pre + dst[:n].tobytes(), 1 + length)
        assert st == 0 and out == b"z" * (1 + length)


def test_asm_flavour_matches_recorded_reference_output(oracle):
    """tests/golden/asm_digests.json holds digests of what the reference's real amd64 assembly
    emitted (recorded through oracle/_ref by tests/golden/make_asm_golden.py).  The restated
    amd64 flavour must reproduce every one of them -- this pins it to reference output even
    where oracle/_ref cannot be built."""
    import hashlib
    import json
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_asm_golden
    rec = json.load(open(corpus.golden_path("asm_digests.json")))["digests"]
    items = make_asm_golden.inputs()
    assert len(items) == len(rec) > 150
    for name, data in items:
        assert rec[name]["n"] == len(data)
        for level in (-1, 1, 2):
            got = hashlib.sha256(oracle.encode_block(data, level, flavor="asm")).hexdigest()
            assert got == rec[name][str(level)], (name, level)
