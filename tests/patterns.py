"""Deterministic input generators restated from the reference's tests:
decode_asm_test.go:188-346 (tag-sequence / overlap patterns), :352 (margin
sizes), minlz_test.go:202-253 (TestSmallCopy/Rand/Regular/Repeat) and
:780 TestEncodeNoiseThenRepeats."""
import numpy as np


def _fill(size, fn):
    i = np.arange(size, dtype=np.int64)
    return bytearray((fn(i) % 256).astype(np.uint8).tobytes())


def large_offset(size, min_offset):
    d = _fill(size, lambda i: i % 251)
    pat = b"LARGEPAT"
    if min_offset < size - len(pat) * 2:
        d[0:len(pat)] = pat
        d[min_offset:min_offset + len(pat)] = pat
    return bytes(d)


def fused_lits(size):
    d = _fill(size, lambda i: i * 3)
    for i in range(100, size - 10, 500):
        d[i:i + 4] = b"ABCD"
        d[i + 4] = i % 256
        d[i + 5] = (i + 1) % 256
    return bytes(d)


def long_literals(size):
    return bytes(_fill(size, lambda i: i * 17 + i * i))


def offset2(size):
    d = bytearray(b"\x35\x7a" * (size // 2 + 1))[:size]
    for i in range(1000, size - 100, 3000):
        d[i:i + 13] = b"UNIQUE_MARKER"
    return bytes(d)


def short_repeat(offset, length):
    d = _fill(10000, lambda i: i * 7)
    pat = bytes(ord("A") + i for i in range(offset))
    d[1000:1000 + offset] = pat
    for i in range(length):
        d[1000 + offset + i] = pat[i % offset]
    return bytes(d)


def pattern_35_7a(size):
    i = np.arange(size)
    m = i % 10
    d = (ord("0") + m).astype(np.uint8)
    d[m == 0] = ord("3")
    d[m == 1] = ord("5")
    d[(m == 2) | (m == 3)] = ord("z")
    d[m == 4] = ord("1")
    d = bytearray(d.tobytes())
    for k in range(0, size - 20, 3000):
        d[k:k + 14] = b"UNIQUE_MARKER_"
    return bytes(d)


def very_large_offset2(pos):
    d = _fill(pos + 1000, lambda i: i * 13 + i // 7)
    d[pos - 100:pos - 94] = b"\x35\x7a\x35\x7a\x35\x7a"
    return bytes(d)


def generate_test_data(size):
    d = _fill(size, lambda i: i * 7 + i // 13)
    pat = b"PATTERN_DATA_HERE"
    for i in range(1000, size - len(pat), 5000):
        d[i:i + len(pat)] = pat
    return bytes(d)


def margin_patterns(size):
    return [b"a" * size, (b"ab" * (size // 2 + 1))[:size], (b"abcd" * (size // 4 + 1))[:size], long_literals(size)]


def reference_patterns():
    """(name, bytes) for every generator case the reference's decoder tests use."""
    out = [("large_offset_copy2", large_offset(100000, 65000)), ("large_offset_copy3", large_offset(200000, 100000)),
           ("fused_lits_small", fused_lits(10000)), ("fused_lits_large", fused_lits(100000)),
           ("long_literals", long_literals(100000))]
    out += [("offset2_%d" % s, offset2(s)) for s in (65549, 70000, 60000, 50000)]
    out += [("short_repeat_%d_%d" % (o, l), short_repeat(o, l)) for o, l in ((1, 4), (2, 4), (2, 10), (3, 9), (4, 16))]
    out += [("35_7a_%d" % s, pattern_35_7a(s)) for s in (65549, 66000, 60000)]
    out += [("vlo2_%d" % p, very_large_offset2(p)) for p in (50000, 55000, 56000, 56500, 56600, 56620, 57000)]
    out += [("testdata_%d" % s, generate_test_data(s)) for s in (100, 1000, 10000, 100000, 1 << 20)]
    for size in (30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 50, 60, 70, 80):
        out += [("margin_%d_%d" % (size, k), p) for k, p in enumerate(margin_patterns(size))]
    return out


def roundtrip_inputs():
    """minlz_test.go:202-253 + :780, sizes trimmed to keep the suite fast."""
    out = []
    for n in (16, 17, 20, 32, 100, 1000, 4096):           # TestSmallCopy-like
        out.append(("smallcopy_%d" % n, (b"aaaa" + bytes(range(n % 200)) + b"aaaa" * 2 + bytes(n))[:max(n, 16)]))
    rng = np.random.default_rng(1)
    for n in (16, 100, 1000, 65535, 65536, 65537, 200000):  # TestSmallRand
        out.append(("rand_%d" % n, rng.integers(0, 256, n, dtype=np.uint8).tobytes()))
    for n in (16, 100, 1000, 65535, 65536, 65537, 200000):  # TestSmallRegular
        out.append(("regular_%d" % n, bytes((np.arange(n) % 10 + ord("a")).astype(np.uint8))))
    for n in (16, 100, 1000, 65536, 65537, 1 << 20):         # TestSmallRepeat
        out.append(("repeat_%d" % n, b"x" * n))
    noise = rng.integers(0, 256, 32768, dtype=np.uint8).tobytes()  # TestEncodeNoiseThenRepeats
    out.append(("noise_then_repeats", noise[:16384] + noise[:16384] * 3))
    out.append(("zeros_8mb", bytes(8 << 20)))
    lowent = rng.integers(0, 4, 3 << 20, dtype=np.uint8).tobytes()
    out.append(("lowentropy_3mb", lowent))
    far = bytearray(rng.integers(0, 256, 2500000, dtype=np.uint8).tobytes())  # copy3 offsets
    far[2300000:2400000] = far[0:100000]
    far[2450000:2460000] = far[100000:110000]
    out.append(("far_matches", bytes(far)))
    return out


def random_structure(rng, max_n=2500000):
    """One input of a fuzz campaign (fuzz_test.go:31 FuzzEncodingBlocks in spirit): a random size
    over every encoder size class and one of six structures -- noise, a small alphabet, phrases from
    a vocabulary, noise with pasted copies at offsets / lengths around every token boundary
    (1, 63/64/65, 1024/1025, 65535..65600, far), byte runs, a noisy period."""
    kind = int(rng.integers(0, 6))
    n = int(rng.choice([rng.integers(17, 200), rng.integers(200, 5000), rng.integers(5000, 70000),
                        rng.integers(70000, 600000), rng.integers(600000, max_n)], p=[0.2, 0.3, 0.3, 0.15, 0.05]))
    if kind == 0:
        d = rng.integers(0, 256, n, dtype=np.uint8)
    elif kind == 1:
        d = (rng.integers(0, int(rng.integers(2, 20)), n, dtype=np.uint8) + 48).astype(np.uint8)
    elif kind == 2:
        vocab = [rng.integers(97, 123, int(rng.integers(2, 12)), dtype=np.uint8) for _ in range(int(rng.integers(4, 400)))]
        parts, tot = [], 0
        while tot < n:
            w = vocab[int(rng.integers(0, len(vocab)))]
            parts += [w, np.array([32], dtype=np.uint8)]
            tot += len(w) + 1
        d = np.concatenate(parts)[:n]
    elif kind == 3:
        d = rng.integers(0, 256, n, dtype=np.uint8)
        for _ in range(int(n * rng.uniform(0.001, 0.05))):
            ln = int(rng.choice([4, 5, 6, 7, 8, 9, 11, 12, 13, 20, 24, 25, 40, 64, 65, 300, 5000]))
            if n < ln + 2:
                continue
            p = int(rng.integers(1, n - ln))
            off = min(int(rng.choice([1, 2, 3, 4, 8, 63, 64, 65, 1000, 1024, 1025, 65535, 65536, 65599, 65600, 70000, max(1, p)])), p)
            for k in range(ln):
                d[p + k] = d[p + k - off]
    elif kind == 4:
        d = np.repeat(rng.integers(0, 256, n // 3 + 1, dtype=np.uint8), rng.integers(1, 9, n // 3 + 1))[:n]
        if d.size < 17:
            d = np.zeros(17, dtype=np.uint8)
    else:
        per = rng.integers(0, 256, int(rng.integers(1, 3000)), dtype=np.uint8)
        d = np.tile(per, n // per.size + 1)[:n].copy()
        idx = rng.integers(0, n, int(n * rng.uniform(0, 0.02)))
        d[idx] = rng.integers(0, 256, idx.size, dtype=np.uint8)
    return np.ascontiguousarray(d, dtype=np.uint8)
