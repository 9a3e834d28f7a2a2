"""GPU parity tests: every call goes through the C ABI of libminlz_cuda.so and
is checked bit-for-bit against the CPU oracle (tests-only) on the same inputs.

Reference tests mirrored: minlz_test.go:632 TestDecodeGoldenInput, :138-194
roundtrip (+ :202-253 drivers, :780, :1538), decode_asm_test.go:28-352,
fuzz_test.go:31 FuzzEncodingBlocks, :120 FuzzDecodeBlock.
"""
import numpy as np
import pytest

import corpus
import patterns

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import minlz_b200 as mz  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected but no CUDA device is visible")
    assert mz.device_count() >= 1


def _cat(blobs):
    sizes = [len(b) for b in blobs]
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    np.cumsum(sizes, out=off[1:])
    flat = np.frombuffer(b"".join(blobs), dtype=np.uint8) if off[-1] else np.zeros(0, dtype=np.uint8)
    return flat, off


def _all_inputs():
    items = patterns.reference_patterns() + patterns.roundtrip_inputs()
    items += [("twain", open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt"), "rb").read())]
    for z in ("enc_regressions.zip", "block-corpus-raw-sample.zip", "block-corpus-enc-sample.zip"):
        items += list(corpus.load_zip(corpus.golden_path(z)))
    return items


# ---------------------------------------------------------------- decode ----

def test_decode_golden():
    want = open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt"), "rb").read()
    mzb = open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt.mzb"), "rb").read()
    assert mz.DecodedLen(mzb) == len(want)
    assert mz.Decode(None, mzb) == want


@pytest.mark.parametrize("level", [-1, 1, 2])
def test_decode_oracle_encoded_inputs(oracle, level):
    """Blocks produced by the oracle encoders decode to the original on the GPU
    (seam level: token streams without header, exact dst length)."""
    items = [(n, d) for n, d in _all_inputs() if len(d) >= 1]
    streams, raws = [], []
    for name, data in items:
        tok = oracle.encode_block(data, level)
        if not tok:
            continue  # incompressible or < 16 bytes: no token stream at the seam
        streams.append(tok)
        raws.append(data)
    assert len(streams) > (60 if level == -1 else 100)
    src, soff = _cat(streams)
    _, doff = _cat(raws)
    dst, status = mz.decode_blocks(src, soff, doff)
    assert not status.any(), np.nonzero(status)[0][:10]
    want = b"".join(raws)
    got = dst.tobytes()
    if got != want:
        for i in range(len(raws)):
            a, b = int(doff[i]), int(doff[i + 1])
            assert got[a:b] == want[a:b], "block %d (%d bytes) differs" % (i, b - a)


def test_decode_adversarial_blocks_match_oracle(oracle):
    """fuzz_test.go:120: corrupt / hostile blocks.  Accept/reject and bytes must
    equal the oracle's; dst must not be written outside its range."""
    blobs = [b for _, b in corpus.load_zip(corpus.golden_path("block-corpus-dec.zip"))]
    blobs += [b for _, b in corpus.load_zip(corpus.golden_path("dec-block-regressions.zip"))]
    res = mz.DecodeBatch(blobs)
    n_ok = 0
    for blob, r in zip(blobs, res):
        want = oracle.decode(blob)
        if isinstance(want, bytes):
            assert r == want
            n_ok += 1
        elif want == oracle.ERR_UNSUPPORTED:
            assert isinstance(r, mz.ErrUnsupported)
        elif want == oracle.ERR_TOO_LARGE:
            assert isinstance(r, mz.ErrTooLarge)
        else:
            assert isinstance(r, mz.ErrCorrupt), (want, r)
    assert n_ok >= 3


def _mutations(oracle, rng, data, level, count):
    tok = bytearray(oracle.encode_block(data, level))
    out = []
    for _ in range(count):
        t = bytearray(tok)
        kind = rng.integers(0, 4)
        if kind == 0:
            for _ in range(int(rng.integers(1, 4))):
                t[int(rng.integers(0, len(t)))] = int(rng.integers(0, 256))
        elif kind == 1:
            t = t[:int(rng.integers(0, len(t)))]
        elif kind == 2:
            p = int(rng.integers(0, len(t)))
            t[p:p] = bytes(rng.integers(0, 256, int(rng.integers(1, 8)), dtype=np.uint8))
        else:
            p = int(rng.integers(0, len(t)))
            del t[p:p + int(rng.integers(1, 8))]
        out.append(bytes(t))
    return out


def test_decode_mutated_streams_match_oracle(oracle):
    """Seam-level fuzz: mutated token streams against the oracle's status and
    bytes, with guard bytes around every dst range."""
    rng = np.random.default_rng(7)
    base = [patterns.generate_test_data(5000), patterns.fused_lits(10000), patterns.offset2(6000),
            open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt"), "rb").read()]
    streams, dlens = [], []
    for data in base:
        for level in (1, 2):
            for m in _mutations(oracle, rng, data, level, 150):
                streams.append(m)
                dlens.append(len(data) if rng.random() < 0.8 else int(rng.integers(0, 2 * len(data))))
    src, soff = _cat(streams)
    doff = np.zeros(len(dlens) + 1, dtype=np.uint64)
    np.cumsum(dlens, out=doff[1:])
    dst, status = mz.decode_blocks(src, soff, doff)
    n_ok = n_bad = 0
    for i, (st, dl) in enumerate(zip(streams, dlens)):
        wst, wout = oracle.decode_block(st, dl)
        assert int(status[i]) == wst, "stream %d: status %d, oracle %d" % (i, status[i], wst)
        if wst == 0:
            assert dst[int(doff[i]):int(doff[i + 1])].tobytes() == wout
            n_ok += 1
        else:
            n_bad += 1
    assert n_ok > 10 and n_bad > 100


def test_decode_device_api_no_overrun(oracle):
    """Device-pointer entry point.  Guard ranges (empty stream, 64 bytes of
    dst: corrupt, nothing may be written) sit between the real blocks and must
    keep their 0xfe fill (fuzz_test.go:169-182 uses guard bytes the same way)."""
    data = [patterns.generate_test_data(100000), patterns.short_repeat(3, 9), b"x" * 70000, patterns.offset2(65549)]
    streams = [oracle.encode_block(d, 1) for d in data]
    streams.append(streams[0][:-3])  # truncated -> corrupt, may only touch its own range
    lens = [len(d) for d in data] + [len(data[0])]
    g_streams, g_lens = [], []
    for s_, n in zip(streams, lens):
        g_streams += [s_, b""]
        g_lens += [n, 64]
    src, soff = _cat(g_streams)
    d_off = np.zeros(len(g_lens) + 1, dtype=np.int64)
    np.cumsum(g_lens, out=d_off[1:])
    dev = torch.device("cuda:0")
    t_src = torch.from_numpy(src.copy()).to(dev)
    t_soff = torch.from_numpy(soff.astype(np.int64)).to(dev)
    t_doff = torch.from_numpy(d_off).to(dev)
    t_dst = torch.full((int(d_off[-1]),), 0xfe, dtype=torch.uint8, device=dev)
    t_status = torch.full((len(g_lens),), -1, dtype=torch.int32, device=dev)
    mz.decode_blocks_dev(t_src, t_soff, t_dst, t_doff, t_status)
    torch.cuda.synchronize()
    status = t_status.cpu().numpy()
    out = t_dst.cpu().numpy()
    for k in range(len(g_lens) // 2):
        g0, g1 = d_off[2 * k + 1], d_off[2 * k + 2]
        assert status[2 * k + 1] == 1                      # guard: d != len(dst)
        assert (out[g0:g1] == 0xfe).all(), "guard %d overwritten" % k
    for k, d in enumerate(data):
        assert status[2 * k] == 0
        assert out[d_off[2 * k]:d_off[2 * k + 1]].tobytes() == d
    assert status[2 * len(data)] == 1


# ---------------------------------------------------------------- encode ----

@pytest.mark.parametrize("level", [-1, 1, 2])
def test_encode_bytes_equal_oracle(oracle, level):
    """Seam level: the CUDA encoder's token stream is byte-identical to the
    oracle's restatement of the Go path, including the 0 = incompressible."""
    items = _all_inputs()
    raws = [d for _, d in items]
    src, soff = _cat(raws)
    dst, doff, out_len = mz.encode_blocks(src, soff, level)
    bad = []
    for i, (name, data) in enumerate(items):
        want = oracle.encode_block(data, level)
        got = dst[int(doff[i]):int(doff[i]) + int(out_len[i])].tobytes()
        if got != want:
            bad.append((name, len(data), len(got), len(want)))
    assert not bad, bad[:10]


@pytest.mark.parametrize("level", [-1, 1, 2])
def test_encode_api_roundtrip(oracle, level):
    """minlz_test.go:138-194 roundtrip through the block API mirror."""
    for name, data in patterns.roundtrip_inputs()[:24] + patterns.reference_patterns()[:12]:
        enc = mz.Encode(None, data, level)
        assert len(enc) <= mz.MaxEncodedLen(len(data))
        assert enc == oracle.encode(data, level), name
        assert mz.DecodedLen(enc) == len(data)
        assert mz.Decode(None, enc) == data, name
        te = mz.TryEncode(None, data, level)
        assert te == oracle.try_encode(data, level)


def test_api_edge_cases(oracle):
    # encode.go:83-85,223-229 and decode.go:55-57
    assert mz.Encode(None, b"", 1) == b"\x00"
    for n in range(1, 16):
        d = bytes(range(n))
        assert mz.Encode(None, d, 1) == b"\x00\x00" + d
        assert mz.Decode(None, b"\x00\x00" + d) == d
    assert mz.Decode(None, b"\x00") == b""
    with pytest.raises(mz.ErrInvalidLevel):
        mz.Encode(None, b"x" * 100, 9)
    with pytest.raises(mz.ErrTooLarge):
        mz.Encode(None, bytes((8 << 20) + 1), 1)
    rnd = np.random.default_rng(3).integers(0, 256, 100000, dtype=np.uint8).tobytes()
    assert mz.Encode(None, rnd, 1) == b"\x00\x00" + rnd       # incompressible -> stored
    assert mz.TryEncode(None, rnd, 1) is None
    assert mz.Encode(None, rnd, mz.LevelUncompressed) == b"\x00\x00" + rnd
    twain = open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt"), "rb").read()
    good = oracle.encode(twain, 1)
    assert oracle.decode(good[:-2]) == oracle.ERR_CORRUPT
    with pytest.raises(mz.ErrCorrupt) as ei:
        mz.Decode(None, good[:-2])
    assert ei.value.partial is not None and len(ei.value.partial) == len(twain)
    # encode_l1.go:194 quirk kept by the restatement: the bail test uses the
    # match END (s), so a long first match after >3 literals reports 0
    assert mz.Encode(None, b"abcdefgh" * 100, 1) == b"\x00\x00" + b"abcdefgh" * 100
    with pytest.raises(mz.ErrUnsupported):
        mz.Decode(None, b"\x05hello")  # Snappy/S2 fallback lives in host Go
    big = bytes(8 << 20)
    enc = mz.Encode(None, big, 1)
    assert enc == oracle.encode(big, 1) and mz.Decode(None, enc) == big


def test_synthetic_batch_parity(oracle):
    """BASELINE configs 2/3 shape at a size the oracle finishes in seconds:
    32 x 1 MiB json blocks, encode bytes == oracle, decode == input."""
    import synth
    blocks = synth.make_blocks("json", 32, 1 << 20, device="cuda").cpu().numpy()
    src = blocks.reshape(-1)
    soff = (np.arange(33, dtype=np.uint64) << 20)
    for level in (1, 2):
        dst, doff, out_len = mz.encode_blocks(src, soff, level)
        enc = [oracle.encode_block(blocks[i], level) for i in range(32)]
        for i in range(32):
            got = dst[int(doff[i]):int(doff[i]) + int(out_len[i])].tobytes()
            assert got == enc[i], "level %d block %d" % (level, i)
        csrc, csoff = _cat(enc)
        out, status = mz.decode_blocks(csrc, csoff, soff)
        assert not status.any()
        assert np.array_equal(out, src)


def test_full_size_roundtrip_property():
    """BASELINE configs 2+3 at full size (4096 x 1 MiB): encode on the GPU,
    decode on the GPU, output equals input; all device resident."""
    import synth
    dev = torch.device("cuda:0")
    nblk, bs = 4096, 1 << 20
    src = synth.make_blocks("json", nblk, bs, device=dev).reshape(-1)
    soff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * bs
    cap = bs + 16
    doff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * cap
    enc = torch.empty(nblk * cap, dtype=torch.uint8, device=dev)
    out_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
    mz.encode_blocks_dev(src, soff, enc, doff, out_len, 1)
    torch.cuda.synchronize()
    lens = out_len.to(torch.int64)
    assert int(lens.min()) > 0 and int(lens.max()) < bs // 2
    # compact the token streams so that src_off is a dense offset table
    coff = torch.zeros(nblk + 1, dtype=torch.int64, device=dev)
    coff[1:] = torch.cumsum(lens, 0)
    idx = torch.arange(int(coff[-1]), device=dev)
    blk = torch.searchsorted(coff[1:], idx, right=True)
    comp = enc[doff[blk] + (idx - coff[blk])]
    del enc, idx, blk
    dec = torch.empty(nblk * bs, dtype=torch.uint8, device=dev)
    status = torch.full((nblk,), -1, dtype=torch.int32, device=dev)
    mz.decode_blocks_dev(comp, coff, dec, soff, status)
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0
    assert torch.equal(dec, src)


def test_encode_packed_host_api(oracle):
    """mzcu_encode_blocks_packed: dense output + offsets, feeds decode as is."""
    items = patterns.reference_patterns()[:20] + [("rnd", np.random.default_rng(5).integers(0, 256, 5000, dtype=np.uint8).tobytes())]
    raws = [d for _, d in items]
    src, soff = _cat(raws)
    dst = np.zeros(max(len(src), 1), dtype=np.uint8)
    poff = np.zeros(len(raws) + 1, dtype=np.uint64)
    total = mz.encode_blocks_packed_into(src, soff, dst, poff, 1)
    assert total == int(poff[-1])
    for i, data in enumerate(raws):
        want = oracle.encode_block(data, 1)
        assert dst[int(poff[i]):int(poff[i + 1])].tobytes() == want
    # blocks the encoder gave up on have empty ranges; decode the rest
    keep = [i for i in range(len(raws)) if poff[i + 1] > poff[i]]
    streams = [dst[int(poff[i]):int(poff[i + 1])].tobytes() for i in keep]
    csrc, csoff = _cat(streams)
    _, doff = _cat([raws[i] for i in keep])
    out, status = mz.decode_blocks(csrc, csoff, doff)
    assert not status.any() and out.tobytes() == b"".join(raws[i] for i in keep)


@pytest.mark.parametrize("level", [-1, 1])
@pytest.mark.parametrize("last", [256 << 10, 100001])
def test_encode_packed_sliced_source(oracle, level, last):
    """Equal-sized contiguous blocks take the sliced host->device pipeline (the
    kernel starts before the source has arrived and chases the arrival front);
    the bytes must not depend on it.  The last block may be short."""
    import synth
    B, nblk = 256 << 10, 80
    body = synth.make_blocks("json", nblk - 1, B, seed=7).cpu().numpy().reshape(-1)
    tail = synth.make_blocks("log", 1, B, seed=8).cpu().numpy().reshape(-1)[:last]
    src = np.ascontiguousarray(np.concatenate([body, tail]))
    soff = np.array([i * B for i in range(nblk)] + [(nblk - 1) * B + last], dtype=np.uint64)
    dst = np.zeros(len(src), dtype=np.uint8)
    poff = np.zeros(nblk + 1, dtype=np.uint64)
    total = mz.encode_blocks_packed_into(src, soff, dst, poff, level)
    assert total == int(poff[-1]) and 0 < total < len(src)
    for i in list(range(0, nblk, 9)) + [nblk - 2, nblk - 1]:
        data = src[int(soff[i]):int(soff[i + 1])].tobytes()
        assert dst[int(poff[i]):int(poff[i + 1])].tobytes() == oracle.encode_block(data, level), i
    out, status = mz.decode_blocks(dst[:total], poff, soff)
    assert not status.any() and np.array_equal(out, src)


def test_concurrent_callers(oracle):
    """The seam must tolerate concurrent calls from arbitrary threads
    (writer.go:670 spawns one goroutine per block; SURVEY 8b threading row)."""
    import threading
    rng = np.random.default_rng(21)
    inputs = [patterns.generate_test_data(200000 + 1000 * i) for i in range(4)] + \
             [rng.integers(0, 8, 150000, dtype=np.uint8).tobytes() for _ in range(4)]
    want = [(oracle.encode(d, 1), oracle.encode(d, 2)) for d in inputs]
    errors = []

    def worker(k):
        try:
            for it in range(6):
                d = inputs[(k + it) % len(inputs)]
                w1, w2 = want[(k + it) % len(inputs)]
                e1 = mz.Encode(None, d, 1)
                e2 = mz.Encode(None, d, 2)
                if e1 != w1 or e2 != w2:
                    errors.append("encode mismatch in thread %d" % k)
                if mz.Decode(None, e1) != d or mz.Decode(None, w2) != d:
                    errors.append("decode mismatch in thread %d" % k)
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
        assert not t.is_alive(), "worker hung"
    assert not errors, errors[:3]


def test_random_structures_roundtrip(oracle):
    """fuzz_test.go:31 FuzzEncodingBlocks in spirit: random mixtures of literals,
    short/long repeats and far copies at random sizes; encoder bytes == oracle,
    decoder output == input, for both levels, all in two batched calls."""
    rng = np.random.default_rng(99)
    blocks = []
    for _ in range(160):
        n = int(rng.choice([17, 100, 4000, 65536, 65537, 150000, 400000]))
        n += int(rng.integers(0, 50))
        alpha = int(rng.choice([2, 4, 16, 256]))
        base = rng.integers(0, alpha, n, dtype=np.uint8)
        for _ in range(int(rng.integers(0, 30))):   # paste earlier slices forward
            ln = int(rng.integers(4, min(n // 2, 5000)))
            a = int(rng.integers(0, n - ln))
            b = int(rng.integers(0, n - ln))
            base[b:b + ln] = base[a:a + ln]
        blocks.append(base.tobytes())
    src, soff = _cat(blocks)
    for level in (-1, 1, 2):
        dst, doff, out_len = mz.encode_blocks(src, soff, level)
        streams, raws = [], []
        for i, data in enumerate(blocks):
            want = oracle.encode_block(data, level)
            got = dst[int(doff[i]):int(doff[i]) + int(out_len[i])].tobytes()
            assert got == want, (level, i, len(data))
            if got:
                streams.append(got)
                raws.append(data)
        csrc, csoff = _cat(streams)
        _, cdoff = _cat(raws)
        out, status = mz.decode_blocks(csrc, csoff, cdoff)
        assert not status.any()
        assert out.tobytes() == b"".join(raws)


def test_random_structures_fuzz_batch(oracle):
    """The fuzz generator of tests/patterns.py (every size class, six structures): 300 inputs, all
    three levels, encoder bytes == oracle (Go flavour) in batched calls."""
    rng = np.random.default_rng(4242)
    blocks = [patterns.random_structure(rng, max_n=1200000).tobytes() for _ in range(300)]
    src, soff = _cat(blocks)
    for level in (-1, 1, 2):
        dst, doff, out_len = mz.encode_blocks(src, soff, level)
        for i, data in enumerate(blocks):
            got = dst[int(doff[i]):int(doff[i]) + int(out_len[i])].tobytes()
            assert got == oracle.encode_block(data, level), (level, i, len(data))


@pytest.mark.parametrize("level", [1, 2, -1])
def test_go_flavour_8mib_class(oracle, level):
    """The library's default flavour (pure-Go functions) at the 8 MiB block size of BASELINE
    config 5, with real data: text, binary structs, a far copy beyond the copy3 range (the
    minSrcPos guards of encode_l1.go:83,147 / encode_l2.go:118 decide), a 3 MiB period and a
    large-offset pattern.  Encoder bytes == oracle restatement, and they decode to the input."""
    import synth
    assert mz.get_encoder_flavor() == mz.FlavorGo
    items = [(k, synth.make_blocks(k, 1, 8 << 20).numpy()[0].tobytes()) for k in ("text", "binary")]
    items.append(("large_offset", patterns.large_offset(8 << 20, (2 << 20) + 70000)))
    rng = np.random.default_rng(7)
    per = rng.integers(0, 256, 3 << 20, dtype=np.uint8)
    items.append(("period3MiB", np.concatenate([per, per, per[: 2 << 20]]).tobytes()))
    text = synth.make_blocks("text", 1, 8 << 20).numpy()[0].copy()
    text[5 << 20:] = text[: 3 << 20]
    items.append(("text-far-copy", text.tobytes()))
    raws = [d for _, d in items]
    src, soff = _cat(raws)
    dst, doff, out_len = mz.encode_blocks(src, soff, level)
    streams = []
    for i, (name, data) in enumerate(items):
        want = oracle.encode_block(data, level)
        got = dst[int(doff[i]):int(doff[i]) + int(out_len[i])].tobytes()
        assert got == want, (name, level, len(got), len(want))
        streams.append(got)
    live = [i for i in range(len(items)) if streams[i]]
    csrc, csoff = _cat([streams[i] for i in live])
    _, cdoff = _cat([raws[i] for i in live])
    out, status = mz.decode_blocks(csrc, csoff, cdoff)
    assert not status.any()
    assert out.tobytes() == b"".join(raws[i] for i in live)
