"""Oracle on the reference's corpora (CPU only).

Mirrors minlz_test.go:138-194 `roundtrip`, :1538 TestDataRoundtrips,
fuzz_test.go:31 FuzzEncodingBlocks and :120 FuzzDecodeBlock: every input
round-trips at L1 and L2 within MaxEncodedLen, and on adversarial blocks the
oracle agrees accept/reject + bytes with an independent spec decoder.
"""
import os

import numpy as np
import pytest

import corpus
import spec_decoder

REF = "/root/reference/testdata"


def _inputs(name):
    return list(corpus.load_zip(corpus.golden_path(name)))


def _roundtrip(oracle, data):
    for level in (-1, 1, 2):
        enc = oracle.encode(data, level)
        assert isinstance(enc, bytes)
        assert len(enc) <= oracle.max_encoded_len(len(data))
        assert oracle.decoded_len(enc) == len(data)
        assert oracle.decode(enc) == data
        te = oracle.try_encode(data, level)
        if te is not None:
            assert len(te) < len(data) and oracle.decode(te) == data


def test_roundtrip_twain(oracle):
    data = open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt"), "rb").read()
    _roundtrip(oracle, data)
    # sizes cross-checked against an independent restatement
    assert len(oracle.encode(data, 1)) == 10337
    assert len(oracle.encode(data, 2)) == 9301


def test_roundtrip_enc_regressions(oracle):
    for name, data in _inputs("enc_regressions.zip"):
        _roundtrip(oracle, data)


def test_roundtrip_raw_sample(oracle):
    for name, data in _inputs("block-corpus-raw-sample.zip"):
        _roundtrip(oracle, data)
    for name, data in _inputs("block-corpus-enc-sample.zip"):
        _roundtrip(oracle, data)


def test_small_and_edge(oracle):
    # encode.go:83-85,223-229: < 16 bytes stored raw; empty -> single 0 byte
    assert oracle.encode(b"", 1) == b"\x00"
    for n in range(1, 16):
        d = bytes(range(n))
        assert oracle.encode(d, 1) == b"\x00\x00" + d
        assert oracle.decode(b"\x00\x00" + d) == d
    assert oracle.decode(b"\x00") == b""
    assert oracle.encode(b"x" * 100, 7) == oracle.ERR_INVALID_LEVEL
    rnd = np.random.default_rng(3).integers(0, 256, 100000, dtype=np.uint8).tobytes()
    assert oracle.encode(rnd, 1) == b"\x00\x00" + rnd  # incompressible -> stored
    assert oracle.try_encode(rnd, 1) is None
    for n in (16, 17, 31, 32, 33, 63, 64, 65, 255, 256, 65535, 65536, 65537):
        _roundtrip(oracle, (b"abcdefgh" * (n // 8 + 1))[:n])
        _roundtrip(oracle, bytes(n))


def test_zeros_8mb(oracle):
    # minlz_test.go:1538 TestDataRoundtrips: 8 MiB of zeros
    data = bytes(8 << 20)
    for level in (1, 2):
        enc = oracle.encode(data, level)
        assert len(enc) < 64 and oracle.decode(enc) == data


def _check_dec(oracle, blob):
    want = spec_decoder.decode_block(blob)
    got = oracle.decode(blob)
    if blob[:1] != b"\x00":
        # Snappy/S2 fallback territory: out of scope for both
        assert got in (oracle.ERR_UNSUPPORTED, oracle.ERR_CORRUPT, oracle.ERR_TOO_LARGE) or isinstance(got, bytes)
        return
    if want is None:
        assert not isinstance(got, bytes), "oracle accepted a block the spec decoder rejects"
    else:
        assert got == want


def test_decode_adversarial(oracle):
    n = 0
    for name, blob in _inputs("dec-block-regressions.zip") + _inputs("block-corpus-dec.zip"):
        if len(blob) > 20000:
            continue  # keep the pure-Python spec decoder fast
        _check_dec(oracle, blob)
        n += 1
    assert n > 500


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_full_raw_corpus_when_reference_present(oracle):
    for name, data in corpus.load_zip(os.path.join(REF, "fuzz/block-corpus-raw.zip")):
        if len(data) > (1 << 18):
            continue
        _roundtrip(oracle, data)
