"""Stream layer: host framing logic (CPU) and GPU parity of Writer / Reader.

Reference tests mirrored: minlz_test.go:1092-1094 (CRC KAT), :707-869 and
writer_test.go:31-72 (option matrix round trips), reader error cases
(reader.go:273-351: missing header, bad CRC, truncated chunk, EOF size)."""
import io

import numpy as np
import pytest

import corpus
import patterns
import stream_ref
from minlz_b200 import stream as mzs


def test_header_and_uvarint():
    # writer.go:1553-1556: 2 MiB -> 11, 4 KiB -> 2, 8 MiB -> 13
    assert mzs.make_header(2 << 20) == b"\xff\x06\x00\x00MinLz\x0b"
    assert mzs.make_header(4 << 10)[-1] == 2
    assert mzs.make_header(8 << 20)[-1] == 13
    assert mzs.make_header((1 << 20) + 1)[-1] == 11
    for v in (0, 1, 127, 128, 300, 1 << 20, (1 << 32) - 1, 1 << 40):
        assert mzs._read_uvarint(mzs._uvarint(v)) == (v, len(mzs._uvarint(v)))
        assert mzs._uvarint(v) == stream_ref.uvarint(v)


def test_writer_options():
    with pytest.raises(ValueError):
        mzs.Writer(None, mzs.WriterBlockSize(1024))
    with pytest.raises(ValueError):
        mzs.Writer(None, mzs.WriterBlockSize((8 << 20) + 1))
    with pytest.raises(Exception):
        mzs.Writer(None, mzs.WriterLevel(3))
    w = mzs.Writer(None)
    assert w.level == 2 and w.block_size == 2 << 20  # writer.go:40, minlz.go:109


def test_empty_stream_is_eof_only():
    buf = io.BytesIO()
    w = mzs.NewWriter(buf)
    w.Close()
    assert buf.getvalue() == b"\x20\x01\x00\x00\x00" == stream_ref.encode_stream(b"", 1, 1 << 20)
    assert w.Written() == (0, 5)
    assert mzs.NewReader(io.BytesIO(buf.getvalue())).Read() == b""


def test_reader_host_side_errors():
    """Error paths that never reach the device."""
    good = stream_ref.encode_stream(b"hello world, hello world, hello world", 0, 4 << 10)
    with pytest.raises(mzs.ErrCorrupt):   # first chunk is not the stream identifier (reader.go:276-283)
        mzs.NewReader(io.BytesIO(good[10:])).Read()
    with pytest.raises(mzs.ErrCorrupt):   # truncated inside a chunk
        mzs.NewReader(io.BytesIO(good[:-9])).Read()
    with pytest.raises(mzs.ErrUnsupported):
        mzs.NewReader(io.BytesIO(b"\xff\x06\x00\x00S2sTwO")).Read()
    with pytest.raises(mzs.ErrCorrupt):   # block size indicator > 13
        mzs.NewReader(io.BytesIO(b"\xff\x06\x00\x00MinLz\x0e")).Read()
    with pytest.raises(mzs.ErrTooLarge):  # stream block size above the reader's limit
        mzs.NewReader(io.BytesIO(good), mzs.ReaderMaxBlockSize(1024)).Read()
    with pytest.raises(mzs.ErrUnsupported):  # reserved non-skippable chunk
        mzs.NewReader(io.BytesIO(good[:10] + b"\x05\x00\x00\x00")).Read()


# --------------------------------------------------------------------- GPU ----

def _payloads():
    twain = open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt"), "rb").read()
    rnd = np.random.default_rng(11).integers(0, 256, 300000, dtype=np.uint8).tobytes()
    return [("twain", twain), ("twain_x40", twain * 40), ("mixed", twain * 3 + rnd + patterns.offset2(70000)),
            ("tiny", b"abc"), ("testdata", patterns.generate_test_data(1 << 20))]


@pytest.mark.gpu
@pytest.mark.parametrize("level", [-1, 0, 1, 2])
@pytest.mark.parametrize("block_size", [4 << 10, 64 << 10, 1 << 20])
def test_writer_bytes_equal_reference_framing(oracle, level, block_size):
    """Writer output == framing restatement over oracle-encoded blocks, byte for byte."""
    for name, data in _payloads():
        want = stream_ref.encode_stream(data, level, block_size)
        buf = io.BytesIO()
        w = mzs.NewWriter(buf, mzs.WriterLevel(level), mzs.WriterBlockSize(block_size), mzs.WriterConcurrency(7))
        # feed in uneven pieces like minlz_test.go:1028 TestNewWriter
        pos = 0
        for step in (1, 10, 1000, 70000, 1 << 22):
            w.Write(data[pos:pos + step])
            pos += step
        w.Write(data[pos:])
        w.Close()
        assert buf.getvalue() == want, (name, level, block_size)
        assert w.Written() == (len(data), len(want))
        # and it reads back
        assert mzs.NewReader(io.BytesIO(want), mzs.ReaderConcurrency(5)).Read() == data


@pytest.mark.gpu
def test_crc_kat_and_blocks(oracle):
    import minlz_b200 as mz
    from minlz_b200 import _lib
    blobs = [b"abcd", b"", b"a", bytes(range(256)) * 37, patterns.generate_test_data(1 << 20), b"x" * 33, b"y" * 31]
    flat = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    np.cumsum([len(b) for b in blobs], out=off[1:])
    crc = np.zeros(len(blobs), dtype=np.uint32)
    r = _lib.load().mzcu_crc32c_blocks(-1, len(blobs), flat.ctypes.data, off.ctypes.data, crc.ctypes.data)
    assert r == 0
    assert int(crc[0]).to_bytes(4, "little") == bytes([0x68, 0x10, 0xe6, 0xb6])  # minlz_test.go:1092-1094
    for b, c in zip(blobs, crc):
        assert int(c) == oracle.crc(b)


@pytest.mark.gpu
def test_reader_stream_errors(oracle):
    twain = open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt"), "rb").read() * 8
    good = stream_ref.encode_stream(twain, 1, 16 << 10)
    assert mzs.decode_stream(good) == twain
    # flip a byte inside the 3rd chunk's payload: CRC or corrupt, and the data before it is delivered
    bad = bytearray(good)
    p = 10
    for _ in range(2):
        p += 4 + (bad[p + 1] | bad[p + 2] << 8 | bad[p + 3] << 16)
    bad[p + 40] ^= 0x55
    with pytest.raises((mzs.ErrCRC, mzs.ErrCorrupt)) as ei:
        mzs.decode_stream(bytes(bad))
    assert ei.value.partial == twain[:2 * (16 << 10)]
    # wrong stored CRC only
    bad = bytearray(good)
    bad[10 + 4] ^= 1
    with pytest.raises(mzs.ErrCRC):
        mzs.decode_stream(bytes(bad))
    assert mzs.decode_stream(bytes(bad), mzs.ReaderIgnoreCRC()) == twain
    # EOF size mismatch (reader.go:478-486)
    bad = good[:-1] + bytes([good[-1] ^ 1])
    with pytest.raises(mzs.ErrCorrupt):
        mzs.decode_stream(bad)
    # concatenated streams (SPEC 4.1 / 4.6) and padding chunks are accepted
    two = good + stream_ref.chunk(0xfe, b"\x00" * 100) + good
    assert mzs.decode_stream(two) == twain + twain
    # type 0x03 chunk: CRC over the compressed bytes
    blk = twain[:5000]
    tok = oracle.encode_block(blk, 1)
    body = stream_ref.uvarint(len(blk)) + tok
    s3 = stream_ref.header(8 << 10) + stream_ref.chunk(0x03, oracle.crc(tok).to_bytes(4, "little") + body) + \
        stream_ref.chunk(0x20, stream_ref.uvarint(len(blk)))
    assert mzs.decode_stream(s3) == blk


@pytest.mark.gpu
def test_config4_shape_stream_roundtrip():
    """BASELINE config 4 shape (stream Writer/Reader, synthetic log text, 2 MiB
    blocks) at 128 MiB: Writer -> Reader reproduces the input; property checks
    that do not need the oracle (size-independent): chunk walk, EOF size."""
    import synth
    data = synth.make_blocks("log", 64, 2 << 20, device="cuda").cpu().numpy().tobytes()
    buf = io.BytesIO()
    w = mzs.NewWriter(buf, mzs.WriterLevel(1), mzs.WriterBlockSize(2 << 20), mzs.WriterConcurrency(32))
    w.ReadFrom(io.BytesIO(data))
    w.Close()
    blob = buf.getvalue()
    assert len(blob) < len(data) // 3
    # walk the chunks: header, 64 x type 0x02, EOF with the total size
    assert blob[:10] == mzs.make_header(2 << 20)
    p, kinds = 10, []
    while p < len(blob):
        kinds.append(blob[p])
        p += 4 + (blob[p + 1] | blob[p + 2] << 8 | blob[p + 3] << 16)
    assert kinds == [2] * 64 + [0x20] and p == len(blob)
    out = io.BytesIO()
    assert mzs.NewReader(io.BytesIO(blob)).WriteTo(out) == len(data)
    assert out.getvalue() == data


@pytest.mark.gpu
@pytest.mark.parametrize("level", [1, -1])
def test_sliced_upload_incompressible_blocks_crc(oracle, level):
    """Sliced host upload (>= 64 equal blocks of >= 256 KiB): blocks that bail out early
    let the encode launch finish while later slices are still in flight; the checksum
    and the pack must still see the whole source (writer.go:672: CRC of the raw block)."""
    from minlz_b200 import _lib
    lib = _lib.load()
    nblk, bs = 96, 1 << 20
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, (nblk, bs), dtype=np.uint8)
    data[::7, : bs // 2] = 0x41  # a few compressible blocks among the stored ones
    flat = data.reshape(-1)
    soff = np.arange(nblk + 1, dtype=np.uint64) * bs
    for rep in range(2):
        dst = np.zeros(flat.size + 64, dtype=np.uint8)
        poff = np.zeros(nblk + 1, dtype=np.uint64)
        crc = np.zeros(nblk, dtype=np.uint32)
        r = lib.mzcu_stream_encode_blocks(-1, level, nblk, flat.ctypes.data, soff.ctypes.data, dst.ctypes.data, dst.size,
                                          poff.ctypes.data, crc.ctypes.data)
        assert r == 0, lib.mzcu_last_error()
        for i in range(nblk):
            assert int(crc[i]) == oracle.crc(data[i].tobytes()), (rep, i)
        for i in range(0, nblk, 5):
            assert dst[int(poff[i]):int(poff[i + 1])].tobytes() == oracle.encode_block(data[i], level), (rep, i)
