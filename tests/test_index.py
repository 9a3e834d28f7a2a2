"""Seek index: host logic on CPU, random access through the GPU decoder.

Reference tests mirrored: index_test.go:30 ExampleIndex_Load (Find + Skip over a
non-seekable input), :119 TestSeeking (Seek / ReadAt at every kind of offset),
:273 TestSeekingStreamIndex and :419 ExampleIndexStream (index made from an
existing stream == the writer's own index).  The reference holds no byte fixture
for the index (its tests draw from Go's math/rand), so the format is checked
structurally against index.go and by appendTo <-> Load round trips.
"""
import io

import numpy as np
import pytest

import stream_ref
import synth
from minlz_b200 import index as mzi
from minlz_b200 import stream as mzs
from minlz_b200 import ErrCorrupt, ErrUnsupported


def test_varint_matches_go_encoding():
    # encoding/binary: zig-zag + base 128; known values from the Go documentation examples
    assert mzi.put_varint(0) == b"\x00"
    assert mzi.put_varint(-1) == b"\x01"
    assert mzi.put_varint(1) == b"\x02"
    assert mzi.put_varint(-64) == b"\x7f"
    assert mzi.put_varint(64) == b"\x80\x01"
    assert mzi.put_varint(-(1 << 63)) == b"\xff" * 9 + b"\x01"
    for v in (0, 1, -1, 63, -64, 64, 300, -300, 1 << 40, -(1 << 40), (1 << 63) - 1, -(1 << 63)):
        b = mzi.put_varint(v)
        assert mzi.varint(b + b"zz") == (v, len(b))
    assert mzi.varint(b"") == (0, 0)
    assert mzi.varint(b"\x80\x80")[1] == 0          # buffer too small
    assert mzi.varint(b"\xff" * 11)[1] < 0           # overflow


def _mk(entries, est, total_u, total_c):
    ix = mzi.Index()
    ix.estBlockUncomp = est
    ix.Offsets = list(entries)
    return ix.appendTo(b"", total_u, total_c), ix


def test_append_load_round_trip_and_layout():
    est = 1 << 20
    # regular spacing: hasUncompressed == 0 (index.go:211-226)
    ent = [(10 + i * 400000 + (i * i) % 977, i * est) for i in range(50)]
    blob, _ = _mk(ent, est, 50 * est, 20_000_000)
    assert blob[0] == 0x40 and blob[4:10] == b"s2idx\x00" and blob[-6:] == b"\x00xdi2s"
    assert int.from_bytes(blob[1:4], "little") == len(blob) - 4                  # skippable chunk length
    assert int.from_bytes(blob[-10:-6], "little") == len(blob)                  # total size, fixed width
    ix = mzi.Index()
    assert ix.Load(blob + b"tail") == b"tail"
    assert ix.Offsets == ent and ix.TotalUncompressed == 50 * est and ix.TotalCompressed == 20_000_000
    assert ix.estBlockUncomp == est
    # irregular spacing: uncompressed deltas are stored
    ent2 = [(c, u + (i % 3) * 1000 + (0 if i else 5)) for i, (c, u) in enumerate(ent)]
    blob2, _ = _mk(ent2, est, 60 * est, -1)
    ix2 = mzi.Index()
    ix2.Load(blob2)
    assert ix2.Offsets == ent2 and ix2.TotalCompressed == -1
    assert len(blob2) > len(blob)
    # headers can be stripped and restored (index.go:581-636)
    bare = mzi.RemoveIndexHeaders(blob)
    assert bare is not None and len(bare) == len(blob) - 4 - 6 - 6 - 4
    assert mzi.RestoreIndexHeaders(bare) == blob
    assert mzi.RemoveIndexHeaders(blob[:-1]) is None and mzi.RestoreIndexHeaders(b"") == b""


def test_load_rejects_what_go_rejects():
    blob, _ = _mk([(0, 0), (500000, 1 << 20)], 1 << 20, 2 << 20, 900000)
    with pytest.raises(mzi.ErrUnexpectedEOF):
        mzi.Index().Load(blob[:12])
    with pytest.raises(ErrCorrupt):
        mzi.Index().Load(b"\x41" + blob[1:])
    with pytest.raises(ErrUnsupported):
        mzi.Index().Load(blob[:4] + b"s2idy\x00" + blob[10:])
    with pytest.raises(ErrCorrupt):
        mzi.Index().Load(blob[:-1] + b"X")
    with pytest.raises(mzi.ErrUnexpectedEOF):
        mzi.Index().Load(blob[:-3])
    mzi.Index().Load(b"\x99" + blob[1:])   # legacy S2 index chunk id is accepted (index.go:277)


def test_find_semantics():
    ix = mzi.Index()
    with pytest.raises(ErrCorrupt):
        ix.Find(0)                               # TotalUncompressed unknown
    ix.TotalUncompressed = 10 << 20
    ix.Offsets = [(i * 300000, i << 20) for i in range(10)]
    assert ix.Find(0) == (0, 0)
    assert ix.Find((3 << 20) - 1) == (600000, 2 << 20)
    assert ix.Find(3 << 20) == (900000, 3 << 20)
    assert ix.Find(-1) == (2700000, 9 << 20)      # from the end
    assert ix.Find(10 << 20) == (2700000, 9 << 20)
    with pytest.raises(mzi.ErrUnexpectedEOF):
        ix.Find((10 << 20) + 1)
    with pytest.raises(mzi.ErrUnexpectedEOF):
        ix.Find(-(10 << 20) - 1)
    big = mzi.Index()                              # > 200 entries: binary search branch
    big.TotalUncompressed = 1000 << 20
    big.Offsets = [(i * 1000, i << 20) for i in range(1000)]
    for off in (0, 1, (1 << 20) - 1, 1 << 20, (777 << 20) + 5, (1000 << 20)):
        assert big.Find(off) == ((min(off >> 20, 999)) * 1000, min(off >> 20, 999) << 20)


def test_add_spacing_and_reduce():
    ix = mzi.Index()
    ix.reset(64 << 10)
    assert ix.estBlockUncomp == 1 << 20           # doubled up to minIndexDist (index.go:59-61)
    for i in range(100):
        ix.add(i * 30000, i * (64 << 10))
    assert [u for _, u in ix.Offsets] == [k << 20 for k in range(7)]   # one entry per MiB
    with pytest.raises(Exception):
        ix.add(5, 1 << 30)                         # compressed offset went backwards (index.go:96-98)
    many = mzi.Index()
    many.estBlockUncomp = 1 << 20
    many.Offsets = [(i * 10, i << 20) for i in range(mzi.MAX_INDEX_ENTRIES + 5)]
    many.reduce()
    assert len(many.Offsets) < mzi.MAX_INDEX_ENTRIES and many.estBlockUncomp == 2 << 20
    assert many.Offsets[1] == (20, 2 << 20)
    light = mzi.Index()
    light.estBlockUncomp = 1 << 20
    light.Offsets = [(i * 10, i << 20) for i in range(10)]
    light.reduceLight()                            # index.go:172-185 incl. its skip-one quirk
    assert light.estBlockUncomp == 2 << 20 and [u >> 20 for _, u in light.Offsets] == [0, 3, 6, 9]


def test_index_stream_on_reference_framing(oracle):
    """IndexStream over a stream built by the test-side framing restatement (CPU only)."""
    data = synth.make_blocks("text", 1, 5 << 20).numpy()[0].tobytes()
    bs = 256 << 10
    blob = stream_ref.encode_stream(data, 1, bs)
    idx_bytes = mzi.IndexStream(io.BytesIO(blob))
    ix = mzi.Index()
    ix.Load(idx_bytes)
    assert ix.TotalUncompressed == len(data) and ix.TotalCompressed == len(blob)
    # entries sit on chunk starts, one per >= first-block-size of uncompressed data
    pos, upos, starts = 10, 0, {}
    while blob[pos] in (1, 2):
        starts[upos] = pos
        pos += 4 + int.from_bytes(blob[pos + 1:pos + 4], "little")
        upos += bs
    assert all(starts[u] == c for c, u in ix.Offsets)
    assert ix.Offsets[0] == (10, 0) and len(ix.Offsets) == 20
    with pytest.raises(ErrCorrupt):
        mzi.IndexStream(io.BytesIO(blob[10:]))                      # must start with the stream identifier
    with pytest.raises(mzi.ErrUnexpectedEOF):
        mzi.IndexStream(io.BytesIO(blob[:-2]))


# ------------------------------------------------------------------ GPU ----

def _compressible(n, seed):
    rng = np.random.default_rng(seed)
    return (ord("0") + (rng.integers(0, 256, n, dtype=np.uint8) & 3)).astype(np.uint8).tobytes()


@pytest.mark.gpu
def test_example_index_load_skip():
    """index_test.go:30 ExampleIndex_Load: index from CloseIndex, input NOT seekable, Find + Skip."""
    tmp = _compressible(5 << 20, 0xbeef)
    buf = io.BytesIO()
    enc = mzs.NewWriter(buf, mzs.WriterBlockSize(100 << 10))
    enc.EncodeBuffer(tmp)
    idx_bytes = enc.CloseIndex()
    compressed = buf.getvalue()
    assert idx_bytes and compressed[-6:] != b"\x00xdi2s"           # returned, not appended
    for want in range(0, len(tmp), 555555):
        ix = mzi.Index()
        ix.Load(idx_bytes)
        c, u = ix.Find(want)
        dec = mzs.NewReader(io.BytesIO(compressed[c:]), mzs.ReaderIgnoreStreamIdentifier())
        dec.Skip(want - u)
        assert dec.Read() == tmp[want:]


@pytest.mark.gpu
def test_seeking_with_appended_index():
    """index_test.go:119 TestSeeking: ReadSeeker over a stream that carries its index."""
    data = synth.make_blocks("json", 1, 7 << 20).numpy()[0].tobytes() + b"tail!"
    buf = io.BytesIO()
    w = mzs.NewWriter(buf, mzs.WriterBlockSize(64 << 10), mzs.WriterAddIndex(), mzs.WriterLevel(1))
    w.Write(data)
    w.Close()
    blob = buf.getvalue()
    assert blob[-6:] == b"\x00xdi2s"
    assert w.Written() == (len(data), len(blob))
    assert mzs.NewReader(io.BytesIO(blob)).Read() == data           # the index chunk is skippable
    rs = mzs.NewReader(io.BytesIO(blob)).ReadSeeker()
    # writer.go:1084-1088: the compressed total is taken before the index chunk is appended
    idx_len = int.from_bytes(blob[-10:-6], "little")
    assert rs.Index().TotalUncompressed == len(data) and rs.Index().TotalCompressed == len(blob) - idx_len
    rng = np.random.default_rng(3)
    offs = [0, 1, (64 << 10) - 1, 64 << 10, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, len(data) - 1, len(data)]
    offs += [int(x) for x in rng.integers(0, len(data), 12)]
    for off in offs:
        assert rs.Seek(off) == off
        assert rs.Read(1000) == data[off:off + 1000]
    assert rs.Seek(-100, io.SEEK_END) == len(data) - 100 and rs.Read() == data[-100:]
    assert rs.Seek(5, io.SEEK_SET) == 5 and rs.Seek(10, io.SEEK_CUR) == 15 and rs.Read(4) == data[15:19]
    for off, n in ((0, 10), (123456, 300000), (len(data) - 50, 100), (3 << 20, 1 << 20)):
        assert rs.ReadAt(n, off) == data[off:off + n]
    with pytest.raises(mzi.ErrUnexpectedEOF):
        rs.Seek(len(data) + 1)
    with pytest.raises(mzs.ErrCantSeek):
        mzs.NewReader(io.BytesIO(blob[:-20])).ReadSeeker()          # no index at the end
    with pytest.raises(mzs.ErrCantSeek):
        mzs.NewReader(_NoSeek(blob)).ReadSeeker()                  # input cannot seek


class _NoSeek:
    def __init__(self, b):
        self.b = io.BytesIO(b)

    def read(self, n=-1):
        return self.b.read(n)


@pytest.mark.gpu
def test_index_stream_equals_writer_index():
    """index_test.go:273,419: indexing an existing stream gives the index the writer would have made;
    it can be handed to ReadSeeker separately."""
    data = _compressible(3 << 20, 7) + synth.make_blocks("random", 1, 1 << 20).numpy()[0].tobytes()
    buf = io.BytesIO()
    w = mzs.NewWriter(buf, mzs.WriterBlockSize(64 << 10))
    w.EncodeBuffer(data)
    own = w.CloseIndex()
    blob = buf.getvalue()
    made = mzi.IndexStream(io.BytesIO(blob))
    a, b = mzi.Index(), mzi.Index()
    a.Load(own)
    b.Load(made)
    assert (a.TotalUncompressed, a.TotalCompressed) == (b.TotalUncompressed, b.TotalCompressed) == (len(data), len(blob))
    # the writer's first entry is the stream header (0, 0) and it spaces entries >= 1 MiB apart
    # (index.go:59-61); IndexStream starts at the first data chunk (10, 0) and spaces them by the
    # first block's size (index.go:492-495).  Same chunk boundaries either way.
    assert a.Offsets[0] == (0, 0) and b.Offsets[0] == (10, 0)
    assert set(a.Offsets[1:]) <= set(b.Offsets) and len(a.Offsets) == 4 and len(b.Offsets) == 64
    assert [u for _, u in b.Offsets] == [k << 16 for k in range(64)]
    rs = mzs.NewReader(io.BytesIO(blob)).ReadSeeker(made)
    for off in (0, 70000, (3 << 20) - 5, (3 << 20) + 12345, len(data) - 3):
        assert rs.ReadAt(4096, off) == data[off:off + 4096]
