"""Stream layer mirror: Writer / Reader over the batched GPU block codec.

Mirrors the reference's stream API (writer.go NewWriter/Write/ReadFrom/
EncodeBuffer/Flush/Close/Written/Reset and the WriterLevel/WriterBlockSize/
WriterConcurrency/WriterUncompressed options; reader.go NewReader/Read/WriteTo/
DecodeConcurrent/Reset and ReaderMaxBlockSize/ReaderIgnoreCRC/
ReaderIgnoreStreamIdentifier) for the part of the stream format that carries
block data: stream identifier, compressed chunks 0x02 (and 0x03 on read),
uncompressed chunks 0x01, EOF 0x20, padding / skippable chunks on read
(SPEC.md "STREAM FORMAT" sections 1-4.11), and the seek index (SURVEY 8(f) N3:
index.py; WriterAddIndex / WriterCreateIndex / CloseIndex on write, Skip /
ReadSeeker.Seek / ReadAt on read -- reader.go:1034,1322-1489).  Search tables,
sidecars, padding on write and the Snappy/S2 fallback are out of scope (SURVEY
2, rows 13-15, 23) and raise ErrUnsupported where a stream requires them.

Where the work happens: framing (a few header bytes per chunk) is host code as
in the reference; the per-block work -- CRC-32C of the uncompressed block,
encode / decode -- is one batched GPU call per `concurrency` blocks through
mzcu_stream_encode_blocks / mzcu_stream_decode_blocks.  The reference spawns
one goroutine per block (writer.go:670, reader.go:830-859); a batch is the GPU
equivalent of that fan-out and preserves stream order the same way.
"""
import io

import numpy as np

from . import _lib
from .index import ErrUnexpectedEOF, Index
from . import (ErrCorrupt, ErrInvalidLevel, ErrTooLarge, ErrUnsupported, LevelBalanced, LevelFastest, LevelSuperFast,
               LevelUncompressed, MinLZError, _raise)

# minlz.go:77-131
MAGIC_CHUNK = b"\xff\x06\x00\x00MinLz"
MAGIC_BODY = b"MinLz"
MAGIC_BODY_S2 = b"S2sTwO"
MAGIC_BODY_SNAPPY = b"sNaPpY"
CHUNK_LEGACY = 0x00
CHUNK_UNCOMPRESSED = 0x01
CHUNK_MINLZ = 0x02
CHUNK_MINLZ_COMP_CRC = 0x03
CHUNK_EOF = 0x20
MAX_NON_SKIPPABLE = 0x3f
CHUNK_PADDING = 0xfe
CHUNK_STREAM_ID = 0xff
MIN_BLOCK_SIZE = 4 << 10
MAX_BLOCK_SIZE = 8 << 20
DEFAULT_BLOCK_SIZE = 2 << 20
MAX_CHUNK = (1 << 24) - 1


class ErrCRC(MinLZError):
    """decode.go:33 ErrCRC: checksum mismatch (streams only)."""

    def __init__(self, msg="minlz: corrupt input, crc mismatch"):
        super().__init__(msg)


def _uvarint(x):
    out = bytearray()
    while x >= 0x80:
        out.append((x & 0x7f) | 0x80)
        x >>= 7
    out.append(x)
    return bytes(out)


def _read_uvarint(buf):
    """encoding/binary.Uvarint: (value, n) with n <= 0 on failure."""
    x = 0
    s = 0
    for i, b in enumerate(buf):
        if i == 10:
            return 0, -(i + 1)
        if b < 0x80:
            if i == 9 and b > 1:
                return 0, -(i + 1)
            return x | (b << s), i + 1
        x |= (b & 0x7f) << s
        s += 7
    return 0, 0


def make_header(block_size):
    """writer.go:1553-1556 makeHeader: magic + log2(block size) - 10."""
    return MAGIC_CHUNK + bytes([(block_size - 1).bit_length() - 10])


# ---- options (functional, like the reference's WriterOption / ReaderOption) ----

def WriterLevel(n):
    def f(w):
        if n not in (LevelSuperFast, LevelUncompressed, LevelFastest, LevelBalanced):
            raise ErrInvalidLevel()  # LevelSmallest is not on the accelerated path
        w.level = n
    return f


def WriterUncompressed():
    def f(w):
        w.level = LevelUncompressed
    return f


def WriterBlockSize(n):
    def f(w):
        if n > MAX_BLOCK_SIZE or n < MIN_BLOCK_SIZE:
            raise ValueError("minlz: block size out of bounds. Must be <= 8MB and >=4KB")
        w.block_size = n
    return f


def WriterConcurrency(n):
    """Blocks in flight.  On the GPU this is the batch size of one launch."""
    def f(w):
        if n <= 0:
            raise ValueError("concurrency must be at least 1")
        w.concurrency = n
    return f


def WriterAddIndex(b=True):
    """writer.go:1194 WriterAddIndex: append the seek index to the end of the stream on Close."""
    def f(w):
        if b and not w.gen_index:
            raise ValueError("WriterAddIndex: WriterCreateIndex has been called with false parameter")
        w.append_index = b
    return f


def WriterCreateIndex(b=True):
    """writer.go:1306 WriterCreateIndex: index generation can be disabled (network streams)."""
    def f(w):
        w.gen_index = b
        if not b and w.append_index:
            raise ValueError("WriterCreateIndex: Cannot disable when WriterAddIndex has been requested")
    return f


def WriterDevice(n):
    def f(w):
        w.device = n
    return f


class Writer:
    """writer.go Writer: frames blocks of <= block size into a MinLZ stream."""

    def __init__(self, w, *opts):
        self.level = LevelBalanced  # writer.go:40
        self.block_size = DEFAULT_BLOCK_SIZE
        self.concurrency = 0
        self.device = -1
        self.gen_index = True       # writer.go:41
        self.append_index = False
        for o in opts:
            o(self)
        if self.concurrency == 0:
            self.concurrency = max(1, (256 << 20) // self.block_size)
        self.Reset(w)

    def Reset(self, w):
        self.writer = w
        self.ibuf = bytearray()
        self.wrote_header = False
        self.uncomp_written = 0
        self.written = 0
        self.closed = False
        self.index = Index() if self.gen_index else None   # writer.go:191-202
        if self.index is not None:
            self.index.reset(self.block_size)

    # -- internals ---------------------------------------------------------
    def _out(self, b):
        if self.writer is not None:
            self.writer.write(b)
        self.written += len(b)

    def _encode_blocks(self, data):
        """Frames `data` (bytes-like) cut into blocks; one GPU call per `concurrency` blocks."""
        mv = memoryview(data)
        n = len(mv)
        if n == 0:
            return
        if not self.wrote_header:
            self.wrote_header = True
            if self.index is not None:
                self.index.add(self.written, 0)   # writer.go:241: the header item is indexed too -> entry (0, 0)
            self._out(make_header(self.block_size))
        bs = self.block_size
        per = self.concurrency * bs
        lib = _lib.load()
        for base in range(0, n, per):
            part = np.frombuffer(mv[base:base + per], dtype=np.uint8)
            nblk = (part.size + bs - 1) // bs
            soff = np.minimum(np.arange(nblk + 1, dtype=np.uint64) * bs, part.size).astype(np.uint64)
            crc = np.zeros(nblk, dtype=np.uint32)
            out = bytearray()
            if self.level == LevelUncompressed:
                r = lib.mzcu_crc32c_blocks(self.device, nblk, part.ctypes.data, soff.ctypes.data, crc.ctypes.data)
                if r < 0:
                    _raise(r)
                doff = np.zeros(nblk + 1, dtype=np.uint64)
                comp = None
            else:
                comp = np.empty(part.size + 64, dtype=np.uint8)
                doff = np.zeros(nblk + 1, dtype=np.uint64)
                r = lib.mzcu_stream_encode_blocks(self.device, self.level, nblk, part.ctypes.data, soff.ctypes.data,
                                                  comp.ctypes.data, comp.size, doff.ctypes.data, crc.ctypes.data)
                if r < 0:
                    _raise(r)
            for i in range(nblk):
                a, b = int(soff[i]), int(soff[i + 1])
                c0, c1 = int(doff[i]), int(doff[i + 1])
                if self.index is not None:   # writer.go:241,945: (stream offset of the chunk, uncompressed start)
                    self.index.add(self.written + len(out), self.uncomp_written + a)
                if c1 > c0:  # writer.go:680-696: compressed chunk = crc + uvarint(len) + tokens
                    lenhdr = _uvarint(b - a)
                    clen = 4 + len(lenhdr) + (c1 - c0)
                    out += bytes([CHUNK_MINLZ, clen & 0xff, (clen >> 8) & 0xff, (clen >> 16) & 0xff])
                    out += int(crc[i]).to_bytes(4, "little") + lenhdr
                    out += comp[c0:c1].tobytes()
                else:        # n2 == 0: uncompressed chunk
                    clen = 4 + (b - a)
                    out += bytes([CHUNK_UNCOMPRESSED, clen & 0xff, (clen >> 8) & 0xff, (clen >> 16) & 0xff])
                    out += int(crc[i]).to_bytes(4, "little")
                    out += part[a:b].tobytes()
            self.uncomp_written += part.size
            self._out(bytes(out))

    # -- API -----------------------------------------------------------------
    def Write(self, p):
        """writer.go:276 Write: buffers; full batches of blocks are encoded as they fill."""
        if self.closed:
            raise ValueError("minlz: writer closed")
        self.ibuf += p
        full = (len(self.ibuf) // self.block_size) * self.block_size
        batch = self.concurrency * self.block_size
        if full >= batch:
            take = (full // batch) * batch
            self._encode_blocks(memoryview(self.ibuf)[:take])
            del self.ibuf[:take]
        return len(p)

    write = Write

    def EncodeBuffer(self, buf):
        """writer.go:441 EncodeBuffer: encodes buf directly (flushes pending data first)."""
        self.Flush()
        self._encode_blocks(buf)

    def ReadFrom(self, r):
        """writer.go:316 ReadFrom: streams everything from r; returns bytes read."""
        total = 0
        chunk = self.concurrency * self.block_size
        while True:
            b = r.read(chunk)
            if not b:
                break
            total += len(b)
            self.Write(b)
        return total

    def Flush(self):
        """writer.go:1006 Flush: everything buffered leaves as (possibly short) blocks."""
        if self.ibuf:
            data = bytes(self.ibuf)
            self.ibuf = bytearray()
            self._encode_blocks(data)
        if self.writer is not None and hasattr(self.writer, "flush"):
            self.writer.flush()

    def Close(self):
        """writer.go:1033 Close: flush, EOF chunk, and the index when WriterAddIndex was given."""
        self._close_index(self.append_index)

    def CloseIndex(self):
        """writer.go:1047 CloseIndex: Close and return the index (also appended with WriterAddIndex)."""
        return self._close_index(True)

    def _close_index(self, want):
        """writer.go:1051-1127 closeIndex: flush, the EOF chunk with the total uncompressed size,
        then the index chunk (returned; written only with WriterAddIndex)."""
        if self.closed:
            return None
        if want and self.index is None:
            raise ValueError("index requested, but was asked to not generate one")
        self.Flush()
        body = _uvarint(self.uncomp_written)
        self._out(bytes([CHUNK_EOF, len(body), 0, 0]) + body)
        index = None
        if want:
            index = self.index.appendTo(b"", self.uncomp_written, self.written)
            if self.append_index:
                self._out(index)
        self.closed = True
        return index

    def Written(self):
        """writer.go:1041 Written: (uncompressed in, compressed out)."""
        return self.uncomp_written, self.written

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.Close()


def NewWriter(w, *opts):
    return Writer(w, *opts)


# ---- reader -------------------------------------------------------------------

def ReaderMaxBlockSize(n):
    def f(r):
        if n > MAX_BLOCK_SIZE or n <= 0:
            raise ValueError("minlz: block size too large. Must be <= 8MB and > 0")
        r.max_block = r.max_block_org = n
    return f


def ReaderIgnoreCRC():
    def f(r):
        r.ignore_crc = True
    return f


def ReaderIgnoreStreamIdentifier():
    def f(r):
        r.ignore_stream_id = True
        r.read_header = True
    return f


def ReaderConcurrency(n):
    def f(r):
        r.concurrency = max(1, n)
    return f


def ReaderDevice(n):
    def f(r):
        r.device = n
    return f


class Reader:
    """reader.go Reader: parses chunks on the host, decodes + checksums blocks in GPU batches."""

    def __init__(self, r, *opts):
        self.max_block = self.max_block_org = MAX_BLOCK_SIZE
        self.ignore_crc = False
        self.ignore_stream_id = False
        self.read_header = False
        self.concurrency = 128
        self.device = -1
        for o in opts:
            o(self)
        self.Reset(r)

    def Reset(self, r):
        self.r = r
        self.err = None
        self.out = bytearray()
        self.block_start = 0  # uncompressed bytes delivered in this stream (reader.go blockStart + j)
        self.want_eof = False
        self.read_header = self.ignore_stream_id
        self.max_block = self.max_block_org
        self.done = False
        self.skip_left = 0   # Skip(): uncompressed bytes still to drop before delivering
        self.index = None

    def _read_full(self, n, allow_eof):
        b = self.r.read(n) if n else b""
        while len(b) < n:
            more = self.r.read(n - len(b))
            if not more:
                break
            b += more
        if len(b) < n:
            if len(b) == 0 and allow_eof:
                return None  # clean end of input
            raise ErrCorrupt("minlz: corrupt input (unexpected end of stream)")
        return b

    def _flush_batch(self, batch):
        """batch: list of (kind, crc, payload, dlen).  Appends decoded data to self.out in order."""
        if not batch:
            return
        lib = _lib.load()
        comp = [(i, e) for i, e in enumerate(batch) if e[0] != CHUNK_UNCOMPRESSED]
        raw = [(i, e) for i, e in enumerate(batch) if e[0] == CHUNK_UNCOMPRESSED]
        results = [None] * len(batch)
        if comp:
            n = len(comp)
            src = np.frombuffer(b"".join(e[2] for _, e in comp), dtype=np.uint8)
            soff = np.zeros(n + 1, dtype=np.uint64)
            np.cumsum([len(e[2]) for _, e in comp], out=soff[1:])
            doff = np.zeros(n + 1, dtype=np.uint64)
            np.cumsum([e[3] for _, e in comp], out=doff[1:])
            dst = np.zeros(max(int(doff[-1]), 1), dtype=np.uint8)
            status = np.zeros(n, dtype=np.int32)
            crc = np.zeros(n, dtype=np.uint32)
            r = lib.mzcu_stream_decode_blocks(self.device, n, src.ctypes.data, soff.ctypes.data, dst.ctypes.data,
                                              doff.ctypes.data, status.ctypes.data, crc.ctypes.data)
            if r < 0:
                _raise(r)
            ccrc = None
            if any(e[0] == CHUNK_MINLZ_COMP_CRC for _, e in comp):  # 0x03: crc of the compressed bytes
                ccrc = np.zeros(n, dtype=np.uint32)
                r = lib.mzcu_crc32c_blocks(self.device, n, src.ctypes.data, soff.ctypes.data, ccrc.ctypes.data)
                if r < 0:
                    _raise(r)
            for k, (i, e) in enumerate(comp):
                if status[k] != 0:
                    results[i] = ErrCorrupt()
                    continue
                got = int(ccrc[k]) if e[0] == CHUNK_MINLZ_COMP_CRC else int(crc[k])
                if not self.ignore_crc and got != e[1]:
                    results[i] = ErrCRC()
                    continue
                results[i] = dst[int(doff[k]):int(doff[k + 1])].tobytes()
        if raw:
            n = len(raw)
            if self.ignore_crc:
                for i, e in raw:
                    results[i] = e[2]
            else:
                src = np.frombuffer(b"".join(e[2] for _, e in raw), dtype=np.uint8)
                soff = np.zeros(n + 1, dtype=np.uint64)
                np.cumsum([len(e[2]) for _, e in raw], out=soff[1:])
                crc = np.zeros(n, dtype=np.uint32)
                if src.size:
                    r = lib.mzcu_crc32c_blocks(self.device, n, src.ctypes.data, soff.ctypes.data, crc.ctypes.data)
                    if r < 0:
                        _raise(r)
                else:
                    crc[:] = 0xa282ead8
                for k, (i, e) in enumerate(raw):
                    results[i] = e[2] if int(crc[k]) == e[1] else ErrCRC()
        for res in results:
            if isinstance(res, Exception):
                self.err = res
                return
            self.block_start += len(res)
            if self.skip_left:
                drop = min(self.skip_left, len(res))
                self.skip_left -= drop
                res = res[drop:]
            self.out += res

    def _fill(self, budget=None):
        """Parses chunks until a batch is complete (or the stream ends, or `budget`
        uncompressed bytes are covered), then decodes it.  Blocks that a pending Skip
        covers entirely are dropped without being decoded (reader.go:1107-1150)."""
        batch = []
        covered = 0
        try:
            while len(batch) < self.concurrency and (budget is None or covered < budget or not batch):
                hdr = self._read_full(4, not self.want_eof)
                if hdr is None:
                    self.done = True
                    break
                ctype = hdr[0]
                clen = hdr[1] | hdr[2] << 8 | hdr[3] << 16
                if not self.read_header:  # reader.go:273-284
                    if ctype == CHUNK_STREAM_ID:
                        self.read_header = True
                    elif ctype <= MAX_NON_SKIPPABLE and ctype != CHUNK_EOF:
                        raise ErrCorrupt()
                if ctype in (CHUNK_MINLZ, CHUNK_MINLZ_COMP_CRC):
                    if clen < 4:
                        raise ErrCorrupt()
                    buf = self._read_full(clen, False)
                    crc = int.from_bytes(buf[:4], "little")
                    body = buf[4:]
                    n, hl = _read_uvarint(body)
                    if hl <= 0 or n > 0xffffffff:
                        raise ErrCorrupt()
                    if n > self.max_block:
                        raise ErrTooLarge()
                    body = body[hl:]
                    if n == 0 or n < len(body):  # reader.go:327-333
                        raise ErrCorrupt()
                    if not batch and self.skip_left >= n:
                        self.skip_left -= n
                        self.block_start += n
                        continue
                    batch.append((ctype, crc, body, n))
                    covered += n
                elif ctype == CHUNK_UNCOMPRESSED:
                    if clen < 4:
                        raise ErrCorrupt()
                    n = clen - 4
                    if n > self.max_block:
                        raise ErrTooLarge()
                    buf = self._read_full(clen, False)
                    if not batch and self.skip_left >= n:
                        self.skip_left -= n
                        self.block_start += n
                        continue
                    batch.append((ctype, int.from_bytes(buf[:4], "little"), buf[4:], n))
                    covered += n
                elif ctype == CHUNK_LEGACY:
                    raise ErrUnsupported()  # Snappy/S2 fallback stays in host Go (reader.go:355-404)
                elif ctype == CHUNK_EOF:
                    if clen > 10:
                        raise ErrCorrupt()
                    # the size check needs everything before it decoded
                    self._flush_batch(batch)
                    batch = []
                    if self.err is not None:
                        return
                    if clen != 0:
                        buf = self._read_full(clen, False)
                        if not self.ignore_stream_id:
                            want, n = _read_uvarint(buf)
                            if n != clen or want != self.block_start:
                                raise ErrCorrupt()
                    self.want_eof = False
                    self.read_header = self.ignore_stream_id
                elif ctype == CHUNK_STREAM_ID:
                    if clen != len(MAGIC_BODY) + 1:
                        raise ErrCorrupt()
                    body = self._read_full(clen, False)
                    self._flush_batch(batch)
                    batch = []
                    if self.err is not None:
                        return
                    self.block_start = 0
                    if body[:5] == MAGIC_BODY:
                        # reader.go:994-1030 minLzHeader: block size indicator
                        bs = body[5]
                        if bs & 0xc0 or (bs & 15) > 13:
                            raise ErrCorrupt()
                        blk = 1 << ((bs & 15) + 10)
                        if blk > self.max_block_org:
                            raise ErrTooLarge()
                        self.max_block = blk
                        self.want_eof = True
                    elif body in (MAGIC_BODY_S2, MAGIC_BODY_SNAPPY):
                        raise ErrUnsupported()
                    else:
                        raise ErrUnsupported()
                elif ctype <= MAX_NON_SKIPPABLE:
                    raise ErrUnsupported()  # reserved unskippable chunk (SPEC 4.8)
                else:
                    self._read_full(clen, False)  # padding / skippable chunks (SPEC 4.7, 4.9-4.10)
        except MinLZError as e:
            # data before the failure is still delivered first, as in the reference
            self._flush_batch(batch)
            if self.err is None:
                self.err = e
            return
        self._flush_batch(batch)

    # -- API -----------------------------------------------------------------
    def Read(self, n=-1):
        """reader.go:248 Read.  Returns up to n bytes; for n < 0 everything that is left.
        A stream error is raised once the data decoded before it has been delivered
        (for n < 0 the partial data travels on the exception as `.partial`)."""
        while (n < 0 or len(self.out) < n) and self.err is None and not self.done:
            self._fill(None if n < 0 else n - len(self.out))
        if n < 0:
            data = bytes(self.out)
            self.out = bytearray()
            if self.err is not None:
                self.err.partial = data
                raise self.err
            return data
        if not self.out:
            if self.err is not None:
                raise self.err
            return b""
        data = bytes(self.out[:n])
        del self.out[:n]
        return data

    read = Read

    def WriteTo(self, w):
        """reader.go:548 WriteTo / :575 DecodeConcurrent: decode everything into w."""
        total = 0
        while True:
            if not self.out and not self.done and self.err is None:
                self._fill()
            if self.out:
                total += len(self.out)
                w.write(bytes(self.out))
                self.out = bytearray()
                continue
            if self.err is not None:
                raise self.err
            if self.done:
                return total

    def Skip(self, n):
        """reader.go:1034 Skip: drop the next n uncompressed bytes.  Blocks lying entirely
        inside the skipped range are not decoded (and, as in the reference, not checked)."""
        if n < 0:
            raise ValueError("attempted negative skip")
        if self.err is not None:
            raise self.err
        take = min(n, len(self.out))
        del self.out[:take]
        n -= take
        if n == 0:
            return
        self.skip_left += n
        while self.skip_left and self.err is None and not self.done:
            self._fill(1)
        if self.err is not None:
            raise self.err
        if self.skip_left:
            self.skip_left = 0
            self.err = ErrUnexpectedEOF()
            raise self.err

    def ReadSeeker(self, index=None):
        """reader.go:1322 ReadSeeker: random access through a seek index -- the bytes given
        here, else the index at the end of the (seekable) input."""
        if index:
            self.index = Index()
            try:
                self.index.Load(index)
            except MinLZError as e:
                raise ErrCantSeek("loading index returned: %s" % e)
        if not (hasattr(self.r, "seek") and hasattr(self.r, "tell")):
            raise ErrCantSeek("input stream isn't seekable")
        if self.index is None:
            pos = self.r.tell()
            idx = Index()
            try:
                idx.LoadStream(self.r)
            except ErrUnsupported:
                raise ErrCantSeek("input stream does not contain an index")
            except MinLZError as e:
                raise ErrCantSeek("reading index returned: %s" % e)
            finally:
                self.r.seek(pos)
            self.index = idx
        return ReadSeeker(self)

    def DecodeConcurrent(self, w, concurrent=0):
        if concurrent > 0:
            self.concurrency = concurrent
        return self.WriteTo(w)


class ErrCantSeek(MinLZError):
    """reader.go ErrCantSeek{Reason}"""

    def __init__(self, reason):
        super().__init__("minlz: Can't seek because " + reason)


class ReadSeeker:
    """reader.go:1363-1491 ReadSeeker: Seek / ReadAt over an indexed stream.  A read after a seek
    frames only the chunks that cover the requested range and decodes those blocks in one
    batched GPU call (index -> pick blocks -> batch decode)."""

    def __init__(self, reader):
        self.Reader = reader
        self.seek_src = reader.r

    def Index(self):
        return self.Reader.index

    def Read(self, n=-1):
        return self.Reader.Read(n)

    read = Read

    def _pos(self):
        r = self.Reader
        return r.block_start - len(r.out)

    def Seek(self, offset, whence=io.SEEK_SET):
        r = self.Reader
        if r.err is not None:
            raise r.err
        if whence == io.SEEK_SET:
            absolute = offset
        elif whence == io.SEEK_CUR:
            absolute = self._pos() + offset
        elif whence == io.SEEK_END:
            absolute = r.index.TotalUncompressed + offset
        else:
            raise ErrUnsupported()
        if absolute < 0:
            raise ValueError("seek before start of file")
        # inside what is already decoded: no need to seek (reader.go:1417-1422)
        lo = r.block_start - len(r.out)
        if r.skip_left == 0 and lo <= absolute < r.block_start:
            del r.out[:absolute - lo]
            return absolute
        c, u = r.index.Find(absolute)
        self.seek_src.seek(c)
        r.out = bytearray()
        r.block_start = u
        r.skip_left = 0
        r.done = False
        r.want_eof = False
        # chunks are self-delimiting, so parsing may start at any indexed chunk; the stream
        # identifier is only found (and then honoured) when the entry is offset 0
        r.read_header = True
        if absolute > u:
            r.Skip(absolute - u)
        return absolute

    seek = Seek

    def ReadAt(self, n, offset):
        """reader.go:1469 ReadAt: n bytes at uncompressed offset; short only at the end of the stream."""
        self.Seek(offset, io.SEEK_SET)
        out = bytearray()
        while len(out) < n:
            b = self.Reader.Read(n - len(out))
            if not b:
                break
            out += b
        return bytes(out)


def NewReader(r, *opts):
    return Reader(r, *opts)


def encode_stream(data, level=LevelBalanced, block_size=DEFAULT_BLOCK_SIZE):
    """Convenience: bytes -> stream bytes."""
    buf = io.BytesIO()
    w = Writer(buf, WriterLevel(level), WriterBlockSize(block_size))
    w.EncodeBuffer(data)
    w.Close()
    return buf.getvalue()


def decode_stream(blob, *opts):
    return Reader(io.BytesIO(blob), *opts).Read()
