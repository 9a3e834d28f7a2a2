"""Stream layer mirror: Writer / Reader over the batched GPU block codec.

Mirrors the reference's stream API (writer.go NewWriter/Write/ReadFrom/
EncodeBuffer/Flush/Close/Written/Reset and the WriterLevel/WriterBlockSize/
WriterConcurrency/WriterUncompressed options; reader.go NewReader/Read/WriteTo/
DecodeConcurrent/Reset and ReaderMaxBlockSize/ReaderIgnoreCRC/
ReaderIgnoreStreamIdentifier) for the part of the stream format that carries
block data: stream identifier, compressed chunks 0x02 (and 0x03 on read),
uncompressed chunks 0x01, EOF 0x20, padding / skippable chunks on read
(SPEC.md "STREAM FORMAT" sections 1-4.11).  Index, search tables, sidecars,
padding on write and the Snappy/S2 fallback are out of scope (SURVEY 2, rows
13-15, 23) and raise ErrUnsupported where a stream requires them.

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
        """writer.go:1051-1074: flush, then the EOF chunk with the total uncompressed size."""
        if self.closed:
            return
        self.Flush()
        body = _uvarint(self.uncomp_written)
        self._out(bytes([CHUNK_EOF, len(body), 0, 0]) + body)
        self.closed = True

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
            self.out += res
            self.block_start += len(res)

    def _fill(self):
        """Parses chunks until a batch is complete (or the stream ends), then decodes it."""
        batch = []
        try:
            while len(batch) < self.concurrency:
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
                    batch.append((ctype, crc, body, n))
                elif ctype == CHUNK_UNCOMPRESSED:
                    if clen < 4:
                        raise ErrCorrupt()
                    n = clen - 4
                    if n > self.max_block:
                        raise ErrTooLarge()
                    buf = self._read_full(clen, False)
                    batch.append((ctype, int.from_bytes(buf[:4], "little"), buf[4:], n))
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
            self._fill()
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

    def DecodeConcurrent(self, w, concurrent=0):
        if concurrent > 0:
            self.concurrency = concurrent
        return self.WriteTo(w)


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
