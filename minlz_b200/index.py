"""Seek index of a MinLZ stream: mirror of the reference's
index.go -- Index.add / Find / reduce / appendTo / Load / LoadStream, IndexStream,
RemoveIndexHeaders / RestoreIndexHeaders -- with the same names, argument meaning
and error behaviour.

The index is host-side integer bookkeeping (pairs of compressed / uncompressed
stream offsets, zig-zag varints); nothing here touches block bytes.  It is what
lets the GPU decoder serve random-access reads: ReadSeeker (stream.py) asks
Find() for the chunk that holds an offset, frames the chunks covering the
requested range and decodes exactly those blocks in one batched GPU call.
"""
import io

from . import ErrCorrupt, ErrUnsupported, MinLZError

INDEX_HEADER = b"s2idx\x00"          # index.go:27
INDEX_TRAILER = b"\x00xdi2s"         # index.go:28
MAX_INDEX_ENTRIES = 1 << 16          # index.go:29
MIN_INDEX_DIST = 1 << 20             # index.go:30
CHUNK_TYPE_INDEX = 0x40              # minlz.go:125
LEGACY_INDEX_CHUNK = 0x99            # minlz.go:130
SKIPPABLE_FRAME_HEADER = 4           # minlz.go:103
MAX_USER_CHUNK_SIZE = (1 << 24) - 1  # minlz.go MaxUserChunkSize


class ErrUnexpectedEOF(MinLZError):
    """io.ErrUnexpectedEOF"""

    def __init__(self, msg="unexpected EOF"):
        super().__init__(msg)


def put_varint(x):
    """encoding/binary.PutVarint: zig-zag, then base-128 little endian."""
    ux = (x << 1) ^ (x >> 63)
    ux &= (1 << 64) - 1
    out = bytearray()
    while ux >= 0x80:
        out.append((ux & 0x7F) | 0x80)
        ux >>= 7
    out.append(ux)
    return bytes(out)


def varint(b, pos=0):
    """encoding/binary.Varint on b[pos:]: (value, bytes read); n <= 0 on error like Go."""
    ux, sh, i = 0, 0, pos
    while True:
        if i >= len(b):
            return 0, 0
        c = b[i]
        i += 1
        if c < 0x80:
            if i - pos > 10 or (i - pos == 10 and c > 1):
                return 0, -(i - pos)
            ux |= c << sh
            break
        ux |= (c & 0x7F) << sh
        sh += 7
        if i - pos >= 10:
            return 0, -(i - pos)
    x = ux >> 1
    if ux & 1:
        x = ~x
    return x, i - pos


class Index:
    """index.go:33-49.  Offsets: list of (CompressedOffset, UncompressedOffset), sorted."""

    def __init__(self):
        self.TotalUncompressed = -1
        self.TotalCompressed = -1
        self.Offsets = []
        self.estBlockUncomp = 0

    # index.go:55-68
    def reset(self, max_block):
        while max_block < MIN_INDEX_DIST:
            max_block *= 2
        self.estBlockUncomp = max_block
        self.TotalCompressed = -1
        self.TotalUncompressed = -1
        self.Offsets = []

    # index.go:80-105
    def add(self, compressed_offset, uncompressed_offset):
        if self.Offsets:
            lc, lu = self.Offsets[-1]
            if uncompressed_offset - lu < self.estBlockUncomp:
                return  # don't add until we have estBlockUncomp
            if lu > uncompressed_offset:
                raise MinLZError("internal error: Earlier uncompressed received (%d > %d)" % (lu, uncompressed_offset))
            if lc > compressed_offset:
                raise MinLZError("internal error: Earlier compressed received (%d > %d)" % (lu, uncompressed_offset))
        self.Offsets.append((compressed_offset, uncompressed_offset))
        if len(self.Offsets) > MAX_INDEX_ENTRIES:
            self.reduceLight()

    # index.go:114-144
    def Find(self, offset):
        """(compressedOff, uncompressedOff) of the entry at or before the uncompressed offset."""
        if self.TotalUncompressed < 0:
            raise ErrCorrupt()
        if offset < 0:
            offset = self.TotalUncompressed + offset
            if offset < 0:
                raise ErrUnexpectedEOF()
        if offset > self.TotalUncompressed:
            raise ErrUnexpectedEOF()
        c = u = 0
        if len(self.Offsets) > 200:
            lo, hi = 0, len(self.Offsets)
            while lo < hi:  # sort.Search: first n with UncompressedOffset > offset
                mid = (lo + hi) // 2
                if self.Offsets[mid][1] > offset:
                    hi = mid
                else:
                    lo = mid + 1
            n = max(lo, 1)
            return self.Offsets[n - 1]
        for co, uo in self.Offsets:
            if uo > offset:
                break
            c, u = co, uo
        return c, u

    # index.go:147-169
    def reduce(self):
        if len(self.Offsets) < MAX_INDEX_ENTRIES:
            return
        remove_n = (len(self.Offsets) + 1) // MAX_INDEX_ENTRIES
        src = self.Offsets
        while self.estBlockUncomp * (remove_n + 1) < MIN_INDEX_DIST and len(src) // (remove_n + 1) > 1000:
            remove_n += 1
        self.Offsets = src[::remove_n + 1]
        self.estBlockUncomp += self.estBlockUncomp * remove_n

    # index.go:172-185
    def reduceLight(self):
        self.estBlockUncomp *= 2
        src, out, idx = self.Offsets, [], 0
        while idx < len(src):
            base = src[idx]
            out.append(base)
            while idx < len(src) and src[idx][1] - base[1] < self.estBlockUncomp:
                idx += 1
            idx += 1  # the for loop's own idx++
        self.Offsets = out

    # index.go:187-270
    def appendTo(self, b, uncomp_total, comp_total):
        self.reduce()
        out = bytearray(b)
        init = len(out)
        out += bytes([CHUNK_TYPE_INDEX, 0, 0, 0]) + INDEX_HEADER
        out += put_varint(uncomp_total) + put_varint(comp_total) + put_varint(self.estBlockUncomp)
        out += put_varint(len(self.Offsets))
        has_uncompressed = 0
        for idx, (_, uo) in enumerate(self.Offsets):
            if idx == 0:
                if uo != 0:
                    has_uncompressed = 1
                    break
                continue
            if uo != self.Offsets[idx - 1][1] + self.estBlockUncomp:
                has_uncompressed = 1
                break
        out.append(has_uncompressed)
        if has_uncompressed:
            for idx, (_, uo) in enumerate(self.Offsets):
                if idx > 0:
                    uo -= self.Offsets[idx - 1][1] + self.estBlockUncomp
                out += put_varint(uo)
        c_predict = self.estBlockUncomp // 2
        for idx, (co, _) in enumerate(self.Offsets):
            if idx > 0:
                co -= self.Offsets[idx - 1][0] + c_predict
                c_predict += _go_div2(co)  # half the error
            out += put_varint(co)
        out += (len(out) - init + 4 + len(INDEX_TRAILER)).to_bytes(4, "little")
        out += INDEX_TRAILER
        chunk_len = len(out) - init - SKIPPABLE_FRAME_HEADER
        out[init + 1:init + 4] = chunk_len.to_bytes(3, "little")
        return bytes(out)

    # index.go:273-410
    def Load(self, b):
        """Loads a binary index; returns the bytes after it.  Errors as in Go."""
        b = bytes(b)
        if len(b) <= 4 + len(INDEX_HEADER) + len(INDEX_TRAILER):
            raise ErrUnexpectedEOF()
        if b[0] != CHUNK_TYPE_INDEX and b[0] != LEGACY_INDEX_CHUNK:
            raise ErrCorrupt()
        chunk_len = int.from_bytes(b[1:4], "little")
        p = 4
        if len(b) - p < chunk_len:
            raise ErrUnexpectedEOF()
        if b[p:p + len(INDEX_HEADER)] != INDEX_HEADER:
            raise ErrUnsupported()
        p += len(INDEX_HEADER)

        def rd(nonneg):
            nonlocal p
            v, n = varint(b, p)
            if n <= 0 or (nonneg and v < 0):
                raise ErrCorrupt()
            p += n
            return v

        self.TotalUncompressed = rd(True)
        self.TotalCompressed = rd(False)
        self.estBlockUncomp = rd(True)
        entries = rd(True)
        if entries > MAX_INDEX_ENTRIES:
            raise ErrCorrupt()
        if len(b) - p < 1:
            raise ErrUnexpectedEOF()
        has_uncompressed = b[p]
        p += 1
        if has_uncompressed & 1 != has_uncompressed:
            raise ErrCorrupt()
        uoffs = []
        for idx in range(entries):
            uo = rd(False) if has_uncompressed else 0
            if idx > 0:
                prev = uoffs[idx - 1]
                uo += prev + self.estBlockUncomp
                if uo <= prev:
                    raise ErrCorrupt()
            if uo < 0:
                raise ErrCorrupt()
            uoffs.append(uo)
        c_predict = self.estBlockUncomp // 2
        coffs = []
        for idx in range(entries):
            co = rd(False)
            if idx > 0:
                c_new = c_predict + _go_div2(co)
                prev = coffs[idx - 1]
                co += prev + c_predict
                if co <= prev:
                    raise ErrCorrupt()
                c_predict = c_new
            if co < 0:
                raise ErrCorrupt()
            coffs.append(co)
        self.Offsets = list(zip(coffs, uoffs))
        if len(b) - p < 4 + len(INDEX_TRAILER):
            raise ErrUnexpectedEOF()
        p += 4
        if b[p:p + len(INDEX_TRAILER)] != INDEX_TRAILER:
            raise ErrCorrupt()
        return b[p + len(INDEX_TRAILER):]

    # index.go:416-449
    def LoadStream(self, rs):
        """Loads the index from the end of a seekable stream."""
        rs.seek(-10, io.SEEK_END)
        tmp = rs.read(10)
        if len(tmp) != 10:
            raise ErrUnexpectedEOF()
        if tmp[4:4 + len(INDEX_TRAILER)] != INDEX_TRAILER:
            raise ErrUnsupported()
        sz = int.from_bytes(tmp[:4], "little")
        if sz > MAX_USER_CHUNK_SIZE + SKIPPABLE_FRAME_HEADER:
            raise ErrCorrupt()
        rs.seek(-sz, io.SEEK_END)
        buf = rs.read(sz)
        if len(buf) != sz:
            raise ErrUnexpectedEOF()
        self.Load(buf)

    # index.go:553-578
    def JSON(self):
        import json
        return json.dumps({
            "total_uncompressed": self.TotalUncompressed, "total_compressed": self.TotalCompressed,
            "offsets": [{"compressed": c, "uncompressed": u} for c, u in self.Offsets],
            "est_block_uncompressed": self.estBlockUncomp}, indent=2).encode()


def _go_div2(x):
    """Go's x / 2 on int64 truncates toward zero (Python's // floors)."""
    return -((-x) // 2) if x < 0 else x // 2


def IndexStream(r):
    """index.go:455-550: index of an existing stream (structure checked, block data not)."""
    from . import DecodedLen
    from . import stream as st
    idx = Index()
    idx.TotalCompressed = 0
    idx.TotalUncompressed = 0
    read_header = False
    while True:
        hdr = r.read(4)
        if len(hdr) == 0:
            return idx.appendTo(b"", idx.TotalUncompressed, idx.TotalCompressed)
        if len(hdr) < 4:
            raise ErrUnexpectedEOF()
        start_chunk = idx.TotalCompressed
        idx.TotalCompressed += 4
        ctype = hdr[0]
        if not read_header:
            if ctype != st.CHUNK_STREAM_ID and ctype != st.CHUNK_EOF:
                raise ErrCorrupt()
            read_header = True
        chunk_len = int.from_bytes(hdr[1:4], "little")
        if chunk_len < 4:
            raise ErrCorrupt()
        idx.TotalCompressed += chunk_len
        buf = r.read(chunk_len)
        if len(buf) != chunk_len:
            raise ErrUnexpectedEOF()
        if ctype in (st.CHUNK_LEGACY, st.CHUNK_MINLZ, st.CHUNK_MINLZ_COMP_CRC):
            # stream chunks carry the block without its leading 0x00 (writer.go:677)
            body = buf[4:]
            # index.go:493 calls DecodedLen on the chunk body, which starts with the uvarint and so
            # takes the plain-varint branch of isMinLZ (decode.go:128-133): no block-header checks
            dlen = _snappy_len(body)
            if dlen > st.MAX_BLOCK_SIZE:
                raise ErrCorrupt()
            n2 = dlen
        elif ctype == st.CHUNK_UNCOMPRESSED:
            n2 = chunk_len - 4
            if n2 > st.MAX_BLOCK_SIZE:
                raise ErrCorrupt()
        elif ctype == st.CHUNK_STREAM_ID:
            if chunk_len != 6 or (buf[:5] != st.MAGIC_BODY and buf != st.MAGIC_BODY_S2 and buf != st.MAGIC_BODY_SNAPPY):
                raise ErrCorrupt()
            continue
        elif ctype == st.CHUNK_EOF:
            continue
        elif ctype <= st.MAX_NON_SKIPPABLE:
            raise ErrUnsupported()
        else:
            continue  # user chunks and padding
        if idx.estBlockUncomp == 0:
            idx.estBlockUncomp = n2
        idx.add(start_chunk, idx.TotalUncompressed)
        idx.TotalUncompressed += n2


def _snappy_len(body):
    v, sh = 0, 0
    for i, c in enumerate(body[:10]):
        v |= (c & 0x7F) << sh
        if c < 0x80:
            if v > 0xFFFFFFFF:
                raise ErrCorrupt()
            return v
        sh += 7
    raise ErrCorrupt()


# index.go:581-613
def RemoveIndexHeaders(b):
    b = bytes(b)
    save = 4 + len(INDEX_HEADER) + len(INDEX_TRAILER) + 4
    if len(b) <= save or b[0] != CHUNK_TYPE_INDEX:
        return None
    chunk_len = int.from_bytes(b[1:4], "little")
    b = b[4:]
    if len(b) < chunk_len:
        return None
    b = b[:chunk_len]
    if b[:len(INDEX_HEADER)] != INDEX_HEADER:
        return None
    b = b[len(INDEX_HEADER):]
    if not b.endswith(INDEX_TRAILER):
        return None
    b = b[:-len(INDEX_TRAILER)]
    if len(b) < 4:
        return None
    return b[:-4]


# index.go:616-636
def RestoreIndexHeaders(inp):
    inp = bytes(inp)
    if len(inp) == 0:
        return inp
    b = bytearray([CHUNK_TYPE_INDEX, 0, 0, 0]) + INDEX_HEADER + inp
    b += (len(b) + 4 + len(INDEX_TRAILER)).to_bytes(4, "little")
    b += INDEX_TRAILER
    chunk_len = len(b) - SKIPPABLE_FRAME_HEADER
    b[1:4] = chunk_len.to_bytes(3, "little")
    return bytes(b)
