"""Spec-level MinLZ block decoder in pure Python.

Independent second restatement, written from SPEC.md sections 1-2 in the order of
the reference's teaching decoder (internal/reference/decoder.go:26-373).  Used
only to cross-check the C oracle's accept/reject decisions and output on small
inputs.  Returns bytes, or None when the block is invalid.
"""
MAX_BLOCK = 8 << 20


def decode_block(src: bytes):
    if len(src) == 0 or src[0] != 0:
        return None
    if len(src) == 1:
        return b""
    p = 1
    want = 0
    shift = 0
    while True:
        if shift == 70 or p >= len(src):
            return None
        v = src[p]
        p += 1
        want |= (v & 0x7F) << shift
        if want > MAX_BLOCK:
            return None
        shift += 7
        if not v & 0x80:
            break
    rest = src[p:]
    if want == 0:
        return bytes(rest)
    if want < len(rest):
        return None
    src = rest
    n = len(src)
    p = 0
    dst = bytearray()
    offset = 1

    def fits(k):
        return k < MAX_BLOCK and len(dst) + k <= want

    while p < n:
        v = src[p]
        p += 1
        tag, value = v & 3, v >> 2
        if tag == 0:
            rep = value & 1
            value >>= 1
            if value < 29:
                length = value + 1
            else:
                k = value - 28
                if p + k > n:
                    return None
                length = int.from_bytes(src[p:p + k], "little") + 30
                p += k
            if not rep:
                if not fits(length) or p + length > n:
                    return None
                dst += src[p:p + length]
                p += length
                continue
        elif tag == 1:
            length = value & 15
            if p + 1 > n:
                return None
            offset = ((src[p] << 2) | (value >> 4)) + 1
            p += 1
            if length == 15:
                if p + 1 > n:
                    return None
                length = src[p] + 18
                p += 1
            else:
                length += 4
        elif tag == 2:
            if p + 2 > n:
                return None
            offset = int.from_bytes(src[p:p + 2], "little") + 64
            p += 2
            if value <= 60:
                length = value + 4
            else:
                k = value - 60
                if p + k > n:
                    return None
                length = int.from_bytes(src[p:p + k], "little") + 64
                p += k
        else:
            lit_len = (value >> 1) & 3
            if not value & 1:
                if p + 2 > n:
                    return None
                offset = int.from_bytes(src[p:p + 2], "little") + 64
                p += 2
                length = (value >> 3) + 4
                lit_len += 1
            else:
                if p + 3 > n:
                    return None
                value |= int.from_bytes(src[p:p + 3], "little") << 6
                p += 3
                offset = (value >> 9) + 65536
                value = (value >> 3) & 63
                if value < 61:
                    length = value + 4
                else:
                    k = value - 60
                    if p + k > n:
                        return None
                    length = int.from_bytes(src[p:p + k], "little") + 64
                    p += k
            if lit_len:
                if p + lit_len > n or not fits(lit_len):
                    return None
                dst += src[p:p + lit_len]
                p += lit_len
        if not fits(length) or offset > len(dst):
            return None
        pos = len(dst) - offset
        if offset >= length:
            dst += dst[pos:pos + length]
        else:
            for i in range(length):
                dst.append(dst[pos + i])
    if len(dst) != want:
        return None
    return bytes(dst)
