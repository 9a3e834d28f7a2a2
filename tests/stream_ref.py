"""Test-side restatement of the MinLZ stream framing (SPEC.md "STREAM FORMAT",
writer.go:604-744,1051-1074, reader.go:248-545) on top of the CPU oracle.
Used to pin the byte-exact output of minlz_b200.stream.Writer and to make
streams for the Reader tests."""
from oracle import binding as oracle


def uvarint(x):
    out = bytearray()
    while x >= 0x80:
        out.append((x & 0x7f) | 0x80)
        x >>= 7
    out.append(x)
    return bytes(out)


def header(block_size):
    return b"\xff\x06\x00\x00MinLz" + bytes([(block_size - 1).bit_length() - 10])


def chunk(ctype, body):
    n = len(body)
    return bytes([ctype, n & 0xff, (n >> 8) & 0xff, (n >> 16) & 0xff]) + body


def data_chunk(block, level):
    crc = oracle.crc(block).to_bytes(4, "little")
    tok = oracle.encode_block(block, level) if level in (-1, 1, 2) else b""
    if tok:
        return chunk(0x02, crc + uvarint(len(block)) + tok)
    return chunk(0x01, crc + block)


def encode_stream(data, level, block_size):
    out = bytearray()
    if data:
        out += header(block_size)
    for a in range(0, len(data), block_size):
        out += data_chunk(data[a:a + block_size], level)
    out += chunk(0x20, uvarint(len(data)))
    return bytes(out)
