"""Loaders for the reference's test corpora (zip files of raw or go-fuzz inputs).

Format follows internal/fuzz/helpers.go:83-161 of the reference: a file that
starts with "go test fuzz" holds one Go `[]byte("...")` literal per line; any
other file is a raw input.
"""
import os
import re
import zipfile

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_SIMPLE = {"a": 7, "b": 8, "f": 12, "n": 10, "r": 13, "t": 9, "v": 11, "\\": 92, "'": 39, '"': 34}


def go_unquote(lit: str) -> bytes:
    """strconv.Unquote for interpreted ("...") and raw (`...`) string literals."""
    if lit[0] == "`":
        return lit[1:-1].replace("\r", "").encode("utf-8")
    assert lit[0] == '"' and lit[-1] == '"', lit[:20]
    s = lit[1:-1]
    out = bytearray()
    i, n = 0, len(s)
    while i < n:
        c = s[i]
        if c != "\\":
            out += c.encode("utf-8")
            i += 1
            continue
        c = s[i + 1]
        if c in _SIMPLE:
            out.append(_SIMPLE[c])
            i += 2
        elif c == "x":
            out.append(int(s[i + 2:i + 4], 16))
            i += 4
        elif c == "u":
            out += chr(int(s[i + 2:i + 6], 16)).encode("utf-8")
            i += 6
        elif c == "U":
            out += chr(int(s[i + 2:i + 10], 16)).encode("utf-8")
            i += 10
        elif c in "01234567":
            out.append(int(s[i + 1:i + 4], 8))
            i += 4
        else:
            raise ValueError("bad escape \\" + c)
    return bytes(out)


_LINE = re.compile(r'^\[\]byte\((.*)\)$', re.S)


def parse_corpus_file(b: bytes):
    if not b.startswith(b"go test fuzz"):
        return [b]
    vals = []
    for line in b.split(b"\n")[1:]:
        line = line.strip()
        if not line:
            continue
        m = _LINE.match(line.decode("utf-8"))
        if not m:
            raise ValueError("malformed corpus line %r" % line[:40])
        vals.append(go_unquote(m.group(1)))
    return vals


def load_zip(path):
    """Yield (name, bytes) for every input in a corpus zip."""
    with zipfile.ZipFile(path) as z:
        for info in z.infolist():
            if info.is_dir():
                continue
            for k, v in enumerate(parse_corpus_file(z.read(info))):
                yield "%s#%d" % (info.filename, k), v


def golden_path(name):
    return os.path.join(GOLDEN, name)
