"""Deterministic synthetic block generators for the MinLZ benchmarks and tests.

Bench / test infrastructure, not part of the product.  Everything is built from
a counter-based integer hash (no torch RNG), so the same (kind, seed, block
index) gives the same bytes on CPU and on CUDA.

Kinds:
  json    newline-delimited JSON-like records over a 4096-word skewed vocabulary
  log     Apache-combined-style access-log lines
  text    vocabulary prose
  binary  packed structs (u64 counter, f32-like noise, u16 enum, 6 zero bytes)
  random  uniform random bytes (incompressible; the encoder must return 0)
"""
import torch

_M64 = (1 << 64) - 1


def _i64(x):
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


_C1, _C2 = _i64(0xbf58476d1ce4e5b9), _i64(0x94d049bb133111eb)


def _lsr(x, k):
    return (x >> k) & ((1 << (64 - k)) - 1)


def mix64(x):
    """splitmix64 finaliser on int64 tensors (wrapping arithmetic)."""
    x = (x ^ _lsr(x, 30)) * _C1
    x = (x ^ _lsr(x, 27)) * _C2
    return x ^ _lsr(x, 31)


def rand_u(shape_idx, stream, seed):
    """Uniform int64 in [0, 2^31) per element of the index tensor."""
    key = _i64(seed * 0x9e3779b97f4a7c15 + stream * 0xd1b54a32d192ed03)
    return _lsr(mix64(shape_idx * _i64(0x2545f4914f6cdd1d) + key), 33)


# ---- piece table -----------------------------------------------------------
_LETTERS = "etaoinshrdlcumwfgypbvkjxqz"


def _make_words(n, seed):
    words = []
    x = seed
    for i in range(n):
        x = (x * 6364136223846793005 + 1442695040888963407) & _M64
        ln = 3 + ((x >> 59) % 8)
        w = []
        y = x
        for _ in range(ln):
            y = (y * 6364136223846793005 + 1442695040888963407) & _M64
            # skew towards frequent letters
            k = (y >> 40) % 676
            w.append(_LETTERS[min(k % 26, k // 26)])
        words.append("".join(w))
    return words


class PieceTable:
    """Byte matrix of all string pieces a generator can emit."""

    def __init__(self, pieces, device):
        self.index = {}
        maxlen = max(len(p) for p in pieces)
        tab = torch.zeros((len(pieces), maxlen), dtype=torch.uint8)
        lens = torch.zeros(len(pieces), dtype=torch.int32)
        for i, p in enumerate(pieces):
            b = p.encode("latin-1") if isinstance(p, str) else p
            tab[i, :len(b)] = torch.tensor(list(b), dtype=torch.uint8)
            lens[i] = len(b)
        self.tab = tab.to(device)
        self.lens = lens.to(device)


_TABLE_CACHE = {}


def _tables(device):
    key = str(device)
    if key in _TABLE_CACHE:
        return _TABLE_CACHE[key]
    words = _make_words(4096, 12345)
    nums = ["%04d" % i for i in range(10000)]
    fixed = ['{"id":', ',"user":"', '","ts":"2026-01-', 'T', ':', 'Z","status":"', '","tags":[', '"', '",', '],"latency_ms":',
             ',"msg":"', ' ', '"}\n', '', '",'
             ]
    status = ["ok", "error", "timeout", "retry", "denied", "created", "accepted", "not_found"]
    d2 = ["%02d" % i for i in range(60)]
    pieces = words + nums + fixed + status + d2
    t = PieceTable(pieces, device)
    base = {"word": 0, "num": 4096, "fixed": 4096 + 10000, "status": 4096 + 10000 + len(fixed),
            "d2": 4096 + 10000 + len(fixed) + len(status)}
    # log pieces
    paths = ["/" + "/".join(words[(i * 7 + k * 13) % 4096] for k in range(1 + i % 3)) for i in range(8192)]
    uas = ["Mozilla/5.0 (%s; rv:%d.0) %s/%d.%d" % (words[(i * 3) % 4096], 50 + i % 70, words[(i * 5 + 1) % 4096], i % 20, i % 9)
           for i in range(256)]
    octets = [str(i) for i in range(256)]
    lfixed = [".", " - - [01/Jan/2026:", ":", " +0000] \"GET ", " HTTP/1.1\" ", " ", " \"-\" \"", "\"\n", "200", "404", "304", "500",
              "301", "POST", ""]
    lt = PieceTable(paths + uas + octets + lfixed + d2 + nums, device)
    lbase = {"path": 0, "ua": 8192, "octet": 8192 + 256, "fixed": 8192 + 512, "d2": 8192 + 512 + len(lfixed),
             "num": 8192 + 512 + len(lfixed) + 60}
    _TABLE_CACHE[key] = (t, base, lt, lbase)
    return _TABLE_CACHE[key]


def _skewed(u31, n):
    """Zipf-like index in [0, n) from a uniform 31-bit integer: n * u^3."""
    f = u31.to(torch.float64) / float(1 << 31)
    f = f * f
    return (f * f * n).to(torch.int64).clamp_(0, n - 1)


def _assemble(pid, table, block_size):
    """pid: [B, P] int64 piece ids -> [B, block_size] uint8 (concatenated, cut)."""
    lens = table.lens[pid].to(torch.int64)
    cum = torch.cumsum(lens, dim=1)
    assert int(cum[:, -1].min()) >= block_size, "not enough pieces for the block size"
    B = pid.shape[0]
    j = torch.arange(block_size, device=pid.device, dtype=torch.int64).unsqueeze(0).expand(B, -1).contiguous()
    k = torch.searchsorted(cum, j, right=True)
    start = torch.gather(cum, 1, k) - torch.gather(lens, 1, k)
    p = torch.gather(pid, 1, k)
    return table.tab[p, j - start]


def _json_blocks(first, count, block_size, seed, device):
    t, b, _, _ = _tables(device)
    R = block_size // 100 + 8  # records per block (every record is > 100 bytes)
    P = 48
    F = b["fixed"]
    blk = torch.arange(first, first + count, device=device, dtype=torch.int64).view(-1, 1, 1)
    rec = torch.arange(R, device=device, dtype=torch.int64).view(1, -1, 1)
    slot = torch.arange(P, device=device, dtype=torch.int64).view(1, 1, -1)
    idx = (blk * R + rec) * P + slot
    u = rand_u(idx, 1, seed)
    word = b["word"] + _skewed(u, 4096)
    num = b["num"] + (u % 10000)
    d2 = b["d2"] + (u % 60)
    pid = torch.full((count, R, P), F + 13, dtype=torch.int64, device=device)  # empty piece
    rid = (blk * R + rec).expand(count, R, 1)

    def put(s, v):
        pid[:, :, s] = v if isinstance(v, int) else v.reshape(count, R)

    put(0, F + 0)
    put(1, b["num"] + (rid[:, :, 0] // 10000) % 10000)
    put(2, b["num"] + rid[:, :, 0] % 10000)
    put(3, F + 1)
    put(4, word[:, :, 4])
    put(5, F + 2)
    put(6, b["d2"] + 1 + (u[:, :, 6] % 28))
    put(7, F + 3)
    put(8, b["d2"] + (rid[:, :, 0] // 3600) % 24)
    put(9, F + 4)
    put(10, b["d2"] + (rid[:, :, 0] // 60) % 60)
    put(11, F + 4)
    put(12, b["d2"] + rid[:, :, 0] % 60)
    put(13, F + 5)
    put(14, b["status"] + _skewed(u[:, :, 14], 8))
    put(15, F + 6)
    ntags = u[:, :, 15] % 4
    for k in range(3):
        on = ntags > k
        last = ntags == k + 1
        put(16 + 3 * k, torch.where(on, torch.full_like(ntags, F + 7), torch.full_like(ntags, F + 13)))
        put(17 + 3 * k, torch.where(on, word[:, :, 17 + 3 * k], torch.full_like(ntags, F + 13)))
        put(18 + 3 * k, torch.where(on, torch.where(last, torch.full_like(ntags, F + 7), torch.full_like(ntags, F + 14)),
                                    torch.full_like(ntags, F + 13)))
    put(25, F + 9)
    put(26, num[:, :, 26])
    put(27, F + 10)
    nmsg = 4 + u[:, :, 27] % 6
    for k in range(9):
        s = 28 + 2 * k
        on = nmsg > k
        put(s, torch.where(on, word[:, :, s], torch.full_like(nmsg, F + 13)))
        put(s + 1, torch.where(on, torch.full_like(nmsg, F + 11), torch.full_like(nmsg, F + 13)))
    put(46, word[:, :, 46])
    put(47, F + 12)
    return _assemble(pid.view(count, R * P), t, block_size)


def _log_blocks(first, count, block_size, seed, device):
    _, _, t, b = _tables(device)
    R = block_size // 110 + 8
    P = 24
    F = b["fixed"]
    blk = torch.arange(first, first + count, device=device, dtype=torch.int64).view(-1, 1, 1)
    rec = torch.arange(R, device=device, dtype=torch.int64).view(1, -1, 1)
    slot = torch.arange(P, device=device, dtype=torch.int64).view(1, 1, -1)
    idx = (blk * R + rec) * P + slot
    u = rand_u(idx, 2, seed)
    rid = (blk * R + rec).expand(count, R, 1)[:, :, 0]
    pid = torch.full((count, R, P), F + 14, dtype=torch.int64, device=device)
    ip = rand_u(_skewed(u[:, :, 0], 65536), 3, seed)  # 64 K address pool

    def put(s, v):
        pid[:, :, s] = v

    put(0, b["octet"] + ip % 223 + 1)
    put(1, F + 0)
    put(2, b["octet"] + (ip >> 8) % 256)
    put(3, F + 0)
    put(4, b["octet"] + (ip >> 16) % 256)
    put(5, F + 0)
    put(6, b["octet"] + (ip >> 24) % 254 + 1)
    put(7, F + 1)
    put(8, b["d2"] + (rid // 3600) % 24)
    put(9, F + 2)
    put(10, b["d2"] + (rid // 60) % 60)
    put(11, F + 2)
    put(12, b["d2"] + rid % 60)
    put(13, F + 3)
    put(14, b["path"] + _skewed(u[:, :, 14], 8192))
    put(15, F + 4)
    put(16, F + 8 + _skewed(u[:, :, 16], 5))
    put(17, F + 5)
    put(18, b["num"] + u[:, :, 18] % 10000)
    put(19, F + 6)
    put(20, b["ua"] + _skewed(u[:, :, 20], 256))
    put(21, F + 7)
    return _assemble(pid.view(count, R * P), t, block_size)


def _text_blocks(first, count, block_size, seed, device):
    t, b, _, _ = _tables(device)
    P = block_size // 3 + 16
    blk = torch.arange(first, first + count, device=device, dtype=torch.int64).view(-1, 1)
    slot = torch.arange(P, device=device, dtype=torch.int64).view(1, -1)
    u = rand_u(blk * P + slot, 4, seed)
    word = b["word"] + _skewed(u, 4096)
    pid = torch.where(slot % 2 == 0, word, torch.full_like(word, b["fixed"] + 11))
    return _assemble(pid, t, block_size)


def _binary_blocks(first, count, block_size, seed, device):
    n = (block_size + 19) // 20
    blk = torch.arange(first, first + count, device=device, dtype=torch.int64).view(-1, 1)
    rec = torch.arange(n, device=device, dtype=torch.int64).view(1, -1)
    ctr = blk * n + rec
    u = rand_u(ctr, 5, seed)
    out = torch.zeros((count, n, 20), dtype=torch.uint8, device=device)
    for k in range(8):
        out[:, :, k] = ((ctr >> (8 * k)) & 0xff).to(torch.uint8)
    out[:, :, 8] = (u & 0xff).to(torch.uint8)          # noisy mantissa
    out[:, :, 9] = ((u >> 8) & 0x3f).to(torch.uint8)
    out[:, :, 10] = 0x80
    out[:, :, 11] = 0x3f
    out[:, :, 12] = ((u >> 16) % 12).to(torch.uint8)   # u16 enum
    return out.view(count, n * 20)[:, :block_size].contiguous()


def _random_blocks(first, count, block_size, seed, device):
    n = (block_size + 7) // 8
    blk = torch.arange(first, first + count, device=device, dtype=torch.int64).view(-1, 1)
    w = torch.arange(n, device=device, dtype=torch.int64).view(1, -1)
    x = mix64((blk * n + w) * _i64(0x2545f4914f6cdd1d) + _i64(seed * 0x9e3779b97f4a7c15 + 77))
    return x.view(torch.uint8).view(count, n * 8)[:, :block_size].contiguous()


_KINDS = {"json": _json_blocks, "log": _log_blocks, "text": _text_blocks, "binary": _binary_blocks,
          "random": _random_blocks}


def make_blocks(kind, nblocks, block_size, seed=0x6d696e6c7a, device="cpu", first=0, chunk=None):
    """Returns a [nblocks, block_size] uint8 tensor on `device`."""
    fn = _KINDS[kind]
    device = torch.device(device)
    if chunk is None:
        chunk = max(1, (64 << 20) // block_size) if device.type == "cuda" else max(1, (8 << 20) // block_size)
    out = torch.empty((nblocks, block_size), dtype=torch.uint8, device=device)
    for b0 in range(0, nblocks, chunk):
        c = min(chunk, nblocks - b0)
        out[b0:b0 + c] = fn(first + b0, c, block_size, seed, device)
    return out
