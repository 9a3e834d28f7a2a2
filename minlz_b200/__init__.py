"""minlz_b200 -- B200-native MinLZ block encode/decode behind the reference API.

Host-side mirror of the reference's block API (minio/minlz encode.go /
decode.go) over the C ABI of libminlz_cuda.so.  Names, argument meaning and
error behaviour follow the Go package so the tests read like the reference's:

    Encode(dst, src, level) / AppendEncoded / TryEncode / MaxEncodedLen
    Decode(dst, src) / AppendDecoded / DecodedLen / IsMinLZ
    EncodeBatch / DecodeBatch               (new: one GPU launch per batch)
    encode_blocks_dev / decode_blocks_dev   (device tensors, the seam level)

All compression work runs in CUDA kernels; there is no CPU codec here and no
fallback: without the built library or without a GPU the calls raise.
"""
import ctypes as C

import numpy as np

from . import _lib

MaxBlockSize = 8 << 20  # minlz.go:24

LevelSuperFast = -1
LevelUncompressed = 0  # encode.go:20-42
LevelFastest = 1
LevelBalanced = 2

_OK, _E_CORRUPT, _E_TOO_LARGE, _E_UNSUPPORTED, _E_LEVEL, _E_DST, _E_CUDA, _E_ARG = 0, -1, -2, -3, -4, -5, -6, -7


# Encoder flavour (include/minlz_cuda.h): which of the reference's two builds the
# encoder mirrors byte for byte -- the pure-Go functions (`noasm`) or the amd64 assembly.
FlavorGo = 0
FlavorAMD64 = 1


def set_encoder_flavor(flavor):
    """Process-wide, like the build tag it stands for (asm_none.go:15 / encode_amd64.go:15)."""
    rc = _lib.load().mzcu_set_encoder_flavor(int(flavor))
    if rc:
        _raise(rc)


def get_encoder_flavor():
    return int(_lib.load().mzcu_get_encoder_flavor())


class MinLZError(Exception):
    """Base of the errors mirrored from decode.go:29-40."""


class ErrCorrupt(MinLZError):
    def __init__(self, msg="minlz: corrupt input", partial=None):
        super().__init__(msg)
        self.partial = partial  # Go returns (dst, ErrCorrupt): the partially decoded buffer


class ErrTooLarge(MinLZError):
    def __init__(self, msg="minlz: decoded block is too large"):
        super().__init__(msg)


class ErrUnsupported(MinLZError):
    def __init__(self, msg="minlz: unsupported input"):
        super().__init__(msg)


class ErrInvalidLevel(MinLZError):
    def __init__(self, msg="minlz: invalid compression level"):
        super().__init__(msg)


class CudaError(RuntimeError):
    pass


def _raise(code):
    msg = _lib.load().mzcu_last_error().decode("utf-8", "replace")
    if code == _E_CORRUPT:
        raise ErrCorrupt()
    if code == _E_TOO_LARGE:
        raise ErrTooLarge()
    if code == _E_UNSUPPORTED:
        raise ErrUnsupported()
    if code == _E_LEVEL:
        raise ErrInvalidLevel()
    if code == _E_CUDA:
        raise CudaError("minlz_b200: " + msg)
    raise ValueError("minlz_b200: error %d: %s" % (code, msg))


def _np(b):
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8).reshape(-1)
    return np.frombuffer(bytes(b), dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data if a.size else None


def device_count():
    return int(_lib.load().mzcu_device_count())


def MaxEncodedLen(srcLen):
    """encode.go:234-244."""
    return int(_lib.load().mzcu_max_encoded_len(int(srcLen)))


def Encode(dst, src, level):
    """encode.go:74-139.  `dst` is accepted for signature parity and ignored
    (Python bytes are immutable); returns the encoded block as bytes."""
    s = _np(src)
    n = MaxEncodedLen(s.size)
    if n < 0:
        raise ErrTooLarge()
    if level not in (LevelSuperFast, LevelUncompressed, LevelFastest, LevelBalanced):
        if s.size < 16:
            return b"\x00" + (b"\x00" + s.tobytes() if s.size else b"")  # encode.go:83-85 runs before the level switch
        raise ErrInvalidLevel()
    out = np.empty(max(n, 1), dtype=np.uint8)
    r = _lib.load().mzcu_encode(out.ctypes.data, out.size, _ptr(s), s.size, level)
    if r < 0:
        _raise(r)
    return out[:r].tobytes()


def AppendEncoded(dst, src, level):
    """encode.go:144-163."""
    return bytes(dst or b"") + Encode(None, src, level)


def TryEncode(dst, src, level):
    """encode.go:168-207: None when Go returns nil."""
    s = _np(src)
    n = MaxEncodedLen(s.size)
    if n < 0 or s.size < 16 or level not in (LevelSuperFast, LevelFastest, LevelBalanced):
        return None
    out = np.empty(n, dtype=np.uint8)
    r = _lib.load().mzcu_try_encode(out.ctypes.data, out.size, _ptr(s), s.size, level)
    if r < 0:
        _raise(r)
    return out[:r].tobytes() if r > 0 else None


def IsMinLZ(src):
    """decode.go:114-118: (ok, size); raises on a malformed header."""
    s = _np(src)
    ok, size = C.c_int(0), C.c_int64(0)
    r = _lib.load().mzcu_is_minlz(_ptr(s), s.size, C.byref(ok), C.byref(size))
    if r < 0:
        _raise(r)
    return bool(ok.value), int(size.value)


def DecodedLen(src):
    """decode.go:107-111."""
    s = _np(src)
    r = _lib.load().mzcu_decoded_len(_ptr(s), s.size)
    if r < 0:
        _raise(r)
    return int(r)


def Decode(dst, src):
    """decode.go:50-78 (MinLZ blocks; the Snappy/S2 fallback for blocks whose
    first byte is not 0 stays in host Go and raises ErrUnsupported here)."""
    s = _np(src)
    ok, dlen = IsMinLZ(s)
    # nothing is allocated before the block is known to be MinLZ of a legal size (decode.go:59-68:
    # a first byte != 0 is a Snappy varint of up to 4 GiB; s2 rejects it above its MaxBlockSize)
    if not ok:
        raise ErrUnsupported()
    if dlen > MaxBlockSize:
        raise ErrTooLarge()
    out = np.zeros(max(dlen, 1), dtype=np.uint8)
    r = _lib.load().mzcu_decode(out.ctypes.data, dlen, _ptr(s), s.size)
    if r == _E_CORRUPT:
        raise ErrCorrupt(partial=out[:dlen].tobytes())
    if r < 0:
        _raise(r)
    return out[:r].tobytes()


def AppendDecoded(dst, src):
    """decode.go:85-103."""
    return bytes(dst or b"") + Decode(None, src)


# ---- batch forms (host memory) ---------------------------------------------

def _offsets(sizes):
    off = np.zeros(len(sizes) + 1, dtype=np.uint64)
    np.cumsum(np.asarray(sizes, dtype=np.uint64), out=off[1:])
    return off


def EncodeBatch(blocks, level, device=-1):
    """Encode() over a list of inputs with one GPU launch; returns list of bytes."""
    if level not in (LevelSuperFast, LevelUncompressed, LevelFastest, LevelBalanced):
        raise ErrInvalidLevel()
    arrs = [_np(b) for b in blocks]
    for a in arrs:
        if a.size > MaxBlockSize:
            raise ErrTooLarge()
    if not arrs:
        return []
    src = np.concatenate(arrs) if sum(a.size for a in arrs) else np.zeros(0, dtype=np.uint8)
    soff = _offsets([a.size for a in arrs])
    doff = _offsets([MaxEncodedLen(a.size) for a in arrs])
    dst = np.empty(int(doff[-1]), dtype=np.uint8)
    enc = np.zeros(len(arrs), dtype=np.uint64)
    r = _lib.load().mzcu_encode_batch(device, level, len(arrs), _ptr(src), soff.ctypes.data, dst.ctypes.data,
                                      doff.ctypes.data, enc.ctypes.data)
    if r < 0:
        _raise(r)
    return [dst[int(doff[i]):int(doff[i]) + int(enc[i])].tobytes() for i in range(len(arrs))]


def DecodeBatch(blocks, device=-1):
    """Decode() over a list of encoded blocks with one GPU launch.  Returns a
    list whose items are bytes or a MinLZError instance (per-block failure)."""
    arrs = [_np(b) for b in blocks]
    if not arrs:
        return []
    sizes = []
    pre = {}
    for i, a in enumerate(arrs):
        try:
            ok, n = IsMinLZ(a)
            if not ok:
                raise ErrUnsupported()
            if n > MaxBlockSize:
                raise ErrTooLarge()
            sizes.append(n)
        except MinLZError as e:
            pre[i] = e
            sizes.append(0)
    src = np.concatenate(arrs) if sum(a.size for a in arrs) else np.zeros(0, dtype=np.uint8)
    soff = _offsets([a.size for a in arrs])
    doff = _offsets(sizes)
    dst = np.zeros(max(int(doff[-1]), 1), dtype=np.uint8)
    dl = np.zeros(len(arrs), dtype=np.int64)
    r = _lib.load().mzcu_decode_batch(device, len(arrs), _ptr(src), soff.ctypes.data, dst.ctypes.data, doff.ctypes.data,
                                      dl.ctypes.data)
    if r < 0:
        _raise(r)
    out = []
    for i in range(len(arrs)):
        if i in pre:
            out.append(pre[i])
        elif dl[i] == _E_CORRUPT:
            out.append(ErrCorrupt(partial=dst[int(doff[i]):int(doff[i + 1])].tobytes()))
        elif dl[i] == _E_UNSUPPORTED:
            out.append(ErrUnsupported())
        elif dl[i] == _E_TOO_LARGE:
            out.append(ErrTooLarge())
        elif dl[i] < 0:
            out.append(MinLZError("error %d" % dl[i]))
        else:
            out.append(dst[int(doff[i]):int(doff[i]) + int(dl[i])].tobytes())
    return out


# ---- seam level on host arrays (encodeBlock / minLZDecode over a batch) -----

def encode_blocks(src, src_off, level, device=-1):
    """Host numpy in/out.  Returns (dst, dst_off, out_len): out_len[i] == 0
    means block i is not compressible (caller stores it raw)."""
    src = _np(src)
    src_off = np.ascontiguousarray(src_off, dtype=np.uint64)
    nblk = len(src_off) - 1
    sizes = (src_off[1:] - src_off[:-1]).astype(np.int64)
    dst_off = _offsets(sizes + 2)
    dst = np.empty(int(dst_off[-1]) + 16, dtype=np.uint8)
    out_len = np.zeros(max(nblk, 1), dtype=np.uint32)
    r = _lib.load().mzcu_encode_blocks(device, level, nblk, _ptr(src), src_off.ctypes.data, dst.ctypes.data,
                                       dst_off.ctypes.data, out_len.ctypes.data)
    if r < 0:
        _raise(r)
    return dst, dst_off, out_len[:nblk]


def decode_blocks(src, src_off, dst_off, device=-1):
    """Host numpy in/out.  Returns (dst, status)."""
    src = _np(src)
    src_off = np.ascontiguousarray(src_off, dtype=np.uint64)
    dst_off = np.ascontiguousarray(dst_off, dtype=np.uint64)
    nblk = len(src_off) - 1
    dst = np.zeros(max(int(dst_off[-1]), 1), dtype=np.uint8)
    status = np.zeros(max(nblk, 1), dtype=np.int32)
    r = _lib.load().mzcu_decode_blocks(device, nblk, _ptr(src), src_off.ctypes.data, dst.ctypes.data, dst_off.ctypes.data,
                                       status.ctypes.data)
    if r < 0:
        _raise(r)
    return dst[:int(dst_off[-1])], status[:nblk]


# ---- seam level on device tensors (what bench.py times) ---------------------

def _stream_ptr(stream):
    import torch
    st = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(st.cuda_stream)


def encode_blocks_dev(src, src_off, dst, dst_off, out_len, level, stream=None):
    """torch CUDA tensors: src/dst uint8, src_off/dst_off int64 (nblk+1),
    out_len int32 (nblk).  Asynchronous on `stream` (default: current)."""
    nblk = src_off.numel() - 1
    r = _lib.load().mzcu_encode_blocks_dev(src.device.index, level, nblk, src.data_ptr(), src_off.data_ptr(),
                                           dst.data_ptr(), dst_off.data_ptr(), out_len.data_ptr(), _stream_ptr(stream))
    if r < 0:
        _raise(r)


def decode_blocks_dev(src, src_off, dst, dst_off, status, stream=None):
    """torch CUDA tensors: src/dst uint8, offsets int64 (nblk+1), status int32."""
    nblk = src_off.numel() - 1
    r = _lib.load().mzcu_decode_blocks_dev(src.device.index, nblk, src.data_ptr(), src_off.data_ptr(), dst.data_ptr(),
                                           dst_off.data_ptr(), status.data_ptr(), _stream_ptr(stream))
    if r < 0:
        _raise(r)


def pack_blocks_dev(src, src_off, lens, dst, dst_off, stream=None):
    """Dense-packs encoder output on the device; writes dst_off (int64, nblk+1)."""
    nblk = lens.numel()
    r = _lib.load().mzcu_pack_blocks_dev(src.device.index, nblk, src.data_ptr(), src_off.data_ptr(), lens.data_ptr(),
                                         dst.data_ptr(), dst_off.data_ptr(), _stream_ptr(stream))
    if r < 0:
        _raise(r)


def encode_blocks_into(src, src_off, dst, dst_off, out_len, level, device=-1):
    """Host numpy buffers, caller-allocated (e.g. pinned): the C ABI host call."""
    nblk = len(src_off) - 1
    r = _lib.load().mzcu_encode_blocks(device, level, nblk, src.ctypes.data, src_off.ctypes.data, dst.ctypes.data,
                                       dst_off.ctypes.data, out_len.ctypes.data)
    if r < 0:
        _raise(r)


def encode_blocks_packed_into(src, src_off, dst, dst_off_out, level, device=-1):
    """Host numpy buffers; token streams packed back to back into dst, offsets
    (uint64, nblk+1) written to dst_off_out.  Returns the packed size."""
    nblk = len(src_off) - 1
    r = _lib.load().mzcu_encode_blocks_packed(device, level, nblk, src.ctypes.data, src_off.ctypes.data, dst.ctypes.data,
                                              dst.size, dst_off_out.ctypes.data)
    if r < 0:
        _raise(r)
    return int(dst_off_out[nblk])


def decode_blocks_into(src, src_off, dst, dst_off, status, device=-1):
    nblk = len(src_off) - 1
    r = _lib.load().mzcu_decode_blocks(device, nblk, src.ctypes.data, src_off.ctypes.data, dst.ctypes.data,
                                       dst_off.ctypes.data, status.ctypes.data)
    if r < 0:
        _raise(r)


def last_kernel_ms():
    return float(_lib.load().mzcu_last_kernel_ms())
