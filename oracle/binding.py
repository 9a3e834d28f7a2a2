"""ctypes binding of oracle/libminlz_oracle.so (see minlz_oracle.h).

TEST INFRASTRUCTURE ONLY -- never imported by minlz_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libminlz_oracle.so")

ERR_CORRUPT, ERR_TOO_LARGE, ERR_UNSUPPORTED, ERR_INVALID_LEVEL, ERR_DST_TOO_SMALL = -1, -2, -3, -4, -5


def build(force=False):
    src = os.path.join(_DIR, "minlz_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u8p, u64p, u32p, i32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
        L.mzo_max_encoded_len.restype = C.c_int64
        L.mzo_max_encoded_len.argtypes = [C.c_int64]
        for f in (L.mzo_encode_block_l0, L.mzo_encode_block_l1, L.mzo_encode_block_l2,
                  L.mzo_encode_block_l0_asm, L.mzo_encode_block_l1_asm, L.mzo_encode_block_l2_asm):
            f.restype = C.c_int64
            f.argtypes = [u8p, u8p, C.c_size_t]
        L.mzo_decode_block.restype = C.c_int
        L.mzo_decode_block.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t]
        for f in (L.mzo_encode, L.mzo_try_encode):
            f.restype = C.c_int64
            f.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_int]
        L.mzo_decode.restype = C.c_int64
        L.mzo_decode.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t]
        L.mzo_decoded_len.restype = C.c_int64
        L.mzo_decoded_len.argtypes = [u8p, C.c_size_t]
        L.mzo_crc.restype = C.c_uint32
        L.mzo_crc.argtypes = [u8p, C.c_size_t]
        L.mzo_emit_literal.restype = C.c_int
        L.mzo_emit_literal.argtypes = [u8p, u8p, C.c_size_t]
        L.mzo_emit_repeat.restype = C.c_int
        L.mzo_emit_repeat.argtypes = [u8p, C.c_int]
        L.mzo_emit_copy.restype = C.c_int
        L.mzo_emit_copy.argtypes = [u8p, C.c_int, C.c_int]
        L.mzo_emit_copy_lits2.restype = C.c_int
        L.mzo_emit_copy_lits2.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int]
        L.mzo_emit_copy_lits3.restype = C.c_int
        L.mzo_emit_copy_lits3.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int]
        L.mzo_encode_batch_mt.restype = C.c_int
        L.mzo_encode_batch_mt.argtypes = [C.c_int, C.c_int, u8p, u64p, u8p, u64p, u32p, C.c_int]
        L.mzo_decode_batch_mt.restype = C.c_int
        L.mzo_decode_batch_mt.argtypes = [C.c_int, u8p, u64p, u8p, u64p, i32p, C.c_int]
        _lib = L
    return _lib


def _in(b):
    a = np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b, dtype=np.uint8)
    if a.size == 0:
        a = np.zeros(1, dtype=np.uint8)[:0]
    return a


def _ptr(a):
    return a.ctypes.data if a.size else 0


def max_encoded_len(n):
    return int(lib().mzo_max_encoded_len(n))


def encode_block(src, level, flavor="go"):
    """encodeBlock / encodeBlockBetter: token stream without header; b'' = 0 (incompressible).
    flavor "go" = the pure-Go functions (noasm build), "asm" = the amd64 assembly flavour."""
    s = _in(src)
    dst = np.empty(s.size + 64, dtype=np.uint8)
    if flavor == "asm":
        f = {-1: lib().mzo_encode_block_l0_asm, 1: lib().mzo_encode_block_l1_asm, 2: lib().mzo_encode_block_l2_asm}[level]
    else:
        f = {-1: lib().mzo_encode_block_l0, 1: lib().mzo_encode_block_l1, 2: lib().mzo_encode_block_l2}[level]
    n = f(dst.ctypes.data, _ptr(s), s.size)
    return dst[:n].tobytes()


def decode_block(src, dlen):
    """minLZDecode: returns (status, dst bytes)."""
    s = _in(src)
    dst = np.zeros(max(dlen, 1), dtype=np.uint8)
    st = lib().mzo_decode_block(dst.ctypes.data, dlen, _ptr(s), s.size)
    return st, dst[:dlen].tobytes()


def encode(src, level):
    s = _in(src)
    cap = max(max_encoded_len(s.size), 1)
    dst = np.empty(cap, dtype=np.uint8)
    n = lib().mzo_encode(dst.ctypes.data, cap, _ptr(s), s.size, level)
    if n < 0:
        return n
    return dst[:n].tobytes()


def try_encode(src, level):
    s = _in(src)
    cap = max(max_encoded_len(s.size), 1)
    dst = np.empty(cap, dtype=np.uint8)
    n = lib().mzo_try_encode(dst.ctypes.data, cap, _ptr(s), s.size, level)
    return dst[:n].tobytes() if n > 0 else None


def decoded_len(src):
    s = _in(src)
    return int(lib().mzo_decoded_len(_ptr(s), s.size))


def decode(src):
    """Decode: bytes, or a negative MZO_ERR_* code."""
    s = _in(src)
    dl = decoded_len(s)
    if dl < 0:
        return dl
    dst = np.zeros(max(dl, 1), dtype=np.uint8)
    n = lib().mzo_decode(dst.ctypes.data, dst.size, _ptr(s), s.size)
    if n < 0:
        return int(n)
    return dst[:n].tobytes()


def crc(b):
    s = _in(b)
    return int(lib().mzo_crc(_ptr(s), s.size))


def encode_batch_mt(level, src, src_off, dst, dst_off, nthreads):
    nblk = len(src_off) - 1
    out_len = np.zeros(nblk, dtype=np.uint32)
    r = lib().mzo_encode_batch_mt(level, nblk, src.ctypes.data, src_off.ctypes.data, dst.ctypes.data,
                                  dst_off.ctypes.data, out_len.ctypes.data, nthreads)
    assert r == 0
    return out_len


def decode_batch_mt(src, src_off, dst, dst_off, nthreads):
    nblk = len(src_off) - 1
    status = np.zeros(nblk, dtype=np.int32)
    r = lib().mzo_decode_batch_mt(nblk, src.ctypes.data, src_off.ctypes.data, dst.ctypes.data,
                                  dst_off.ctypes.data, status.ctypes.data, nthreads)
    assert r == 0
    return status
