"""ctypes binding of oracle/_ref/libminlz_ref.so: the reference's own AMD64
assembly (asm_amd64.s), made runnable by p9_to_gas.py + ref_shim.c.

TEST INFRASTRUCTURE ONLY -- never imported by minlz_b200.  The library is built
from /root/reference where it lies (`make -C oracle ref`); the built .so is
git-ignored but travels to the GPU box, where /root/reference does not exist.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "_ref", "libminlz_ref.so")
REF_SRC = "/root/reference/asm_amd64.s"


def available():
    return os.path.exists(_SO) or os.path.exists(REF_SRC)


def build(force=False):
    """Builds oracle/_ref when the reference tree is present; returns the path or None."""
    if os.path.exists(REF_SRC):
        if force or not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", _DIR, "-s", "ref"])
    return _SO if os.path.exists(_SO) else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            raise RuntimeError("oracle/_ref/libminlz_ref.so is missing and /root/reference is not here to build it")
        L = C.CDLL(_SO)
        p = C.c_void_p
        for f in (L.mzr_encode_block_l0, L.mzr_encode_block_l1, L.mzr_encode_block_l2):
            f.restype = C.c_int64
            f.argtypes = [p, C.c_size_t, p, C.c_size_t]
        L.mzr_decode_block.restype = C.c_int
        L.mzr_decode_block.argtypes = [p, C.c_size_t, p, C.c_size_t]
        L.mzr_emit_literal.restype = C.c_int64
        L.mzr_emit_literal.argtypes = [p, C.c_size_t, p, C.c_size_t]
        L.mzr_emit_repeat.restype = C.c_int64
        L.mzr_emit_repeat.argtypes = [p, C.c_size_t, C.c_int64]
        L.mzr_emit_copy.restype = C.c_int64
        L.mzr_emit_copy.argtypes = [p, C.c_size_t, C.c_int64, C.c_int64]
        for f in (L.mzr_emit_copy_lits2, L.mzr_emit_copy_lits3):
            f.restype = C.c_int64
            f.argtypes = [p, C.c_size_t, p, C.c_size_t, C.c_int64, C.c_int64]
        L.mzr_match_len.restype = C.c_int64
        L.mzr_match_len.argtypes = [p, C.c_size_t, p, C.c_size_t]
        L.mzr_encode_batch_mt.restype = C.c_int
        L.mzr_encode_batch_mt.argtypes = [C.c_int, C.c_int, p, p, p, p, p, C.c_int]
        L.mzr_decode_batch_mt.restype = C.c_int
        L.mzr_decode_batch_mt.argtypes = [C.c_int, p, p, p, p, p, C.c_int]
        _lib = L
    return _lib


def _in(b):
    return np.ascontiguousarray(b, dtype=np.uint8) if isinstance(b, np.ndarray) else np.frombuffer(bytes(b), dtype=np.uint8)


def encode_block(src, level):
    """encodeBlockFast / encodeBlock / encodeBlockBetter of encode_amd64.go (header not written);
    b'' when the assembly returns 0 (incompressible)."""
    s = _in(src)
    # len(dst) as Encode hands it over: MaxEncodedLen(n) minus the header bytes already written
    cap = s.size + 2
    dst = np.zeros(cap + 64, dtype=np.uint8)      # slack: the assembly's wide copies may write past dst[d:]
    f = {-1: lib().mzr_encode_block_l0, 1: lib().mzr_encode_block_l1, 2: lib().mzr_encode_block_l2}[level]
    n = f(dst.ctypes.data, cap, s.ctypes.data if s.size else 0, s.size)
    assert 0 <= n <= cap
    return dst[:n].tobytes()


def decode_block(src, dlen):
    """minLZDecode of decode_amd64.go: (status, dst bytes)."""
    s = _in(src)
    pad = np.zeros(s.size + 64, dtype=np.uint8)   # the Go slice has no slack either; keep reads in our memory
    pad[:s.size] = s
    dst = np.zeros(max(dlen, 1) + 64, dtype=np.uint8)
    st = lib().mzr_decode_block(dst.ctypes.data, dlen, pad.ctypes.data, s.size)
    return st, dst[:dlen].tobytes()


def encode_batch_mt(level, src, src_off, dst, dst_off, nthreads):
    nblk = len(src_off) - 1
    out_len = np.zeros(nblk, dtype=np.uint32)
    r = lib().mzr_encode_batch_mt(level, nblk, src.ctypes.data, src_off.ctypes.data, dst.ctypes.data,
                                  dst_off.ctypes.data, out_len.ctypes.data, nthreads)
    assert r == 0
    return out_len


def decode_batch_mt(src, src_off, dst, dst_off, nthreads):
    nblk = len(src_off) - 1
    status = np.zeros(nblk, dtype=np.int32)
    r = lib().mzr_decode_batch_mt(nblk, src.ctypes.data, src_off.ctypes.data, dst.ctypes.data,
                                  dst_off.ctypes.data, status.ctypes.data, nthreads)
    assert r == 0
    return status
