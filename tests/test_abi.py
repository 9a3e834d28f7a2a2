"""C-ABI checks that need no GPU: the library loads, exports every symbol the
header declares, and the pure helpers behave like the reference's."""
import ctypes as C
import os
import re

import pytest

import minlz_b200
from minlz_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "minlz_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mzcu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 15
    lib = C.CDLL(_lib.SO_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_lib.SYMBOLS) == names, "python binding and header disagree"
    assert lib.mzcu_abi_version() == 2


def test_encoder_flavour_setting():
    # include/minlz_cuda.h: process-wide, Go flavour by default, unknown values refused
    assert minlz_b200.get_encoder_flavor() == minlz_b200.FlavorGo
    minlz_b200.set_encoder_flavor(minlz_b200.FlavorAMD64)
    assert minlz_b200.get_encoder_flavor() == minlz_b200.FlavorAMD64
    minlz_b200.set_encoder_flavor(minlz_b200.FlavorGo)
    with pytest.raises(Exception):
        minlz_b200.set_encoder_flavor(7)
    assert minlz_b200.get_encoder_flavor() == minlz_b200.FlavorGo


def test_max_encoded_len():
    # minlz_test.go:42-69 TestMaxEncodedLen / encode.go:234-244
    assert minlz_b200.MaxEncodedLen(0) == 1
    assert minlz_b200.MaxEncodedLen(1) == 3
    assert minlz_b200.MaxEncodedLen(8 << 20) == (8 << 20) + 2
    assert minlz_b200.MaxEncodedLen((8 << 20) + 1) == -1


def test_header_helpers(oracle):
    # decode.go:120-156 isMinLZ cases
    assert minlz_b200.IsMinLZ(b"\x00") == (True, 0)
    assert minlz_b200.IsMinLZ(b"\x00\x00abc") == (True, 3)
    assert minlz_b200.IsMinLZ(b"\x00\x05hello") == (True, 5)
    with pytest.raises(minlz_b200.ErrCorrupt):
        minlz_b200.IsMinLZ(b"")
    with pytest.raises(minlz_b200.ErrCorrupt):
        minlz_b200.IsMinLZ(b"\x00\x05")           # header only
    with pytest.raises(minlz_b200.ErrCorrupt):
        minlz_b200.IsMinLZ(b"\x00\x02abcdef")     # compressed larger than decoded
    with pytest.raises(minlz_b200.ErrTooLarge):
        minlz_b200.IsMinLZ(b"\x00\x81\x80\x80\x04x")  # 8 MiB + 1
    assert minlz_b200.IsMinLZ(b"\x05hello")[0] is False  # Snappy/S2 style
    for blob in (b"\x00", b"\x00\x00abc", b"\x00\x05hello", b"\x00\x80\x80\x80\x04" + b"x" * 10):
        assert minlz_b200.DecodedLen(blob) == oracle.decoded_len(blob)


def test_no_cpu_fallback_without_gpu():
    if minlz_b200.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(minlz_b200.CudaError):
        minlz_b200.Encode(None, b"x" * 100, minlz_b200.LevelFastest)
    with pytest.raises(minlz_b200.CudaError):
        minlz_b200.Decode(None, b"\x00\x64" + b"\x00" * 4)
    # paths that never reach the device still work, as in the reference
    assert minlz_b200.Encode(None, b"", 1) == b"\x00"
    assert minlz_b200.Encode(None, b"abc", 1) == b"\x00\x00abc"
    assert minlz_b200.Decode(None, b"\x00\x00abc") == b"abc"


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "minlz_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU codec", ""), f
