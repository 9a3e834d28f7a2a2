"""The C++ host mirror (include/minlz.hpp: block API, Index, Writer / Reader over the C ABI).

The reference's host code is Go; without a Go toolchain the host side above the C ABI is C++
(the prompt's rule for compiled references), with the Python package as the second, test-driving
mirror.  tests/cpp/host_mirror_test.cpp runs its own checks (round trips, Skip / Seek / ReadAt like
index_test.go:30,119, error behaviour like decode.go:74-76) and writes what it produced; this file
compares those bytes with the oracle and the Python mirror.
"""
import io
import os
import subprocess

import numpy as np
import pytest

import synth
from minlz_b200 import index as mzi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("cpp") / "host_mirror_test"
    lib = os.path.join(ROOT, "minlz_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-L", lib, "-lminlz_cuda",
                           "-Wl,-rpath," + lib, "-o", str(out)])
    return str(out)


def test_index_matches_python_mirror(exe, tmp_path):
    r = subprocess.run([exe, "index", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stderr
    ix = mzi.Index()
    ix.estBlockUncomp = 1 << 20
    ix.Offsets = [(10 + i * 400000 + (i * i) % 977, i << 20) for i in range(50)]
    assert (tmp_path / "index_regular.bin").read_bytes() == ix.appendTo(b"", 50 << 20, 20000000)
    ix.Offsets = [(c, u + (i % 3) * 1000 + (0 if i else 5)) for i, (c, u) in enumerate(ix.Offsets)]
    assert (tmp_path / "index_irregular.bin").read_bytes() == ix.appendTo(b"", 60 << 20, -1)


@pytest.mark.gpu
def test_block_and_stream_bytes_match(exe, tmp_path, oracle):
    import minlz_b200 as mz
    from minlz_b200 import stream as mzs
    data = synth.make_blocks("json", 1, 3 << 20).numpy()[0].tobytes() + b"the tail"
    src = tmp_path / "input.bin"
    src.write_bytes(data)
    r = subprocess.run([exe, "gpu", str(src), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stderr + r.stdout
    for level in (-1, 1, 2):
        assert (tmp_path / ("block_%d.mzb" % level)).read_bytes() == oracle.encode(data, level)
    assert (tmp_path / "block_0.mzb").read_bytes() == b"\x00\x00" + data
    asm = (tmp_path / "block_1_amd64.mzb").read_bytes()
    assert asm.endswith(oracle.encode_block(data, 1, flavor="asm")) and mz.Decode(None, asm) == data
    # the Python mirror, driven the same way, frames the same bytes
    buf = io.BytesIO()
    w = mzs.NewWriter(buf, mzs.WriterLevel(1), mzs.WriterBlockSize(64 << 10), mzs.WriterAddIndex())
    w.Write(data[:len(data) // 3])
    w.Write(data[len(data) // 3:])
    idx = w.CloseIndex()
    assert (tmp_path / "stream.mz").read_bytes() == buf.getvalue()
    assert (tmp_path / "index.bin").read_bytes() == idx
    assert (tmp_path / "index_stream.bin").read_bytes() == mzi.IndexStream(io.BytesIO(buf.getvalue()))
    assert mzs.NewReader(io.BytesIO((tmp_path / "stream.mz").read_bytes())).Read() == data
