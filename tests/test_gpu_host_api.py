"""GPU tests of the host-side entry points a Go host drives (include/minlz_cuda.h):
several devices behind one call, asynchronous submit / wait, validate mode, host
placement.  Reference behaviour mirrored: Writer.EncodeBuffer / Reader.DecodeConcurrent
fan-out (writer.go:441-563, reader.go:575-992: results in stream order whatever ran
where), debugValidateBlocks (encode.go:108-133)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import minlz_b200 as mz  # noqa: E402
from minlz_b200 import _lib  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected but no CUDA device is visible")


def _batch(nblk, bs, seed=0, kind="json"):
    import synth
    data = synth.make_blocks(kind, nblk, bs, device="cpu", seed=0x6d696e6c7a + seed).numpy().copy()
    rng = np.random.default_rng(seed)
    data[nblk // 3] = rng.integers(0, 256, bs, dtype=np.uint8)  # one incompressible block: out_len 0
    return data


def _device_lists():
    n = mz.device_count()
    lists = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [[0, 1], [1, 0]]
    if n >= 4:
        lists += [[0, 1, 2, 3]]
    return lists


@pytest.mark.parametrize("level", [1, 2])
def test_multi_device_batch_equals_oracle(oracle, level):
    """One call shards the batch over a device list; bytes, sizes, CRCs and decoded output are
    what the single-block oracle gives, in stream order, for every list (a list may name a
    device twice: two host threads then share it)."""
    lib = _lib.load()
    nblk, bs = 37, 192 << 10
    data = _batch(nblk, bs, seed=level)
    flat = data.reshape(-1)
    soff = np.arange(nblk + 1, dtype=np.uint64) * bs
    want = [oracle.encode_block(data[i], level) for i in range(nblk)]
    want_crc = [oracle.crc(data[i].tobytes()) for i in range(nblk)]
    assert want[nblk // 3] == b""
    for devs in _device_lists():
        d = (C.c_int * len(devs))(*devs)
        dst = np.zeros(flat.size + 64, dtype=np.uint8)
        doff = np.zeros(nblk, dtype=np.uint64)
        olen = np.zeros(nblk, dtype=np.uint32)
        crc = np.zeros(nblk, dtype=np.uint32)
        r = lib.mzcu_stream_encode_blocks_multi(len(devs), d, level, nblk, flat.ctypes.data, soff.ctypes.data, dst.ctypes.data,
                                                dst.size, doff.ctypes.data, olen.ctypes.data, crc.ctypes.data)
        assert r == 0, lib.mzcu_last_error()
        for i in range(nblk):
            assert dst[int(doff[i]):int(doff[i]) + int(olen[i])].tobytes() == want[i], (devs, i)
            assert int(crc[i]) == want_crc[i], (devs, i)
        # decode the compressible blocks from the very layout the encode call produced
        live = np.nonzero(olen)[0]
        sb = np.ascontiguousarray(doff[live])
        sl = np.ascontiguousarray(olen[live])
        out = np.zeros(live.size * bs, dtype=np.uint8)
        ooff = np.arange(live.size + 1, dtype=np.uint64) * bs
        status = np.ones(live.size, dtype=np.int32)
        dcrc = np.zeros(live.size, dtype=np.uint32)
        r = lib.mzcu_stream_decode_blocks_multi(len(devs), d, live.size, dst.ctypes.data, sb.ctypes.data, sl.ctypes.data,
                                                out.ctypes.data, ooff.ctypes.data, status.ctypes.data, dcrc.ctypes.data)
        assert r == 0, lib.mzcu_last_error()
        assert not status.any()
        assert np.array_equal(out.reshape(live.size, bs), data[live])
        assert [int(c) for c in dcrc] == [want_crc[i] for i in live]


def test_multi_device_argument_errors():
    lib = _lib.load()
    soff = np.array([0, 100], dtype=np.uint64)
    src = np.zeros(100, dtype=np.uint8)
    dst = np.zeros(200, dtype=np.uint8)
    doff = np.zeros(1, dtype=np.uint64)
    olen = np.zeros(1, dtype=np.uint32)
    args = (1, src.ctypes.data, soff.ctypes.data, dst.ctypes.data, dst.size, doff.ctypes.data, olen.ctypes.data, None)
    assert lib.mzcu_stream_encode_blocks_multi(0, None, 1, *args) == -7  # MZCU_ERR_INVALID_ARG: empty device list
    d = (C.c_int * 1)(0)
    assert lib.mzcu_stream_encode_blocks_multi(1, d, 9, *args) == -4     # ErrInvalidLevel
    bad = (C.c_int * 1)(63)
    assert lib.mzcu_stream_encode_blocks_multi(1, bad, 1, *args) == -6   # no such device: CUDA error, no fallback
    assert b"device 63" in lib.mzcu_last_error()


def test_async_submit_wait_overlapping_calls(oracle):
    """Two encode jobs and a decode job in flight at once; every result equals the oracle's."""
    lib = _lib.load()
    nblk, bs = 64, 256 << 10
    jobs = []
    for k in range(3):
        data = _batch(nblk, bs, seed=10 + k)
        flat = data.reshape(-1)
        soff = np.arange(nblk + 1, dtype=np.uint64) * bs
        dst = np.zeros(flat.size + 64, dtype=np.uint8)
        poff = np.zeros(nblk + 1, dtype=np.uint64)
        crc = np.zeros(nblk, dtype=np.uint32)
        j = lib.mzcu_submit_stream_encode_blocks(-1, 1, nblk, flat.ctypes.data, soff.ctypes.data, dst.ctypes.data, dst.size,
                                                 poff.ctypes.data, crc.ctypes.data)
        assert j > 0
        jobs.append((j, data, flat, soff, dst, poff, crc))
    dec_jobs = []
    for j, data, flat, soff, dst, poff, crc in jobs:
        assert lib.mzcu_wait(j) == 0, lib.mzcu_last_error()
        for i in range(0, nblk, 7):
            assert dst[int(poff[i]):int(poff[i + 1])].tobytes() == oracle.encode_block(data[i], 1)
            assert int(crc[i]) == oracle.crc(data[i].tobytes())
        # stored block: empty range; decode the others
        live = np.nonzero(poff[1:] > poff[:-1])[0]
        assert live.size == nblk - 1
        # the stored block's range is empty, so dropping one of its two equal offsets leaves the
        # offset table of the live blocks
        so = np.ascontiguousarray(np.delete(poff, nblk // 3))
        out = np.zeros(live.size * bs, dtype=np.uint8)
        ooff = np.arange(live.size + 1, dtype=np.uint64) * bs
        status = np.ones(live.size, dtype=np.int32)
        dj = lib.mzcu_submit_stream_decode_blocks(-1, live.size, dst.ctypes.data, so.ctypes.data, out.ctypes.data, ooff.ctypes.data,
                                                  status.ctypes.data, None)
        assert dj > 0
        dec_jobs.append((dj, data, live, out, status, so, ooff))
    for dj, data, live, out, status, so, ooff in dec_jobs:
        assert lib.mzcu_wait(dj) == 0, lib.mzcu_last_error()
        assert not status.any()
        assert np.array_equal(out.reshape(live.size, -1), data[live])
    assert lib.mzcu_wait(123456) == -7  # unknown job


def test_validate_mode(oracle):
    """MZCU_VALIDATE / mzcu_set_validate: decode-after-encode on the device (encode.go:108-133).
    Clean encodes pass through every entry point; results are unchanged."""
    lib = _lib.load()
    assert lib.mzcu_get_validate() == 0
    lib.mzcu_set_validate(1)
    try:
        assert lib.mzcu_get_validate() == 1
        nblk, bs = 70, 256 << 10
        data = _batch(nblk, bs, seed=5)
        flat = data.reshape(-1)
        soff = np.arange(nblk + 1, dtype=np.uint64) * bs
        for level in (-1, 1, 2):
            dst, doff, out_len = mz.encode_blocks(flat, soff, level)            # mzcu_encode_blocks
            for i in range(0, nblk, 9):
                assert dst[int(doff[i]):int(doff[i]) + int(out_len[i])].tobytes() == oracle.encode_block(data[i], level)
            pk = np.zeros(flat.size + 64, dtype=np.uint8)
            poff = np.zeros(nblk + 1, dtype=np.uint64)
            mz.encode_blocks_packed_into(flat, soff, pk, poff, level)            # mzcu_encode_blocks_packed
            for i in range(0, nblk, 9):
                assert pk[int(poff[i]):int(poff[i + 1])].tobytes() == oracle.encode_block(data[i], level)
        dev = torch.device("cuda:0")
        t_src = torch.from_numpy(flat).to(dev)
        t_soff = torch.from_numpy(soff.astype(np.int64)).to(dev)
        cap = bs + 16
        t_eoff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * cap
        t_enc = torch.empty(nblk * cap, dtype=torch.uint8, device=dev)
        t_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
        mz.encode_blocks_dev(t_src, t_soff, t_enc, t_eoff, t_len, 1)             # mzcu_encode_blocks_dev
        torch.cuda.synchronize()
        assert int(t_len[0]) == len(oracle.encode_block(data[0], 1))
        assert mz.Encode(None, data[1].tobytes(), 1) == oracle.encode(data[1].tobytes(), 1)
    finally:
        lib.mzcu_set_validate(0)


def test_validate_pass_catches_a_damaged_block():
    """The validate pass itself: a slot whose bytes were damaged after encoding is named."""
    lib = _lib.load()
    nblk, bs = 12, 128 << 10
    data = _batch(nblk, bs, seed=8)
    dev = torch.device("cuda:0")
    t_src = torch.from_numpy(data.reshape(-1)).to(dev)
    t_soff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * bs
    cap = bs + 16
    t_eoff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * cap
    t_enc = torch.zeros(nblk * cap, dtype=torch.uint8, device=dev)
    t_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
    mz.encode_blocks_dev(t_src, t_soff, t_enc, t_eoff, t_len, 1)
    torch.cuda.synchronize()
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    args = (0, nblk, t_src.data_ptr(), t_soff.data_ptr(), t_enc.data_ptr(), t_eoff.data_ptr(), t_len.data_ptr(), sp)
    assert lib.mzcu_validate_blocks_dev(*args) == 0, lib.mzcu_last_error()
    for victim in (7, 2):
        pos = victim * cap + int(t_len[victim]) // 2
        t_enc[pos] ^= 0x5a
        torch.cuda.synchronize()
        assert lib.mzcu_validate_blocks_dev(*args) == -8  # MZCU_ERR_VALIDATE
        assert ("block %d " % victim).encode() in lib.mzcu_last_error()


def test_bind_host_to_device():
    lib = _lib.load()
    node = lib.mzcu_bind_host_to_device(0)
    assert node >= -1
    # still able to run work afterwards
    assert mz.Decode(None, mz.Encode(None, b"abcd" * 1000, 1)) == b"abcd" * 1000
