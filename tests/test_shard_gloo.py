"""World-size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: block
ranges, the all-gather of encoded sizes and the stream offsets every rank
derives from it.  Encoded sizes come from the oracle (tests only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minlz_b200 import shard


def test_block_range_partitions():
    for n in (0, 1, 7, 8, 4096, 4097):
        for w in (1, 2, 3, 8):
            got = []
            for r in range(w):
                lo, hi = shard.block_range(n, r, w)
                assert 0 <= lo <= hi <= n
                got += list(range(lo, hi))
            assert got == list(range(n))
            sizes = [shard.block_range(n, r, w)[1] - shard.block_range(n, r, w)[0] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lens, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = len(lens)
        lo, hi = shard.block_range(n, rank, world)
        local = torch.tensor(lens[lo:hi], dtype=torch.int32)
        all_len = shard.gather_block_lengths(local, n)
        off = shard.stream_offsets(all_len, per_block_overhead=8)
        q.put((rank, all_len.tolist(), off.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_lengths_and_offsets(oracle, world):
    import synth
    blocks = synth.make_blocks("json", 7, 1 << 14).numpy()
    lens = [len(oracle.encode_block(blocks[i], 1)) for i in range(7)]
    assert all(x > 0 for x in lens)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lens, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_off = np.concatenate([[0], np.cumsum(np.array(lens) + 8)]).tolist()
    for rank, all_len, off in res:
        assert all_len == lens
        assert off == want_off
