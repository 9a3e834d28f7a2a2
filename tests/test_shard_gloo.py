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


def _funnel_worker(rank, world, port, nblk, bs, q):
    """Rank 0 holds the batch: scatter raw blocks -> encode (oracle, CPU) -> gather the packed
    stream to rank 0 -> scatter it back -> decode -> gather the blocks.  SURVEY 8(e)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import synth
        from oracle import binding as oracle
        full = synth.make_blocks("json", nblk, bs).reshape(-1) if rank == 0 else None
        local = shard.scatter_rows(full, nblk, bs, src=0, device=torch.device("cpu"))
        lo, hi = shard.block_range(nblk, rank, world)
        assert local.numel() == (hi - lo) * bs
        toks = [oracle.encode_block(local[i * bs:(i + 1) * bs].numpy(), 1) for i in range(hi - lo)]
        packed = torch.from_numpy(np.frombuffer(b"".join(toks) + b"\0", dtype=np.uint8).copy())
        lens = torch.tensor([len(t) for t in toks], dtype=torch.int32)
        stream, all_len, off = shard.gather_packed(packed, lens, nblk, dst=0)
        if rank == 0:
            want = b"".join(oracle.encode_block(full[i * bs:(i + 1) * bs].numpy(), 1) for i in range(nblk))
            assert stream.numpy().tobytes() == want and int(off[-1]) == len(want)
        else:
            assert stream is None
        mine, moff = shard.scatter_packed(stream, all_len, nblk, src=0, device=torch.device("cpu"))
        assert mine.numpy().tobytes() == b"".join(toks) and moff.tolist()[-1] == mine.numel()
        dec = bytearray()
        for i in range(hi - lo):
            st, out = oracle.decode_block(mine[int(moff[i]):int(moff[i + 1])].numpy(), bs)
            assert st == 0
            dec += out
        back = shard.gather_rows(torch.from_numpy(np.frombuffer(bytes(dec) + b"\0", dtype=np.uint8).copy()), nblk, bs, dst=0)
        ok = True if rank != 0 else bool(torch.equal(back, full))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nblk", [(2, 7), (3, 2), (3, 8)])
def test_funnel_scatter_gather(world, nblk):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_funnel_worker, args=(r, world, port, nblk, 1 << 14, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
