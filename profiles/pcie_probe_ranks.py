"""PCIe floor of the e2e step with N processes at once (torchrun): every rank moves the byte counts of
one bench step over PCIe from pinned memory -- H2D 4.3 GB + 1.41 GB, D2H 1.41 GB + 4.3 GB -- first one
direction at a time, then both directions concurrently, all ranks in lockstep (barrier before the
timed region, max over ranks).  What the e2e leg of bench.py cannot beat on this host."""
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
try:
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from minlz_b200 import _lib
    node = _lib.load().mzcu_bind_host_to_device(local)
except Exception:
    node = None
n_big, n_small = 4096 << 20, 1414 << 20
d_big = torch.empty(n_big, dtype=torch.uint8, device="cuda")
d_small = torch.empty(n_small, dtype=torch.uint8, device="cuda")
h_big = torch.empty(n_big, dtype=torch.uint8).pin_memory()
h_small = torch.empty(n_small, dtype=torch.uint8).pin_memory()
h_big2 = torch.empty(n_big, dtype=torch.uint8).pin_memory()
h_small2 = torch.empty(n_small, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) * 1e3


def h2d():
    d_big.copy_(h_big, non_blocking=True)
    d_small.copy_(h_small, non_blocking=True)


def d2h():
    h_small2.copy_(d_small, non_blocking=True)
    h_big2.copy_(d_big, non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        h2d()
    with torch.cuda.stream(s2):
        d2h()
    s1.synchronize()
    s2.synchronize()


a, b, c = timed(h2d), timed(d2h), timed(both)
if rank == 0:
    gb = (n_big + n_small) / 1e9
    print("ranks %d (numa node of rank 0: %s): H2D %.2f GB %.1f ms (%.1f GB/s per rank), D2H %.1f ms (%.1f GB/s), both directions %.1f ms"
          % (world, node, gb, a, gb / a * 1e3, b, gb / b * 1e3, c))
    print("e2e floor of one bench step per rank: serial directions %.1f ms -> %.1f GB/s per rank; full duplex %.1f ms -> %.1f GB/s per rank"
          % (a + b, n_big / (a + b) / 1e6, c, n_big / c / 1e6))
if world > 1:
    dist.destroy_process_group()
