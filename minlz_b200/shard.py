"""Sharding of independent blocks across the GPUs of a box.

MinLZ blocks are independent by format design (reference SPEC.md "Independent
block streams"; repeat offset and tables reset per block), so a batch shards
by contiguous block ranges: rank r owns blocks [r*N/G, (r+1)*N/G) and stream
order is rank order.  The data path needs no collective; the only exchange a
sharded stream writer needs is every rank's compressed block sizes (4 bytes
per block) so that each rank can place its chunks in the output stream -- one
all-gather, done here with torch.distributed (NCCL on GPUs, gloo in the CPU
tests).
"""
import torch
import torch.distributed as dist


def block_range(nblk, rank, world):
    """Contiguous, balanced block range [lo, hi) of `rank` (first ranks get the remainder)."""
    base, rem = divmod(nblk, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_block_lengths(local_len, nblk_total, group=None):
    """All-gather of per-block encoded sizes in stream order.

    local_len: int32 tensor with this rank's block sizes (its block_range).
    Returns an int32 tensor [nblk_total] on the same device, identical on all ranks.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = block_range(nblk_total, rank, world)
    assert local_len.numel() == hi - lo
    width = (nblk_total + world - 1) // world  # equal-size slots for the collective
    slot = torch.zeros(width, dtype=torch.int32, device=local_len.device)
    slot[: hi - lo] = local_len.to(torch.int32)
    out = torch.empty(world * width, dtype=torch.int32, device=local_len.device)
    dist.all_gather_into_tensor(out, slot, group=group)
    parts = []
    for r in range(world):
        a, b = block_range(nblk_total, r, world)
        parts.append(out[r * width: r * width + (b - a)])
    return torch.cat(parts)


def stream_offsets(all_len, per_block_overhead=0):
    """Exclusive prefix sum: byte offset of every block's chunk in the stream.
    `per_block_overhead` is the framing a stream writer adds per chunk (header
    + CRC = 8 bytes in the MinLZ stream format, writer.go:689-696)."""
    sizes = all_len.to(torch.int64) + per_block_overhead
    off = torch.zeros(all_len.numel() + 1, dtype=torch.int64, device=all_len.device)
    off[1:] = torch.cumsum(sizes, 0)
    return off
