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


# ---- funnelling a batch through one rank (one exchange step in, one out) ----
#
# When the data lives on ONE GPU (a stream being written or read there) the blocks have to
# travel: raw blocks are scattered from the source rank, every rank encodes its contiguous
# range, the sizes are all-gathered (above) and the token streams are gathered to the sink
# rank straight into their place in the packed stream -- the sink posts one receive per rank
# at the offset the prefix sum gives, so nothing is padded or re-copied.  Decoding mirrors
# it.  Point-to-point sends over the process group: NVLink / NVSwitch under NCCL, sockets
# under gloo in the CPU tests.  No reduction is involved anywhere.

def _exchange(ops):
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def scatter_rows(full, rows_total, row_bytes, src=0, group=None, device=None):
    """Fixed-size rows (uncompressed blocks): rank `src` holds uint8[rows_total * row_bytes];
    every rank gets the rows of its block_range.  Returns uint8[rows_local * row_bytes]."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = block_range(rows_total, rank, world)
    if rank == src:
        device = full.device
    local = torch.empty((hi - lo) * row_bytes, dtype=torch.uint8, device=device)
    ops = []
    if rank == src:
        for r in range(world):
            a, b = block_range(rows_total, r, world)
            if r == rank:
                local.copy_(full[a * row_bytes:b * row_bytes])
            elif b > a:
                ops.append(dist.P2POp(dist.isend, full[a * row_bytes:b * row_bytes], r, group))
    elif hi > lo:
        ops.append(dist.P2POp(dist.irecv, local, src, group))
    _exchange(ops)
    return local


def gather_rows(local, rows_total, row_bytes, dst=0, group=None):
    """Inverse of scatter_rows: the decoded blocks return to rank `dst` in stream order
    (None on the other ranks)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = block_range(rows_total, rank, world)
    ops = []
    full = None
    if rank == dst:
        full = torch.empty(rows_total * row_bytes, dtype=torch.uint8, device=local.device)
        for r in range(world):
            a, b = block_range(rows_total, r, world)
            if r == rank:
                full[a * row_bytes:b * row_bytes].copy_(local[: (b - a) * row_bytes])
            elif b > a:
                ops.append(dist.P2POp(dist.irecv, full[a * row_bytes:b * row_bytes], r, group))
    elif hi > lo:
        ops.append(dist.P2POp(dist.isend, local[: (hi - lo) * row_bytes], dst, group))
    _exchange(ops)
    return full


def gather_packed(local_bytes, local_len, nblk_total, dst=0, group=None):
    """Variable-size token streams: every rank holds its blocks packed back to back
    (`local_bytes`, sizes `local_len`).  Rank `dst` receives the whole packed stream in
    block order.  Returns (stream or None, all_len, offsets) -- all_len / offsets on every rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    all_len = gather_block_lengths(local_len, nblk_total, group)
    off = stream_offsets(all_len)
    host_off = off.tolist()
    lo, hi = block_range(nblk_total, rank, world)
    mine = host_off[hi] - host_off[lo]
    ops = []
    stream = None
    if rank == dst:
        stream = torch.empty(host_off[-1], dtype=torch.uint8, device=local_bytes.device)
        for r in range(world):
            a, b = block_range(nblk_total, r, world)
            if r == rank:
                stream[host_off[a]:host_off[b]].copy_(local_bytes[:mine])
            elif host_off[b] > host_off[a]:
                ops.append(dist.P2POp(dist.irecv, stream[host_off[a]:host_off[b]], r, group))
    elif mine:
        ops.append(dist.P2POp(dist.isend, local_bytes[:mine], dst, group))
    _exchange(ops)
    return stream, all_len, off


def scatter_packed(stream, all_len, nblk_total, src=0, group=None, device=None):
    """Inverse of gather_packed: rank `src` holds the packed stream and `all_len` is known
    everywhere (it travels in the stream's chunk headers / index).  Every rank gets the bytes
    of its block_range.  Returns (local_bytes, local_offsets[int64, rows_local + 1])."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    off = stream_offsets(all_len)
    host_off = off.tolist()
    lo, hi = block_range(nblk_total, rank, world)
    if rank == src:
        device = stream.device
    local = torch.empty(host_off[hi] - host_off[lo], dtype=torch.uint8, device=device)
    ops = []
    if rank == src:
        for r in range(world):
            a, b = block_range(nblk_total, r, world)
            if r == rank:
                local.copy_(stream[host_off[a]:host_off[b]])
            elif host_off[b] > host_off[a]:
                ops.append(dist.P2POp(dist.isend, stream[host_off[a]:host_off[b]], r, group))
    elif local.numel():
        ops.append(dist.P2POp(dist.irecv, local, src, group))
    _exchange(ops)
    return local, (off[lo:hi + 1] - off[lo]).to(local.device)
