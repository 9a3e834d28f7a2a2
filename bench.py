#!/usr/bin/env python
"""bench.py -- MinLZ block encode+decode throughput on B200 (one JSON line).

Workloads (BASELINE.json `configs`):

  blocks  (default; configs[1]+[2], the configuration the metric is quoted on)
          4096 x 1 MiB synthetic JSON-like blocks per GPU.  One step = LevelFastest encode of
          the whole batch (encode_l1 kernel + dense pack), then decode of the packed token
          streams (decode kernel), all through the C ABI of libminlz_cuda.so.
  stream  (configs[3]) 2048 x 2 MiB synthetic log-text blocks per GPU per step, the batch
          points of the stream layer: Writer.EncodeBuffer (writer.go:441: CRC-32C of every raw
          block + encode + ordered hand-off) and Reader.DecodeConcurrent (reader.go:575: decode
          + CRC of the decoded block).  A 32 GiB stream is 8 such steps per GPU at N=1.
  sweep   (configs[4]) LevelBalanced on 8 MiB blocks, mixed entropy: text, binary structs and
          uniform random (= pre-compressed; the encoder must answer "stored"), 512 blocks of each
          per GPU, the CPU arm timed beside every kind.

`value` = uncompressed bytes / (encode + decode time), data resident in HBM, CUDA events on
the launch stream; `e2e` = the same round trip through the host-pointer C ABI calls with
pinned host buffers (H2D and D2H inside the timed region).  With N > 1 ranks the default run
also reports `funnel`: the batch living on rank 0, scattered / gathered over NCCL every step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--workload blocks|stream|sweep] [--blocks B] [--block-size S] [--kind json] [--level L]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "uncompressed GB/s encode+decode, 1 MB blocks, 1/2/4/8 GPU vs host asm"
UNIT = "GB/s"
LEVEL_NAME = {-1: "LevelSuperFast (encode_l0)", 1: "LevelFastest (encode_l1)", 2: "LevelBalanced (encode_l2)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """Per-launch DRAM bytes of `kernel` from the committed ncu summary, or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _once(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            f = [x.strip() for x in out.strip().split(",")]
            self.samples.append(float(f[0]))
            self.max_mhz = float(f[1])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._once()
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=10)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------ CPU arm ----

def cpu_engine():
    """The CPU arm: the reference's own AMD64 assembly through oracle/_ref when that
    library is here (built in the dev container from /root/reference, travels with
    the snapshot), else the oracle port of the pure-Go path."""
    from oracle import refasm
    if refasm.build() is not None:
        return refasm, "reference", ("minio/minlz asm_amd64.s (encodeBlockAsm* / encodeBetterBlockAsm* / decodeBlockAsm) "
                                     "run natively via oracle/_ref, one block per task over a pthread pool")
    from oracle import binding
    binding.build()
    return binding, "port", "oracle port of the pure-Go path (oracle/_ref not built)"


class CpuArm:
    """Encode + decode of [n, bs] host blocks on all host cores.  Every buffer is allocated
    ONCE and every page touched before anything is timed, and the buffers are reused across
    passes -- the reference arm gets the same treatment as the GPU arm's pinned, warmed
    buffers (no first-touch page faults inside a timed pass)."""

    def __init__(self, n, bs, level, nthreads):
        import numpy as np
        self.np = np
        self.engine, self.kind, self.what = cpu_engine()
        self.n, self.bs, self.level, self.nthreads = n, bs, level, nthreads
        self.cap = bs + 16
        self.soff = np.arange(n + 1, dtype=np.uint64) * bs
        self.doff = np.arange(n + 1, dtype=np.uint64) * self.cap
        self.enc = np.empty(n * self.cap, dtype=np.uint8)
        self.comp = np.empty(n * bs + 64, dtype=np.uint8)
        self.dec = np.empty(n * bs, dtype=np.uint8)
        for a in (self.enc, self.comp, self.dec):
            a.fill(1)  # touch every page

    def encode(self, src):
        t0 = time.perf_counter()
        out_len = self.engine.encode_batch_mt(self.level, src, self.soff, self.enc, self.doff, self.nthreads)
        return time.perf_counter() - t0, out_len

    def compact(self, out_len):
        np = self.np
        coff = np.zeros(self.n + 1, dtype=np.uint64)
        np.cumsum(out_len, out=coff[1:])
        for i in range(self.n):
            a = int(self.doff[i])
            self.comp[int(coff[i]):int(coff[i + 1])] = self.enc[a:a + int(out_len[i])]
        return coff

    def decode(self, coff, live):
        """Decodes the blocks with a token stream (`live`); stored blocks have none."""
        np = self.np
        if live.all():
            soff, doff = coff, self.soff
        else:  # only the compressible blocks reach minLZDecode; stored ones are a memcpy in the wrapper
            idx = np.nonzero(live)[0]
            if idx.size == 0:
                return 0.0
            return self._decode_ranges(coff, idx)
        t0 = time.perf_counter()
        status = self.engine.decode_batch_mt(self.comp, soff, self.dec, doff, self.nthreads)
        dt = time.perf_counter() - t0
        assert not status.any()
        return dt

    def _decode_ranges(self, coff, idx):
        # contiguous re-pack of the live blocks (not timed), then one batch call
        np = self.np
        sizes = (coff[1:] - coff[:-1])[idx]
        so = np.zeros(idx.size + 1, dtype=np.uint64)
        np.cumsum(sizes, out=so[1:])
        tmp = np.empty(int(so[-1]) + 64, dtype=np.uint8)
        for k, i in enumerate(idx):
            tmp[int(so[k]):int(so[k + 1])] = self.comp[int(coff[i]):int(coff[i + 1])]
        do = np.arange(idx.size + 1, dtype=np.uint64) * self.bs
        t0 = time.perf_counter()
        status = self.engine.decode_batch_mt(tmp, so, self.dec, do, self.nthreads)
        dt = time.perf_counter() - t0
        assert not status.any()
        return dt

    def round_trip(self, host_blocks, check=True):
        np = self.np
        src = host_blocks.reshape(-1)
        te, out_len = self.encode(src)
        coff = self.compact(out_len)
        live = out_len > 0
        td = self.decode(coff, live)
        if check and live.all():
            assert np.array_equal(self.dec, src), "CPU arm round trip mismatch"
        return {"enc_s": te, "dec_s": td, "seconds": te + td, "stored": int((~live).sum()),
                "comp_bytes": int(coff[-1])}


def cpu_best(host_blocks, level, nthreads, passes=2, warm=1):
    """Best of `passes` warmed passes; GB/s of uncompressed bytes."""
    n, bs = host_blocks.shape
    arm = CpuArm(n, bs, level, nthreads)
    for _ in range(warm):
        arm.round_trip(host_blocks)
    best_e = best_d = 1e30
    r = None
    for _ in range(passes):
        r = arm.round_trip(host_blocks, check=False)
        best_e, best_d = min(best_e, r["enc_s"]), min(best_d, r["dec_s"])
    total = n * bs
    return {"value": total / (best_e + best_d) / 1e9, "encode_gbps": total / best_e / 1e9,
            "decode_gbps": (total - r["stored"] * bs) / best_d / 1e9 if best_d > 0 else None,
            "stored_blocks": r["stored"], "kind": arm.kind, "what": arm.what}


# ------------------------------------------------------------ workloads ----

def resolve_workload(args):
    """Fills the defaults of the chosen workload; returns the list of (kind, level) legs."""
    wl = args.workload
    if wl == "blocks":
        args.kind = args.kind or "json"
        args.block_size = args.block_size or (1 << 20)
        args.blocks = args.blocks or 4096
        args.level = 1 if args.level is None else args.level
        legs = [args.kind]
    elif wl == "stream":
        args.kind = args.kind or "log"
        args.block_size = args.block_size or (2 << 20)
        args.blocks = args.blocks or 2048
        args.level = 1 if args.level is None else args.level
        legs = [args.kind]
    else:  # sweep
        args.block_size = args.block_size or (8 << 20)
        args.blocks = args.blocks or 512
        args.level = 2 if args.level is None else args.level
        legs = [args.kind] if args.kind else ["text", "binary", "random"]
        args.kind = "+".join(legs)
    if args.flavor == "auto":
        # the flavour the reference arm executes on this box (its amd64 assembly)
        args.flavor = "amd64"
    return legs


def workload_config(args):
    what = {
        "blocks": "configs[1]+[2]: %d x %d B synthetic %s blocks per GPU, %s encode then batched decode of the packed token streams",
        "stream": "configs[3]: stream batch points (Writer.EncodeBuffer / Reader.DecodeConcurrent): %d x %d B synthetic %s blocks per GPU "
                  "per step, CRC-32C of every raw block + %s encode + pack, then decode + CRC of the decoded blocks; a 32 GiB stream "
                  "is 8 such steps per GPU at N=1",
        "sweep": "configs[4]: mixed-entropy sweep, %d x %d B blocks per kind per GPU, kinds %s, %s encode then decode of the "
                 "compressible blocks (incompressible ones are stored by the wrapper, encode.go:137-138)",
    }[args.workload] % (args.blocks, args.block_size, args.kind, LEVEL_NAME[args.level])
    return {"workload": what, "blocks_per_gpu": args.blocks, "block_size": args.block_size, "level": args.level,
            "encoder_flavor": args.flavor,
            "cache": "inputs (%.1f GB per pass) larger than the 126 MB L2, no flush needed" %
                     (args.blocks * args.block_size / 1e9)}


def run_reference(args, legs):
    """--impl reference: the reference's CPU implementation of the path on the box's host
    cores, all of them: its own AMD64 assembly via oracle/_ref (see cpu_engine), falling
    back to the oracle port only if that library is absent.  Buffers are allocated and
    touched once, `--warmup` full passes run untimed, then K timed passes."""
    import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    bs = args.block_size
    nsample = min(args.blocks, max(1, args.cpu_bytes // bs))
    t_tot = 0.0
    per_leg = {}
    kind = what = None
    for leg in legs:
        blocks = synth.make_blocks(leg, nsample, bs, device="cpu").numpy()
        arm = CpuArm(nsample, bs, args.level, cores)
        kind, what = arm.kind, arm.what
        for _ in range(max(args.warmup, 1)):
            arm.round_trip(blocks)
        te = td = 0.0
        stored = 0
        for _ in range(args.steps):
            r = arm.round_trip(blocks, check=False)
            te += r["enc_s"]
            td += r["dec_s"]
            stored = r["stored"]
        t_tot += te + td
        tb = nsample * bs * args.steps
        per_leg[leg] = {"value": round(tb / (te + td) / 1e9, 4), "encode_gbps": round(tb / te / 1e9, 4),
                        "decode_gbps": round((tb - stored * bs * args.steps) / td / 1e9, 4) if td > 0 else None,
                        "stored_blocks": stored}
    value = len(legs) * nsample * bs * args.steps / t_tot / 1e9
    first = per_leg[legs[0]]
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": round(t_tot / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d x %d B blocks of %s per step, level %d encode + decode, buffers allocated and touched "
                                   "once, %d untimed warm-up passes; %s" % (nsample, bs, args.kind, args.level,
                                                                            max(args.warmup, 1), what),
                         "encode_gbps": first["encode_gbps"], "decode_gbps": first["decode_gbps"]},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if len(legs) > 1:
        line["sweep"] = per_leg
    print(json.dumps(line))


def parity_gate_all(torch, np, args, mz, src, nblk, bs, comp, coff, out_len, kind_name):
    """Encoder bytes of EVERY block against the checker (never timed): the reference's real
    assembly when oracle/_ref is on the box and the flavour is amd64, else the oracle
    restatement of the flavour.  The CPU encodes the batch in chunks on all cores."""
    from oracle import binding as oracle_port
    ref_engine, ref_kind, _ = cpu_engine()
    use_ref = args.flavor == "amd64" and ref_kind == "reference"
    who = "the reference's amd64 assembly (oracle/_ref)" if use_ref else "the oracle restatement of the %s flavour" % args.flavor
    cores = os.cpu_count() or 1
    chunk = max(1, min(nblk, (1 << 30) // bs))
    cap = bs + 16
    enc = np.empty(chunk * cap, dtype=np.uint8)
    h_len = out_len.cpu().numpy().astype(np.int64)
    h_coff = coff.cpu().numpy().astype(np.int64)
    bad = []
    for lo in range(0, nblk, chunk):
        hi = min(nblk, lo + chunk)
        m = hi - lo
        hb = src[lo * bs:hi * bs].cpu().numpy()
        soff = np.arange(m + 1, dtype=np.uint64) * bs
        doff = np.arange(m + 1, dtype=np.uint64) * cap
        if use_ref:
            want_len = ref_engine.encode_batch_mt(args.level, hb, soff, enc, doff, cores)
        else:
            want_len = np.zeros(m, dtype=np.uint32)
            for i in range(m):
                w = oracle_port.encode_block(hb[i * bs:(i + 1) * bs], args.level, flavor="asm" if args.flavor == "amd64" else "go")
                enc[i * cap:i * cap + len(w)] = np.frombuffer(w, dtype=np.uint8)
                want_len[i] = len(w)
        got = comp[int(h_coff[lo]):int(h_coff[hi])].cpu().numpy()
        for i in range(m):
            g = got[int(h_coff[lo + i] - h_coff[lo]):int(h_coff[lo + i + 1] - h_coff[lo])]
            if int(want_len[i]) != int(h_len[lo + i]) or not np.array_equal(g, enc[i * cap:i * cap + int(want_len[i])]):
                bad.append(lo + i)
    assert not bad, "encoder bytes differ from %s on %s blocks %s" % (who, kind_name, bad[:10])
    return "encoder bytes identical to %s on %d/%d %s blocks" % (who, nblk, nblk, kind_name)


def run_funnel(args, mz, shard, dist, world, rank, dev, src, soff, enc, eoff, out_len, comp, coff, dec, status, steps):
    """The batch lives on rank 0: scatter raw blocks -> encode -> gather the packed stream to rank 0 ->
    scatter it back -> decode -> gather the blocks on rank 0.  Everything inside the timed region."""
    import torch
    nblk, bs = args.blocks, args.block_size
    total = world * nblk
    parts = [torch.empty_like(src) for _ in range(world)] if rank == 0 else None
    dist.gather(src, parts, dst=0)                      # build the full batch on rank 0 (not timed)
    full = torch.cat(parts) if rank == 0 else None
    del parts

    def step():
        local = shard.scatter_rows(full, total, bs, src=0, device=dev)
        mz.encode_blocks_dev(local, soff, enc, eoff, out_len, args.level)
        mz.pack_blocks_dev(enc, eoff, out_len, comp, coff)
        stream, all_len, _ = shard.gather_packed(comp, out_len, total, dst=0)
        mine, moff = shard.scatter_packed(stream, all_len, total, src=0, device=dev)
        mz.decode_blocks_dev(mine, moff, dec, soff, status)
        return shard.gather_rows(dec, total, bs, dst=0), stream

    for _ in range(2):
        back, stream = step()
    torch.cuda.synchronize()
    if rank == 0:
        assert torch.equal(back, full), "funnel round trip mismatch"
    assert int(status.abs().sum()) == 0
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        back, stream = step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = None
    if rank == 0:
        ms = float(t[0]) / steps
        sb = int(stream.numel())
        raw_out = (world - 1) * nblk * bs          # rank 0 egress: raw blocks of the other ranks
        res = {"value": round(total * bs / (ms * 1e-3) / 1e9, 4), "unit": UNIT, "ms_per_step": round(ms, 4), "steps": steps,
               "blocks_total": total,
               "what": "the batch of %d blocks lives on rank 0: raw blocks scattered, packed token streams gathered, scattered "
                       "back and decoded blocks gathered with point-to-point NCCL sends inside every step" % total,
               "nccl_bytes_per_step": int(2 * raw_out + 2 * (sb * (world - 1)) // world),
               "rank0_egress_bytes_per_step": int(raw_out + (sb * (world - 1)) // world),
               "rank0_link_floor_ms": round((raw_out + (sb * (world - 1)) // world) / 770e9 * 1e3 * 2, 2),
               "ratio": round(total * bs / sb, 4)}
    del full
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="blocks", choices=["blocks", "stream", "sweep"])
    ap.add_argument("--blocks", type=int, default=None)
    ap.add_argument("--block-size", type=int, default=None)
    ap.add_argument("--kind", default=None)
    ap.add_argument("--level", type=int, default=None, choices=(-1, 1, 2),
                    help="1 = LevelFastest (headline), 2 = LevelBalanced, -1 = LevelSuperFast")
    ap.add_argument("--flavor", default="auto", choices=("auto", "go", "amd64"),
                    help="which reference build the encoder mirrors byte for byte: the amd64 assembly (what the "
                         "reference arm runs on this box; default) or the pure-Go functions")
    ap.add_argument("--funnel", action="store_true", help="N > 1: only run the funnel measurement and print its line")
    ap.add_argument("--no-funnel", action="store_true", help="N > 1: skip the funnel sub-measurement")
    ap.add_argument("--cpu-bytes", type=int, default=1 << 30, help="bounded sample (bytes per kind) for the CPU legs")
    ap.add_argument("--e2e-blocks", type=int, default=None, help="blocks per e2e step (host buffers)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e: blocking calls only")
    ap.add_argument("--pipeline-ahead", default="1,2", help="e2e: encode jobs submitted ahead of the oldest decode (comma list)")
    ap.add_argument("--pipeline-iters", type=int, default=0, help="e2e: batches per pipelined measurement (0: from --steps)")
    ap.add_argument("--no-gate", action="store_true", help="skip the all-blocks encoder byte comparison (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    legs = resolve_workload(args)

    if args.impl == "reference":
        run_reference(args, legs)
        return

    import numpy as np
    import torch
    import minlz_b200 as mz
    from minlz_b200 import _lib, shard
    import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    lib.mzcu_bind_host_to_device(local)  # NUMA: this process's threads and pinned buffers next to its GPU
    nblk, bs = args.blocks, args.block_size
    mz.set_encoder_flavor(mz.FlavorAMD64 if args.flavor == "amd64" else mz.FlavorGo)
    with_crc = args.workload == "stream"

    soff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * bs
    cap = bs + 16
    eoff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * cap
    enc = torch.empty(nblk * cap, dtype=torch.uint8, device=dev)
    out_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
    comp = torch.empty(nblk * bs + 64, dtype=torch.uint8, device=dev)
    coff = torch.zeros(nblk + 1, dtype=torch.int64, device=dev)
    dec = torch.empty(nblk * bs, dtype=torch.uint8, device=dev)
    status = torch.zeros(nblk, dtype=torch.int32, device=dev)
    crc_a = torch.zeros(nblk, dtype=torch.int32, device=dev)
    crc_b = torch.zeros(nblk, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def crc_dev(buf, off, out):
        r = lib.mzcu_crc32c_blocks_dev(local, nblk, buf.data_ptr(), off.data_ptr(), out.data_ptr(), sptr)
        assert r == 0, lib.mzcu_last_error()

    K = args.steps
    U = nblk * bs
    legs_out = {}
    agg = {"total_ms": 0.0, "enc": 0.0, "pack": 0.0, "dec": 0.0, "comp": 0, "bytes": 0}
    clocks_all = None
    parity_notes = []
    src = None
    for leg in legs:
        src = None
        # ---- synthetic input, resident in HBM; independent blocks shard by rank
        src = synth.make_blocks(leg, nblk, bs, device=dev, first=rank * nblk).reshape(-1)
        stored_leg = leg == "random"

        def step(timed):
            if timed is not None:
                ev[0].record(stream)
            if with_crc:
                crc_dev(src, soff, crc_a)  # writer.go:672
            mz.encode_blocks_dev(src, soff, enc, eoff, out_len, args.level)
            if timed is not None:
                ev[1].record(stream)
            mz.pack_blocks_dev(enc, eoff, out_len, comp, coff)
            if world > 1:
                # the one real exchange of the sharded stream: every rank learns all
                # compressed block sizes (stream order = rank order), 4 B per block
                all_len = shard.gather_block_lengths(out_len, world * nblk)
                shard.stream_offsets(all_len, per_block_overhead=8)
            if timed is not None:
                ev[2].record(stream)
            if not stored_leg:
                mz.decode_blocks_dev(comp, coff, dec, soff, status)
                if with_crc:
                    crc_dev(dec, soff, crc_b)  # reader.go:341-351
            if timed is not None:
                ev[3].record(stream)
                torch.cuda.synchronize()
                timed["enc"] += ev[0].elapsed_time(ev[1])
                timed["pack"] += ev[1].elapsed_time(ev[2])
                timed["dec"] += ev[2].elapsed_time(ev[3])

        if args.funnel and world > 1:
            res = run_funnel(args, mz, shard, dist, world, rank, dev, src, soff, enc, eoff, out_len, comp, coff, dec, status, K)
            if rank == 0:
                cfg = workload_config(args)
                cfg["workload"] += "; FUNNEL: " + res["what"]
                print(json.dumps({"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": K,
                                  "warmup": 2, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                                  "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg, "funnel": res}))
            dist.destroy_process_group()
            return

        for _ in range(max(args.warmup, 1)):
            step(None)
        torch.cuda.synchronize()
        # correctness gate before timing: the round trip must reproduce the input
        if stored_leg:
            assert int((out_len != 0).sum()) == 0, "incompressible blocks must come back as 0 = stored"
        else:
            assert int(status.abs().sum()) == 0, "decode reported corrupt blocks"
            assert int((out_len <= 0).sum()) == 0, "encoder returned incompressible on compressible data"
            assert torch.equal(dec, src), "round trip mismatch"
            if with_crc:
                assert torch.equal(crc_a, crc_b), "CRC of decoded blocks differs from CRC of the raw blocks"
        comp_bytes = int(coff[-1])
        note = "round trip verified bit-exact before timing" if not stored_leg else "all blocks answered 0 = stored"
        if rank == 0 and not args.no_cpu and not args.no_gate:
            note = parity_gate_all(torch, np, args, mz, src, nblk, bs, comp, coff, out_len, leg) + "; " + note
        parity_notes.append(note)

        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        parts = {"enc": 0.0, "pack": 0.0, "dec": 0.0}
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
            t_start.record(stream)
            for _ in range(K):
                step(parts)
            t_end.record(stream)
            torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        tms = torch.tensor([t_start.elapsed_time(t_end), parts["enc"], parts["pack"], parts["dec"]], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        total_ms, enc_ms, pack_ms, dec_ms = [float(x) for x in tms.tolist()]
        cs = clocks.summary()
        if clocks_all is None:
            clocks_all = cs
        else:
            clocks_all["reasons"] = sorted(set(clocks_all["reasons"]) | set(cs["reasons"]))
            clocks_all["samples"] += cs["samples"]
            if cs["sm_mhz"] is not None and (clocks_all["sm_mhz"] is None or cs["sm_mhz"] < clocks_all["sm_mhz"]):
                clocks_all["sm_mhz"] = cs["sm_mhz"]
        legs_out[leg] = {"value": round(world * U / (total_ms / K * 1e-3) / 1e9, 4),
                         "encode_gbps": round(world * U / (enc_ms / K * 1e-3) / 1e9, 3),
                         "decode_gbps": round(world * U / (dec_ms / K * 1e-3) / 1e9, 3) if not stored_leg else None,
                         "ms": {"encode": round(enc_ms / K, 4), "pack": round(pack_ms / K, 4), "decode": round(dec_ms / K, 4)},
                         "ratio": round(U / comp_bytes, 4) if comp_bytes else None,
                         "stored_blocks": nblk if stored_leg else 0}
        agg["total_ms"] += total_ms
        agg["enc"] += enc_ms
        agg["pack"] += pack_ms
        agg["dec"] += dec_ms
        agg["comp"] += comp_bytes
        agg["bytes"] += U

    total_ms, enc_ms, pack_ms, dec_ms, comp_bytes = agg["total_ms"], agg["enc"], agg["pack"], agg["dec"], agg["comp"]
    UB = agg["bytes"]                      # uncompressed bytes of one step over all legs
    step_ms = total_ms / K
    value = world * UB / (step_ms * 1e-3) / 1e9

    # ---- e2e: host-pointer C ABI calls with pinned buffers (last leg's data for a sweep) -----------------
    e2e = None
    if not args.no_e2e:
        nb = min(nblk, args.e2e_blocks or nblk)
        h_src = torch.empty(nb * bs, dtype=torch.uint8).pin_memory()
        h_dec = torch.empty(nb * bs, dtype=torch.uint8).pin_memory()
        h_comp = torch.empty(nb * bs + 64, dtype=torch.uint8).pin_memory()
        n_src, n_dec, n_comp = h_src.numpy(), h_dec.numpy(), h_comp.numpy()
        hs = np.arange(nb + 1, dtype=np.uint64) * bs
        hc = np.zeros(nb + 1, dtype=np.uint64)
        hst = np.zeros(nb, dtype=np.int32)
        hcrc_a = np.zeros(nb, dtype=np.uint32)
        hcrc_b = np.zeros(nb, dtype=np.uint32)
        e2e_tot = {"enc": 0.0, "dec": 0.0, "t": 0.0, "cb": 0, "bytes": 0}
        extra = []
        aheads = sorted({max(1, int(a)) for a in args.pipeline_ahead.split(",")})
        pipe_t = {a: 0.0 for a in aheads}
        pipe_bytes = {a: 0 for a in aheads}
        api = ("mzcu_stream_encode_blocks + mzcu_stream_decode_blocks (host pointers, pinned, CRC-32C on the device)" if with_crc
               else "mzcu_encode_blocks_packed + mzcu_decode_blocks (host pointers, pinned)")
        for leg in legs:
            if len(legs) > 1:
                src = None
                src = synth.make_blocks(leg, nblk, bs, device=dev, first=rank * nblk).reshape(-1)
            h_src.copy_(src[: nb * bs])
            stored_leg = leg == "random"
            parts2 = {"enc": 0.0, "dec": 0.0}

            def e2e_step():
                # host blocks -> packed token streams on the host -> host blocks again
                t0_ = time.perf_counter()
                if with_crc:
                    r_ = lib.mzcu_stream_encode_blocks(local, args.level, nb, n_src.ctypes.data, hs.ctypes.data, n_comp.ctypes.data,
                                                       n_comp.size, hc.ctypes.data, hcrc_a.ctypes.data)
                    assert r_ == 0, lib.mzcu_last_error()
                    cb_ = int(hc[nb])
                else:
                    cb_ = mz.encode_blocks_packed_into(n_src, hs, n_comp, hc, args.level, device=local)
                t1_ = time.perf_counter()
                if not stored_leg:
                    if with_crc:
                        r_ = lib.mzcu_stream_decode_blocks(local, nb, n_comp.ctypes.data, hc.ctypes.data, n_dec.ctypes.data,
                                                           hs.ctypes.data, hst.ctypes.data, hcrc_b.ctypes.data)
                        assert r_ == 0, lib.mzcu_last_error()
                    else:
                        mz.decode_blocks_into(n_comp, hc, n_dec, hs, hst, device=local)
                parts2["enc"] += t1_ - t0_
                parts2["dec"] += time.perf_counter() - t1_
                return cb_

            for _ in range(min(max(args.warmup, 1), 2)):
                e2e_step()
            if not stored_leg:
                assert np.array_equal(n_dec, n_src) and not hst.any()
                if with_crc:
                    assert np.array_equal(hcrc_a, hcrc_b)
            if dist is not None:
                dist.barrier()
            parts2["enc"] = parts2["dec"] = 0.0
            t0 = time.perf_counter()
            ek = max(1, min(K, 3))
            for _ in range(ek):
                cb = e2e_step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / ek
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_tot["t"] += float(tt[0])
            e2e_tot["enc"] += parts2["enc"] / ek
            e2e_tot["dec"] += parts2["dec"] / ek
            e2e_tot["cb"] += cb
            e2e_tot["bytes"] += nb * bs
            legs_out[leg]["e2e_gbps"] = round(world * nb * bs / float(tt[0]) / 1e9, 4)

            # the same round trip with the asynchronous calls (mzcu_submit_* / mzcu_wait): `ahead` encode
            # jobs are in flight before the decode of the oldest one is submitted, so the device always has
            # the next kernel queued (launch order = submission order) and the download of one call overlaps
            # the upload and the kernels of the next.  ahead + 1 packed-stream buffers; one decoded buffer
            # (decode k has been waited for before decode k+1 is submitted).
            if not stored_leg and not args.no_pipeline:
                need = cb + 64
                while len(extra) < aheads[-1] or any(x.size < need for x in extra):
                    extra = [x for x in extra if x.size >= need]
                    extra.append(torch.empty(need, dtype=torch.uint8).pin_memory().numpy())
                comps = [n_comp] + extra[: aheads[-1]]
                hcs = [hc] + [np.zeros(nb + 1, dtype=np.uint64) for _ in range(aheads[-1])]
                crcs_a = [hcrc_a] + [np.zeros(nb, dtype=np.uint32) for _ in range(aheads[-1])]
                crcs_b = [hcrc_b] + [np.zeros(nb, dtype=np.uint32) for _ in range(aheads[-1])]
                sts = [hst] + [np.zeros(nb, dtype=np.int32) for _ in range(aheads[-1])]

                def sub_enc(i, nbuf):
                    b_ = i % nbuf
                    j_ = lib.mzcu_submit_stream_encode_blocks(local, args.level, nb, n_src.ctypes.data, hs.ctypes.data,
                                                              comps[b_].ctypes.data, comps[b_].size, hcs[b_].ctypes.data,
                                                              crcs_a[b_].ctypes.data if with_crc else None)
                    assert j_ > 0, lib.mzcu_last_error()
                    return j_

                def sub_dec(i, nbuf):
                    b_ = i % nbuf
                    j_ = lib.mzcu_submit_stream_decode_blocks(local, nb, comps[b_].ctypes.data, hcs[b_].ctypes.data,
                                                              n_dec.ctypes.data, hs.ctypes.data, sts[b_].ctypes.data,
                                                              crcs_b[b_].ctypes.data if with_crc else None)
                    assert j_ > 0, lib.mzcu_last_error()
                    return j_

                def pipeline(nit, ahead):
                    nbuf = ahead + 1
                    je_ = {i: sub_enc(i, nbuf) for i in range(min(ahead, nit))}
                    for i in range(nit):
                        assert lib.mzcu_wait(je_.pop(i)) == 0, lib.mzcu_last_error()
                        jd_ = sub_dec(i, nbuf)
                        if i + ahead < nit:
                            je_[i + ahead] = sub_enc(i + ahead, nbuf)
                        assert lib.mzcu_wait(jd_) == 0, lib.mzcu_last_error()

                for ahead in aheads:
                    n_dec[:] = 0
                    pipeline(ahead + 1, ahead)
                    assert np.array_equal(n_dec, n_src) and not any(s_.any() for s_ in sts[: ahead + 1])
                    if with_crc:
                        assert all(np.array_equal(hcrc_a, c_) for c_ in crcs_a[: ahead + 1] + crcs_b[: ahead + 1])
                    if dist is not None:
                        dist.barrier()
                    pk = args.pipeline_iters or 4 * (ahead + 1)  # 8 / 12 batches of 4 GiB: a 32 - 48 GiB stream
                    t0 = time.perf_counter()
                    pipeline(pk, ahead)
                    dtp = (time.perf_counter() - t0) / pk
                    tp_ = torch.tensor([dtp], dtype=torch.float64, device=dev)
                    if dist is not None:
                        dist.all_reduce(tp_, op=dist.ReduceOp.MAX)
                    pipe_t[ahead] += float(tp_[0])
                    pipe_bytes[ahead] += nb * bs
        nl = len(legs)
        dec_legs = sum(1 for leg in legs if leg != "random")
        serial_v = round(world * e2e_tot["bytes"] / e2e_tot["t"] / 1e9, 4)
        pipe_vs = {a: round(world * pipe_bytes[a] / pipe_t[a] / 1e9, 4)
                   for a in pipe_t if pipe_t[a] > 0 and pipe_bytes[a] == e2e_tot["bytes"]}
        ahead_best = max(pipe_vs, key=pipe_vs.get) if pipe_vs else None
        pipe_v = pipe_vs[ahead_best] if pipe_vs else None
        e2e = {"value": pipe_v if pipe_v and pipe_v > serial_v else serial_v, "unit": UNIT,
               "serial_calls": serial_v, "pipelined_calls": pipe_v,
               "pipelined_by_jobs_in_flight": {str(a + 1): v for a, v in pipe_vs.items()},
               "how": "value = the best of: one blocking encode call then one blocking decode call per step (serial_calls); "
                      "the same calls submitted asynchronously (mzcu_submit_* / mzcu_wait) with two or three jobs in "
                      "flight - the decode of batch k overlaps the encodes of batches k+1 (and k+2) - ramp-up and drain "
                      "inside the timed region (pipelined_calls).  All of them move every input and output byte over PCIe "
                      "inside the timed region.",
               "h2d_bytes_per_step": int(e2e_tot["bytes"] + e2e_tot["cb"] + nl * 2 * 8 * (nb + 1) * 2),
               "d2h_bytes_per_step": int(e2e_tot["cb"] + dec_legs * nb * bs + nl * 8 * nb),
               "blocks_per_step": nb * nl,
               "ms_per_step": round((pipe_t[ahead_best] if pipe_v and pipe_v > serial_v else e2e_tot["t"]) * 1e3, 3),
               "serial_ms_per_step": round(e2e_tot["t"] * 1e3, 3),
               "encode_call_ms": round(e2e_tot["enc"] * 1e3, 3), "decode_call_ms": round(e2e_tot["dec"] * 1e3, 3),
               "api": api}

    # ---- funnel (N > 1): the scatter / gather path of SURVEY 8e, inside the default line --------
    funnel = None
    if world > 1 and not args.no_funnel and args.workload == "blocks":
        funnel = run_funnel(args, mz, shard, dist, world, rank, dev, src, soff, enc, eoff, out_len, comp, coff, dec, status,
                            max(1, min(K, 3)))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    enc_bytes = UB + comp_bytes          # algorithmic: read source once, write tokens once
    dec_bytes = comp_bytes + (UB - sum(U for leg in legs if leg == "random"))  # read tokens once, write output once
    enc_gbs = enc_bytes / (enc_ms / K * 1e-3) / 1e9
    dec_gbs = dec_bytes / (dec_ms / K * 1e-3) / 1e9 if dec_ms > 0 else 0.0
    dominant_is_enc = enc_ms >= dec_ms
    enc_kernel = {-1: "encode_l1_kernel<true> (L0 params)", 1: "encode_l1_kernel<false>", 2: "encode_l2_kernel"}[args.level]
    if args.flavor == "amd64":
        enc_kernel = enc_kernel.replace("encode_l1_kernel", "encode_l1_asm_kernel").replace("encode_l2_kernel", "encode_l2_asm_kernel")
    headline = args.workload == "blocks" and args.level == 1 and nblk == 4096 and bs == (1 << 20)
    enc_traffic = ncu_traffic("encode") if headline else None  # the ncu captures are of the headline batch
    dec_traffic = ncu_traffic("decode") if headline else None

    def roofline(kernel, gbs, traffic, alg, ms):
        r = {"bound": "hbm", "kernel": kernel, "achieved": round(gbs, 3), "peak": peak, "unit": "GB/s",
             "frac": round(gbs / peak, 5), "traffic": traffic, "peak_source": peak_src,
             "algorithmic_bytes_per_launch": alg // max(1, len(legs)), "share_of_step": round(ms / total_ms, 4)}
        if traffic:
            # what the kernel actually moves: table probes are random 32-byte reads and the B200 L2 fetches a
            # whole 128-byte line per miss (profiles/r01_micro_gather32.txt: 4.7 TB/s ceiling for such gathers)
            r["traffic_gbs"] = round(traffic / (ms / K * 1e-3) / 1e9, 1)
            r["traffic_frac_of_peak"] = round(r["traffic_gbs"] / peak, 4)
            r["traffic_over_algorithmic"] = round(traffic / max(1, alg), 2)
        return r

    roof_enc = roofline(enc_kernel, enc_gbs, enc_traffic, enc_bytes, enc_ms)
    roof_dec = roofline("decode_pc_kernel", dec_gbs, dec_traffic, dec_bytes, dec_ms)
    roof = roof_enc if dominant_is_enc else roof_dec

    cpu = None
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        nsample = min(nblk, max(1, args.cpu_bytes // bs))
        t_cpu = 0.0
        first = None
        for leg in legs:
            if len(legs) > 1 or src is None:
                hb = synth.make_blocks(leg, nsample, bs, device="cpu", first=rank * nblk).numpy()
            else:
                hb = src[: nsample * bs].cpu().numpy().reshape(nsample, bs)
            r = cpu_best(hb, args.level, cores, passes=2, warm=1)
            t_cpu += nsample * bs / (r["value"] * 1e9)
            legs_out[leg]["cpu_gbps"] = round(r["value"], 4)
            legs_out[leg]["cpu_encode_gbps"] = round(r["encode_gbps"], 4)
            legs_out[leg]["cpu_decode_gbps"] = round(r["decode_gbps"], 4) if r["decode_gbps"] else None
            first = first or r
        cpu = {"value": round(len(legs) * nsample * bs / t_cpu / 1e9, 4), "unit": UNIT, "cores": cores, "kind": first["kind"],
               "sample": "first %d of the %d blocks%s, level %d encode + decode, buffers touched and one warm-up pass before the "
                         "best of 2 timed passes; %s" % (nsample, nblk, " of every kind" if len(legs) > 1 else "", args.level,
                                                         first["what"]),
               "encode_gbps": round(first["encode_gbps"], 4),
               "decode_gbps": round(first["decode_gbps"], 4) if first["decode_gbps"] else None}

    # per step and leg: [crc,] encode, scan_lengths, pack_blocks and, unless the leg is stored, decode [, crc]
    launches_per_step = sum(3 + (1 if with_crc else 0) + (0 if leg == "random" else 1 + (1 if with_crc else 0)) for leg in legs)
    line = {
        "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 1),
        "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(args),
        "encode_gbps": round(world * UB / (enc_ms / K * 1e-3) / 1e9, 3),
        "decode_gbps": round(world * (dec_bytes - comp_bytes) / (dec_ms / K * 1e-3) / 1e9, 3) if dec_ms > 0 else None,
        "ms": {"encode": round(enc_ms / K, 4), "pack": round(pack_ms / K, 4), "decode": round(dec_ms / K, 4)},
        "ratio": round(UB / comp_bytes, 4) if comp_bytes else None,
        "roofline": roof, "roofline_decode": roof_dec, "roofline_encode": roof_enc,
        "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks_all,
        "gpu_launches": K * launches_per_step,
        "parity": "; ".join(parity_notes),
    }
    if len(legs) > 1:
        line["sweep"] = legs_out
    if funnel is not None:
        line["funnel"] = funnel
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
