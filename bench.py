#!/usr/bin/env python
"""bench.py -- MinLZ block encode+decode throughput on B200 (one JSON line).

Workload (BASELINE.json configs[1]+[2], the configuration the metric is quoted
on): 4096 x 1 MiB synthetic JSON-like blocks per GPU.  One step = LevelFastest
encode of the whole batch (encode_l1 kernel + dense pack), then decode of the
packed token streams (decode kernel), all through the C ABI of
libminlz_cuda.so.  `value` = uncompressed bytes / (encode + decode time), data
resident in HBM; `e2e` = the same round trip through the host-pointer C ABI
calls with pinned host buffers (H2D and D2H inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--blocks B] [--block-size S] [--kind json]
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


def cpu_engine():
    """The CPU arm: the reference's own AMD64 assembly through oracle/_ref when that
    library is here (built in the dev container from /root/reference, travels with
    the snapshot), else the oracle port of the pure-Go path."""
    from oracle import refasm
    if refasm.build() is not None:
        return refasm, "reference", ("minio/minlz asm_amd64.s (encodeBlockAsm* / decodeBlockAsm) run natively via "
                                     "oracle/_ref, one block per task over a pthread pool")
    from oracle import binding
    binding.build()
    return binding, "port", "oracle port of the pure-Go path (oracle/_ref not built)"


def cpu_round_trip(host_blocks, nthreads, repeats=2, level=1):
    """Times the CPU arm (test/bench infrastructure) on host cores: encode +
    decode of `host_blocks` ([n, bs] uint8 numpy).  Returns GB/s of
    uncompressed bytes over the encode+decode time, and the parts."""
    import numpy as np
    oracle, _, _ = cpu_engine()
    n, bs = host_blocks.shape
    src = host_blocks.reshape(-1)
    soff = np.arange(n + 1, dtype=np.uint64) * bs
    cap = bs + 16
    doff = np.arange(n + 1, dtype=np.uint64) * cap
    enc = np.zeros(n * cap, dtype=np.uint8)  # pre-touched: no page faults in the timed region
    best_e = best_d = 1e30
    for _ in range(repeats):
        t0 = time.perf_counter()
        out_len = oracle.encode_batch_mt(level, src, soff, enc, doff, nthreads)
        t1 = time.perf_counter()
        best_e = min(best_e, t1 - t0)
    assert (out_len > 0).all()
    coff = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(out_len, out=coff[1:])
    comp = np.empty(int(coff[-1]), dtype=np.uint8)
    for i in range(n):
        comp[int(coff[i]):int(coff[i + 1])] = enc[int(doff[i]):int(doff[i]) + int(out_len[i])]
    dec = np.zeros(n * bs, dtype=np.uint8)
    for _ in range(repeats):
        t0 = time.perf_counter()
        status = oracle.decode_batch_mt(comp, coff, dec, soff, nthreads)
        t1 = time.perf_counter()
        best_d = min(best_d, t1 - t0)
    assert not status.any() and np.array_equal(dec, src)
    total = n * bs
    return {"value": total / (best_e + best_d) / 1e9, "encode_gbps": total / best_e / 1e9,
            "decode_gbps": total / best_d / 1e9, "seconds": best_e + best_d}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the
    box's host cores, all of them: its own AMD64 assembly via oracle/_ref (see
    cpu_engine), falling back to the oracle port only if that library is absent."""
    import numpy as np
    import torch
    import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    _, kind, what = cpu_engine()
    nsample = min(args.blocks, args.cpu_blocks)
    blocks = synth.make_blocks(args.kind, nsample, args.block_size, device="cpu").numpy()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_round_trip(blocks[: max(1, nsample // 8)], cores, repeats=1)
    t_tot = 0.0
    res = None
    for _ in range(args.steps):
        res = cpu_round_trip(blocks, cores, repeats=1, level=args.level)
        t_tot += res["seconds"]
    value = nsample * args.block_size * args.steps / t_tot / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t_tot / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d x %d B %s blocks per step, level %d encode + decode; %s"
                                   % (nsample, args.block_size, args.kind, args.level, what),
                         "encode_gbps": round(res["encode_gbps"], 4), "decode_gbps": round(res["decode_gbps"], 4)},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_funnel(args, mz, shard, dist, world, rank, dev, src, soff, enc, eoff, out_len, comp, coff, dec, status):
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

    for _ in range(max(args.warmup, 1)):
        back, stream = step()
    torch.cuda.synchronize()
    if rank == 0:
        assert torch.equal(back, full), "funnel round trip mismatch"
    assert int(status.abs().sum()) == 0
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        back, stream = step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0]) / args.steps
        cfg = workload_config(args)
        cfg["workload"] += "; FUNNEL: the batch of %d blocks lives on rank 0, raw blocks scattered and token streams / " \
                           "decoded blocks gathered over NCCL inside every step" % total
        print(json.dumps({"metric": METRIC, "value": round(total * bs / (ms * 1e-3) / 1e9, 4), "unit": UNIT, "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
                          "ratio": round(total * bs / int(stream.numel()), 4),
                          "nccl_bytes_per_step": int(2 * (world - 1) * nblk * bs + 2 * (int(stream.numel()) * (world - 1)) // world)}))


def workload_config(args):
    return {"workload": "configs[1]+[2]: %d x %d B synthetic %s blocks per GPU, %s encode then "
                        "batched decode of the packed token streams" %
                        (args.blocks, args.block_size, args.kind,
                         {-1: "LevelSuperFast (encode_l0)", 1: "LevelFastest (encode_l1)", 2: "LevelBalanced (encode_l2)"}[args.level]),
            "blocks_per_gpu": args.blocks, "block_size": args.block_size, "level": args.level,
            "encoder_flavor": getattr(args, "flavor", "amd64"),
            "cache": "inputs (%.1f GB per pass) larger than the 126 MB L2, no flush needed" %
                     (args.blocks * args.block_size / 1e9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--blocks", type=int, default=4096)
    ap.add_argument("--block-size", type=int, default=1 << 20)
    ap.add_argument("--kind", default="json")
    ap.add_argument("--level", type=int, default=1, choices=(-1, 1, 2),
                    help="1 = LevelFastest (headline), 2 = LevelBalanced, -1 = LevelSuperFast")
    ap.add_argument("--flavor", default="auto", choices=("auto", "go", "amd64"),
                    help="which reference build the encoder mirrors byte for byte: the amd64 assembly (what the "
                         "reference arm runs on this box; default for levels -1/1) or the pure-Go functions")
    ap.add_argument("--funnel", action="store_true",
                    help="N > 1 only: the whole batch lives on rank 0 and is scattered / gathered over NCCL every step "
                         "(SURVEY 8e 'including scatter/gather'); the default keeps the data pre-sharded")
    ap.add_argument("--cpu-blocks", type=int, default=1024, help="bounded sample for the CPU legs")
    ap.add_argument("--e2e-blocks", type=int, default=4096, help="blocks per e2e step (host buffers)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.flavor == "auto":
        args.flavor = "go" if args.level == 2 else "amd64"

    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import minlz_b200 as mz
    from minlz_b200 import shard
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
    nblk, bs = args.blocks, args.block_size
    mz.set_encoder_flavor(mz.FlavorAMD64 if args.flavor == "amd64" else mz.FlavorGo)

    # ---- synthetic input, resident in HBM; independent blocks shard by rank
    src = synth.make_blocks(args.kind, nblk, bs, device=dev, first=rank * nblk).reshape(-1)
    soff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * bs
    cap = bs + 16
    eoff = torch.arange(nblk + 1, dtype=torch.int64, device=dev) * cap
    enc = torch.empty(nblk * cap, dtype=torch.uint8, device=dev)
    out_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
    comp = torch.empty(nblk * bs, dtype=torch.uint8, device=dev)
    coff = torch.zeros(nblk + 1, dtype=torch.int64, device=dev)
    dec = torch.empty(nblk * bs, dtype=torch.uint8, device=dev)
    status = torch.zeros(nblk, dtype=torch.int32, device=dev)

    stream = torch.cuda.current_stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step(timed):
        if timed is not None:
            ev[0].record(stream)
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
        mz.decode_blocks_dev(comp, coff, dec, soff, status)
        if timed is not None:
            ev[3].record(stream)
            torch.cuda.synchronize()
            timed["enc"] += ev[0].elapsed_time(ev[1])
            timed["pack"] += ev[1].elapsed_time(ev[2])
            timed["dec"] += ev[2].elapsed_time(ev[3])

    if args.funnel and world > 1:
        run_funnel(args, mz, shard, dist, world, rank, dev, src, soff, enc, eoff, out_len, comp, coff, dec, status)
        dist.destroy_process_group()
        return

    for _ in range(args.warmup):
        step(None)
    torch.cuda.synchronize()
    # correctness gate before timing: the round trip must reproduce the input
    if args.warmup == 0:
        step(None)
        torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0, "decode reported corrupt blocks"
    assert int((out_len <= 0).sum()) == 0, "encoder returned incompressible on compressible data"
    assert torch.equal(dec, src), "round trip mismatch"
    comp_bytes = int(coff[-1])
    # encoder bytes against the checker (never timed here): the reference's real assembly when
    # oracle/_ref is on the box and the flavour is amd64, else the oracle restatement of the flavour
    parity = "round trip verified bit-exact before timing"
    if rank == 0 and not args.no_cpu:
        from oracle import binding as oracle_port
        ref_engine, ref_kind, _ = cpu_engine()
        picks = sorted({0, 1, nblk // 2, nblk - 1})
        h_len = out_len.cpu().numpy()
        for i in picks:
            got = enc[i * cap:i * cap + int(h_len[i])].cpu().numpy().tobytes()
            blk = src[i * bs:(i + 1) * bs].cpu().numpy()
            if args.flavor == "amd64" and ref_kind == "reference":
                want, who = ref_engine.encode_block(blk, args.level), "the reference's amd64 assembly (oracle/_ref)"
            else:
                want = oracle_port.encode_block(blk, args.level, flavor="asm" if args.flavor == "amd64" else "go")
                who = "the oracle restatement of the %s flavour" % args.flavor
            assert got == want, "encoder bytes differ from %s on block %d" % (who, i)
        parity = "encoder bytes identical to %s on blocks %s; %s" % (who, picks, parity)

    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    parts = {"enc": 0.0, "pack": 0.0, "dec": 0.0}
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        t_start.record(stream)
        for _ in range(args.steps):
            step(parts)
        t_end.record(stream)
        torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    total_ms = t_start.elapsed_time(t_end)
    tms = torch.tensor([total_ms, parts["enc"], parts["pack"], parts["dec"]], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, pack_ms, dec_ms = [float(x) for x in tms.tolist()]
    K = args.steps
    U = nblk * bs
    step_ms = total_ms / K
    value = world * U / (step_ms * 1e-3) / 1e9

    # ---- e2e: host-pointer C ABI calls with pinned buffers -----------------
    e2e = None
    if not args.no_e2e:
        nb = min(nblk, args.e2e_blocks)
        h_src = torch.empty(nb * bs, dtype=torch.uint8).pin_memory()
        h_src.copy_(src[: nb * bs])
        h_dec = torch.empty(nb * bs, dtype=torch.uint8).pin_memory()
        h_comp = torch.empty(nb * bs, dtype=torch.uint8).pin_memory()
        n_src, n_dec, n_comp = h_src.numpy(), h_dec.numpy(), h_comp.numpy()
        hs = np.arange(nb + 1, dtype=np.uint64) * bs
        hc = np.zeros(nb + 1, dtype=np.uint64)
        hst = np.zeros(nb, dtype=np.int32)

        e2e_parts = {"enc": 0.0, "dec": 0.0}

        def e2e_step():
            # host blocks -> packed token streams on the host -> host blocks again
            t0_ = time.perf_counter()
            cb_ = mz.encode_blocks_packed_into(n_src, hs, n_comp, hc, args.level, device=local)
            t1_ = time.perf_counter()
            mz.decode_blocks_into(n_comp, hc, n_dec, hs, hst, device=local)
            e2e_parts["enc"] += t1_ - t0_
            e2e_parts["dec"] += time.perf_counter() - t1_
            return cb_

        for _ in range(min(args.warmup, 2)):
            e2e_step()
        assert np.array_equal(n_dec, n_src) and not hst.any()
        if dist is not None:
            dist.barrier()
        e2e_parts["enc"] = e2e_parts["dec"] = 0.0
        t0 = time.perf_counter()
        ek = max(1, min(K, 3))
        for _ in range(ek):
            cb = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / ek
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        e2e = {"value": round(world * nb * bs / dt / 1e9, 4), "unit": UNIT,
               "h2d_bytes_per_step": int(nb * bs + cb + 2 * 8 * (nb + 1) * 2),
               "d2h_bytes_per_step": int(cb + nb * bs + 8 * nb),
               "blocks_per_step": nb, "ms_per_step": round(dt * 1e3, 3),
               "encode_call_ms": round(e2e_parts["enc"] / ek * 1e3, 3), "decode_call_ms": round(e2e_parts["dec"] / ek * 1e3, 3),
               "api": "mzcu_encode_blocks_packed + mzcu_decode_blocks (host pointers, pinned)"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    enc_bytes = U + comp_bytes          # algorithmic: read source once, write tokens once
    dec_bytes = comp_bytes + U          # read tokens once, write output once
    enc_gbs = enc_bytes / (enc_ms / K * 1e-3) / 1e9
    dec_gbs = dec_bytes / (dec_ms / K * 1e-3) / 1e9
    dominant_is_enc = enc_ms >= dec_ms
    enc_kernel = {-1: "encode_l1_kernel<true> (L0 params)", 1: "encode_l1_kernel<false>", 2: "encode_l2_kernel"}[args.level]
    if args.flavor == "amd64":
        enc_kernel = enc_kernel.replace("encode_l1_kernel", "encode_l1_asm_kernel")
    enc_traffic = ncu_traffic("encode") if args.level == 1 else None  # the ncu capture is of the L1 headline
    roof = {"bound": "hbm", "kernel": enc_kernel if dominant_is_enc else "decode_pc_kernel",
            "achieved": round(enc_gbs if dominant_is_enc else dec_gbs, 3), "peak": peak, "unit": "GB/s",
            "frac": round((enc_gbs if dominant_is_enc else dec_gbs) / peak, 5),
            "traffic": enc_traffic if dominant_is_enc else ncu_traffic("decode"), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": enc_bytes if dominant_is_enc else dec_bytes,
            "share_of_step": round((enc_ms if dominant_is_enc else dec_ms) / total_ms, 4)}
    if roof["traffic"]:
        # what the kernel actually moves: table probes are random 32-byte reads and the B200 L2 fetches a
        # whole 128-byte line per miss (profiles/r01_micro_gather32.txt: 4.7 TB/s ceiling for such gathers)
        t_ms = (enc_ms if dominant_is_enc else dec_ms) / K
        roof["traffic_gbs"] = round(roof["traffic"] / (t_ms * 1e-3) / 1e9, 1)
        roof["traffic_frac_of_peak"] = round(roof["traffic_gbs"] / peak, 4)
        roof["random_line_gather_ceiling_gbs"] = 4700.0
    roof_dec = {"bound": "hbm", "kernel": "decode_pc_kernel", "achieved": round(dec_gbs, 3), "peak": peak, "unit": "GB/s",
                "frac": round(dec_gbs / peak, 5), "traffic": ncu_traffic("decode"),
                "algorithmic_bytes_per_launch": dec_bytes, "share_of_step": round(dec_ms / total_ms, 4)}
    roof_enc = {"bound": "hbm", "kernel": enc_kernel, "achieved": round(enc_gbs, 3), "peak": peak, "unit": "GB/s",
                "frac": round(enc_gbs / peak, 5), "traffic": enc_traffic,
                "algorithmic_bytes_per_launch": enc_bytes, "share_of_step": round(enc_ms / total_ms, 4)}

    cpu = None
    if not args.no_cpu and world >= 1:
        cores = os.cpu_count() or 1
        nsample = min(nblk, args.cpu_blocks)
        hb = src[: nsample * bs].cpu().numpy().reshape(nsample, bs)
        r = cpu_round_trip(hb, cores, repeats=2, level=args.level)
        _, ckind, cwhat = cpu_engine()
        cpu = {"value": round(r["value"], 4), "unit": UNIT, "cores": cores, "kind": ckind,
               "sample": "first %d of the %d blocks, level %d encode + decode, best of 2; %s"
                         % (nsample, nblk, args.level, cwhat),
               "encode_gbps": round(r["encode_gbps"], 4), "decode_gbps": round(r["decode_gbps"], 4)}

    line = {
        "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(args),
        "encode_gbps": round(world * U / (enc_ms / K * 1e-3) / 1e9, 3),
        "decode_gbps": round(world * U / (dec_ms / K * 1e-3) / 1e9, 3),
        "ms": {"encode": round(enc_ms / K, 4), "pack": round(pack_ms / K, 4), "decode": round(dec_ms / K, 4)},
        "ratio": round(U / comp_bytes, 4),
        "roofline": roof, "roofline_decode": roof_dec, "roofline_encode": roof_enc,
        "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks.summary(),
        "gpu_launches": K * 4,  # per step: encode_l1, scan_lengths, pack_blocks, decode
        "parity": parity,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
