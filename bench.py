#!/usr/bin/env python
"""Benchmark of the .spz per-gaussian codec hot path (BASELINE.json: encode/decode Mgaussians/s and
HBM GB/s vs roofline at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--sh-degree D]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own CPU pack/unpack on the host cores

Workload (config 5 of BASELINE.json; it fits one GPU, so it is also the N=1 workload): a synthetic
100M-gaussian SH-degree-3 cloud, v3 stream, sharded by contiguous point range over the N ranks
(strong scaling: the cloud is fixed, each rank owns n/N points; no collective on the data path).
One "step" = encode the rank's shard (float planes -> byte planes), then decode the result back
(byte planes -> float planes, with a coordinate flip folded in).  `value` = gaussians through that
encode+decode round trip per second, whole job, inputs resident in HBM, timed with CUDA events on
the launching stream, max over ranks.  `e2e` = the same step through the host-pointer C-ABI
(spzb200_encode_host / spzb200_decode_host: pinned host planes in, pinned host planes out, H2D and
D2H inside the timed region), issued full duplex -- step i's encode and the decode of step i-1's
stream run concurrently from two host threads so both directions of the PCIe link are busy; the
one-thread "encode, then decode its result" figure is reported beside it as e2e.sequential.  The working set (30.1 GB per step at N=1) is far larger than the
126 MB L2, so no explicit flush is needed between iterations.

The oracle / reference build under oracle/ is used here ONLY as the timed CPU baseline
(`cpu_baseline`, `--impl reference`), never on the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "encode+decode Mgaussians/s (v3, SH3)"
UNIT = "Mgaussians/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=100_000_000, help="gaussians in the whole cloud")
    ap.add_argument("--sh-degree", type=int, default=3)
    ap.add_argument("--from-coord", type=int, default=6, help="PackOptions.from (6 = RDF)")
    ap.add_argument("--to-coord", type=int, default=6, help="UnpackOptions.to (6 = RDF)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ply", action="store_true", help="skip the fused PLY-rows kernel timings (N=1 only)")
    ap.add_argument("--cpu-sample-points", type=int, default=0, help="0 = auto")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from an ncu --set full capture")
    return ap.parse_args()


def workload_name(n, deg):
    return f"synthetic {n / 1e6:g}M gaussians SH degree {deg} v3 encode+decode, sharded by point range"


# -------------------------------------------------------------------------------------------------
# clocks: NVML sampled from a thread during the timed region
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# -------------------------------------------------------------------------------------------------
# the reference's CPU implementation, timed (cpu_baseline and --impl reference)
# -------------------------------------------------------------------------------------------------
def cpu_reference_run(points_total, deg, frm, to, steps, warmup, threads, sample_points):
    """Times encode+decode of a bounded sample of the workload with the reference's own code
    (oracle/_ref/libspz_ref.so = the unmodified C++ compiled in place) or, if that build is not
    present, the C restatement (oracle/_build/libspz_oracle.so).  `threads` independent point
    ranges run concurrently (the reference itself is single-threaded; points are independent)."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np

    import oracle as O
    from spz_b200.synth import numpy_cloud

    kind = "reference"
    try:
        impl_factory = O.Ref
        impl_factory()
    except Exception:  # noqa: BLE001
        kind = "port"
        impl_factory = O.Oracle
    per_thread = max(1, sample_points // threads)
    clouds = []
    for t in range(threads):
        c = numpy_cloud(per_thread, deg, seed=1000 + t)
        clouds.append(O.Cloud(per_thread, deg, *c.planes()))
    impls = [impl_factory() for _ in range(threads)]

    def one(i):
        p = impls[i].pack(clouds[i], frm)
        t_pack = impls[i].last_seconds
        g = impls[i].unpack(p, to)
        t_unpack = impls[i].last_seconds
        return t_pack, t_unpack, int(g.n)

    times, inner = [], []
    with ThreadPoolExecutor(threads) as ex:
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            res = list(ex.map(one, range(threads)))
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                inner.append(res)
    n_step = per_thread * threads
    mean = sum(times) / len(times)
    out = {"value": n_step / mean / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
           "sample": f"{n_step} gaussians per step ({threads} x {per_thread}-point ranges of the synthetic SH{deg} "
                     f"workload), {steps} steps; wall time of pack+unpack calls incl. marshalling copies",
           "ms_per_step": mean * 1e3}
    if kind == "reference" and inner and inner[0][0][0] is not None:
        # time inside packGaussians / unpackGaussians alone (marshalling excluded), slowest thread
        pk = statistics.mean(max(r[0] for r in step) for step in inner)
        up = statistics.mean(max(r[1] for r in step) for step in inner)
        out["pack_mgs"] = n_step / pk / 1e6
        out["unpack_mgs"] = n_step / up / 1e6
        out["value"] = n_step / (pk + up) / 1e6
        out["ms_per_step"] = (pk + up) * 1e3
        out["sample"] += "; value = time inside packGaussians+unpackGaussians only"
    return out


def traffic_for(kernel, gaussians, override):
    """DRAM bytes per launch of `kernel`: the ncu --set full capture recorded in profiles/traffic.json
    (taken at a different launch size), scaled per gaussian to this launch."""
    if override is not None:
        return override
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
        return t["dram_bytes"] / t["gaussians"] * gaussians
    except Exception:  # noqa: BLE001
        return None


def ply_rows_kernels(ctx, codec, dev, deg, n=40_000_000, steps=8):
    """the fused kernels on the other side of the codec, .ply vertex records <-> packed
    planes (canonical property order), device-resident, CUDA events, median of `steps`."""
    import torch
    names = codec.ply_property_names(deg)
    w = len(names)
    rows = torch.empty(n * w, dtype=torch.float32, device=dev).uniform_(-1, 1)
    out = codec.alloc_packed(n, deg, 3, device=dev)
    back = torch.empty_like(rows)
    per = 4 * w + codec.packed_bytes_per_gaussian(deg)
    res = {"points": n, "sh_degree": deg, "bytes_per_gaussian": per,
           "note": "packGaussians(loadSplatFromPly(...)) / saveSplatToPly(unpackGaussians(...)) row layout in one kernel each; "
                   "not part of `value`"}
    for key, fn in (("rows_to_packed", lambda: ctx.encode_ply_device(rows, n, names, deg, 6, out=out)),
                    ("packed_to_rows", lambda: ctx.decode_ply_device(out, names, 6, out=back))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            e[0].record()
            fn()
            e[1].record()
            torch.cuda.synchronize()
            ts.append(e[0].elapsed_time(e[1]))
        ms = statistics.median(ts)
        res[key] = {"ms": ms, "hbm_gbs": per * n / ms / 1e6, "mgaussians_s": n / ms / 1e3}
    return res


def small_cloud_latency(ctx, codec, torch_cloud, dev, n, deg, args):
    import torch
    c = torch_cloud(n, deg, dev, seed=60)
    p = codec.alloc_packed(n, deg, 3, device=dev)
    g = codec.alloc_cloud(n, deg, device=dev)
    dev_us, host_us = [], []
    for i in range(60):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        ctx.encode_device(c, args.from_coord, out=p)
        ctx.decode_device(p, args.to_coord, out=g)
        e[1].record()
        torch.cuda.synchronize()
        if i >= 10:
            dev_us.append(e[0].elapsed_time(e[1]) * 1e3)
    hc = codec.alloc_cloud(n, deg, pinned=True, numpy_arrays=True)
    hp = codec.alloc_packed(n, deg, 3, pinned=True, numpy_arrays=True)
    hb = codec.alloc_cloud(n, deg, pinned=True, numpy_arrays=True)
    for src, dst in zip(c.planes(), hc.planes()):
        if dst.size:
            torch.from_numpy(dst).copy_(src)
    torch.cuda.synchronize()
    for i in range(60):
        t0 = time.perf_counter()
        ctx.encode_host(hc, args.from_coord, out=hp)
        ctx.decode_host(hp, args.to_coord, out=hb)
        if i >= 10:
            host_us.append((time.perf_counter() - t0) * 1e6)
    return {"points": n, "sh_degree": deg, "device_encode_plus_decode_us": statistics.median(dev_us),
            "host_api_encode_plus_decode_us": statistics.median(host_us),
            "note": "median of 50; device = four launches (two tile kernels + two scalar tails) on resident planes, issued from Python through ctypes, so it is launch-bound, not bandwidth-bound (the data moves in ~6 us); host = spzb200_encode_host + spzb200_decode_host with pinned planes, i.e. 18 MB over PCIe each way"}


def time_host_zlib(packed, deg, sample_points):
    """gzip stays on the host (north star) and is reported apart from the codec: deflate / inflate of
    the container bytes of the first `sample_points` encoded gaussians with the reference's zlib
    parameters (default level, gzip wrapper, memLevel 9; load-spz.cc:186-214), one thread."""
    import struct
    import zlib

    from spz_b200 import codec
    widths = codec.byte_plane_widths(deg, 3)
    planes = [p[:w * sample_points].cpu().numpy().tobytes() for p, w in zip(packed.planes(), widths)]
    order = [planes[0], planes[3], planes[4], planes[1], planes[2], planes[5]]  # stream order
    stream = struct.pack("<IIIBBBB", 0x5053474e, 3, sample_points, deg, 12, 0, 0) + b"".join(order)
    threads = host_threads()
    t0 = time.perf_counter()
    gz = codec.gzip_bytes(stream, 1)          # the reference's single-thread zlib stream, byte for byte
    t1 = time.perf_counter()
    back = codec.gunzip_bytes(gz, 1)
    t2 = time.perf_counter()
    gzp = codec.gzip_bytes(stream, threads)   # block-parallel zlib, one standard gzip member
    t3 = time.perf_counter()
    backp = codec.gunzip_bytes(gzp, threads)
    t4 = time.perf_counter()
    assert back == stream and backp == stream and zlib.decompress(gzp, 16 + zlib.MAX_WBITS) == stream
    mg = lambda dt: sample_points / dt / 1e6  # noqa: E731
    return {"sample": f"container of the first {sample_points} encoded gaussians ({len(stream)} bytes), zlib {zlib.ZLIB_RUNTIME_VERSION}",
            "reference_1_thread": {"deflate_mb_s": len(stream) / (t1 - t0) / 1e6, "inflate_mb_s": len(stream) / (t2 - t1) / 1e6,
                                   "deflate_mgaussians_s": mg(t1 - t0), "inflate_mgaussians_s": mg(t2 - t1), "ratio": len(gz) / len(stream)},
            "block_parallel": {"threads": threads, "deflate_mb_s": len(stream) / (t3 - t2) / 1e6, "inflate_mb_s": len(stream) / (t4 - t3) / 1e6,
                               "deflate_mgaussians_s": mg(t3 - t2), "inflate_mgaussians_s": mg(t4 - t3), "ratio": len(gzp) / len(stream)}}


def host_memory_available():
    """min(MemAvailable, cgroup limit - usage) in bytes, or None."""
    vals = []
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                vals.append(int(ln.split()[1]) * 1024)
    except OSError:
        pass
    try:
        mx = open("/sys/fs/cgroup/memory.max").read().strip()
        if mx != "max":
            vals.append(int(mx) - int(open("/sys/fs/cgroup/memory.current").read()))
    except (OSError, ValueError):
        pass
    return min(vals) if vals else None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = max(1, min(host_threads(), 32))
    sample = args.cpu_sample_points or threads * 250_000
    r = cpu_reference_run(args.points, args.sh_degree, args.from_coord, args.to_coord, args.steps, args.warmup, threads, sample)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32->u8", "data": "synthetic",
            "config": {"workload": workload_name(args.points, args.sh_degree), "from": args.from_coord, "to": args.to_coord},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    for k in ("pack_mgs", "unpack_mgs"):
        if k in r:
            line[k] = r[k]
    print(json.dumps(line), flush=True)
    return 0


# -------------------------------------------------------------------------------------------------
# multi-rank plumbing (no data-path collective: only the barrier and these scalar reductions)
# -------------------------------------------------------------------------------------------------
def reduce_scalar(dist, x, op, device):
    """max / sum of a python scalar over all ranks; identity when not distributed."""
    if dist is None:
        return x
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return t.item()


def rank_shard(points_total, sh_degree, world, rank):
    """[a, b) of this rank: contiguous point range with tile-aligned boundaries (spzb200_shard_range)."""
    from spz_b200 import codec
    return codec.shard_range(points_total, sh_degree, world, rank)


# -------------------------------------------------------------------------------------------------
# the B200 arm
# -------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch

    from spz_b200 import codec
    from spz_b200.synth import torch_cloud

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the codec has no CPU path (use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    deg, n_total = args.sh_degree, args.points
    a, b = rank_shard(n_total, deg, world, rank)
    n = b - a
    ctx = codec.Context(local)
    dev = torch.device("cuda", local)
    cloud = torch_cloud(n, deg, dev, seed=1 + rank)
    packed = codec.alloc_packed(n, deg, 3, device=dev)
    decoded = codec.alloc_cloud(n, deg, device=dev)
    alg_bytes = codec.algorithmic_bytes_per_gaussian(deg, 3)

    def step(evs=None):
        if evs:
            evs[0].record()
        ctx.encode_device(cloud, args.from_coord, out=packed)
        if evs:
            evs[1].record()
        ctx.decode_device(packed, args.to_coord, out=decoded)
        if evs:
            evs[2].record()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = ctx.info()["kernel_launches"]
    sampler = ClockSampler(local)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    e0.record()
    for s in range(args.steps):
        step(evs[s])
    e1.record()
    barrier()
    clocks = sampler.stop()
    total_ms = e0.elapsed_time(e1)
    launches = ctx.info()["kernel_launches"] - launches0
    enc_ms = statistics.mean(ev[0].elapsed_time(ev[1]) for ev in evs)
    dec_ms = statistics.mean(ev[1].elapsed_time(ev[2]) for ev in evs)

    # ---- e2e: the same step through the host-pointer C-ABI with pinned host planes ---------------
    e2e = None
    need = 2 * codec.float_bytes_per_gaussian(deg) * n + 2 * codec.packed_bytes_per_gaussian(deg) * n
    avail = host_memory_available()
    if not args.no_e2e and avail is not None and need * world > 0.6 * avail:
        raise SystemExit(f"bench.py: e2e needs {need * world / 1e9:.1f} GB of pinned host memory, only {avail / 1e9:.1f} GB "
                         "available; rerun with --no-e2e or fewer --points")
    if not args.no_e2e:
        del decoded
        torch.cuda.empty_cache()
        h_cloud = codec.alloc_cloud(n, deg, pinned=True, numpy_arrays=True)
        h_packed = codec.alloc_packed(n, deg, 3, pinned=True, numpy_arrays=True)
        h_back = codec.alloc_cloud(n, deg, pinned=True, numpy_arrays=True)
        for src, dst in zip(cloud.planes(), h_cloud.planes()):
            if dst.size:
                torch.from_numpy(dst).copy_(src)
        torch.cuda.synchronize()
        phases = []

        def e2e_step():
            _, t_enc = ctx.encode_host(h_cloud, args.from_coord, out=h_packed)
            _, t_dec = ctx.decode_host(h_packed, args.to_coord, out=h_back)
            return t_enc, t_dec

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            phases.append(e2e_step())
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        h2d = phases[-1][0]["h2d_bytes"] + phases[-1][1]["h2d_bytes"]
        d2h = phases[-1][0]["d2h_bytes"] + phases[-1][1]["d2h_bytes"]
        e2e = {"s": e2e_s, "h2d": h2d, "d2h": d2h,
               "enc": {k: statistics.mean(p[0][k] for p in phases) for k in ("h2d_ms", "kernel_ms", "d2h_ms", "wall_ms")},
               "dec": {k: statistics.mean(p[1][k] for p in phases) for k in ("h2d_ms", "kernel_ms", "d2h_ms", "wall_ms")}}
        # the e2e path is the same kernels: spot-check its bytes against the device-resident result
        for name, hp, dp in zip(codec.PLANES, h_packed.planes(), packed.planes()):
            k = min(hp.size, 1 << 22)
            if k and not torch.equal(torch.from_numpy(hp[:k]), dp[:k].cpu()):
                raise SystemExit(f"bench.py: e2e plane {name} differs from the device-resident encode")

        # Full-duplex form of the same step: a second host thread (own context) decodes the stream
        # the previous step produced while this step's cloud is being encoded, so both directions of
        # the PCIe link carry a step's worth of planes at once.  Same calls, same bytes per step.
        from concurrent.futures import ThreadPoolExecutor
        ctx2 = codec.Context(local)
        h_packed2 = codec.alloc_packed(n, deg, 3, pinned=True, numpy_arrays=True)
        streams = [h_packed, h_packed2]
        ctx.encode_host(h_cloud, args.from_coord, out=streams[1])  # step 0 decodes this one
        pool = ThreadPoolExecutor(2)

        def duplex_step(i):
            fa = pool.submit(ctx.encode_host, h_cloud, args.from_coord, streams[i % 2])
            fb = pool.submit(ctx2.decode_host, streams[(i + 1) % 2], args.to_coord, h_back)
            return fa.result()[1], fb.result()[1]

        duplex_step(0)
        barrier()
        t0 = time.perf_counter()
        dphases = [duplex_step(1 + i) for i in range(args.e2e_steps)]
        barrier()
        e2e["duplex_s"] = (time.perf_counter() - t0) / args.e2e_steps
        e2e["duplex_enc_wall_ms"] = statistics.mean(p[0]["wall_ms"] for p in dphases)
        e2e["duplex_dec_wall_ms"] = statistics.mean(p[1]["wall_ms"] for p in dphases)
        pool.shutdown()
        ctx2.close()
        hb, db = torch.from_numpy(h_back.sh[:1 << 20] if deg else h_back.alphas[:1 << 20]), None
        ref_back = ctx.decode_device(packed, args.to_coord)
        db = (ref_back.sh if deg else ref_back.alphas)[:hb.numel()].cpu()
        if not torch.equal(hb.view(torch.int32), db.view(torch.int32)):
            raise SystemExit("bench.py: duplex e2e decode differs from the device-resident decode")
        del ref_back

    # ---- reduce over ranks (max time; sums of bytes and launches) -----------------------------------
    def allmax(x):
        return reduce_scalar(dist, x, "max", dev)

    def allsum(x):
        return reduce_scalar(dist, x, "sum", dev)

    total_ms = allmax(total_ms)
    enc_ms_max, dec_ms_max = allmax(enc_ms), allmax(dec_ms)
    launches = int(allsum(launches))
    if e2e:
        e2e["s"] = allmax(e2e["s"])
        e2e["duplex_s"] = allmax(e2e["duplex_s"])
        e2e["h2d"] = int(allsum(e2e["h2d"]))
        e2e["d2h"] = int(allsum(e2e["d2h"]))

    # ---- config 1 of BASELINE.json is a 60k-gaussian file: far too small to be bandwidth-bound, so
    # what matters there is latency.  Device-resident launch pair and the host-pointer call, median of 50.
    latency = None
    if rank == 0:
        latency = small_cloud_latency(ctx, codec, torch_cloud, dev, 60_000, deg, args)

    ply_rows = None
    if rank == 0 and world == 1 and not args.no_ply:
        ply_rows = ply_rows_kernels(ctx, codec, dev, deg, n=min(n, 40_000_000))

    host_zlib = None
    if rank == 0 and not args.no_cpu_baseline:
        host_zlib = time_host_zlib(packed, deg, min(n, 400_000))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample_points or 2_000_000
        cpu = cpu_reference_run(n_total, deg, args.from_coord, args.to_coord, steps=2, warmup=1, threads=1, sample_points=sample)
        cpu = {k: cpu[k] for k in cpu if k != "ms_per_step"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        ms_per_step = total_ms / args.steps
        value = n_total / (ms_per_step * 1e-3) / 1e6
        # per-kernel roofline on this rank's shard (rank 0's durations; all shards are equal-sized)
        enc_gbs = alg_bytes * n / (enc_ms * 1e-3) / 1e9
        dec_gbs = alg_bytes * n / (dec_ms * 1e-3) / 1e9
        dec_env = os.environ.get("SPZB200_DECODE")
        dec_kernel = ("decodeTilesKernel" if deg == 0 or dec_env == "direct" else
                      "decodeTilesBulkKernel" if dec_env == "bulk" else "decodePerGaussianKernel")
        enc_env = os.environ.get("SPZB200_ENCODE")
        enc_kernel = ("encodePerGaussianKernel" if enc_env == "bulk" or (enc_env != "tiles" and deg == 3 and n <= 24_000_000)
                      else "encodeTilesKernel")
        dom = (enc_kernel, enc_gbs, enc_ms) if enc_ms >= dec_ms else (dec_kernel, dec_gbs, dec_ms)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32->u8", "data": "synthetic",
            "config": {"workload": workload_name(n_total, deg), "points_total": n_total, "points_per_gpu": n,
                       "sh_degree": deg, "stream_version": 3, "from": args.from_coord, "to": args.to_coord,
                       "l2": "working set per step >> 126 MB L2, no flush needed", "sharding": f"point-range x{world}, no collective"},
            "encode_mgs": n_total / (enc_ms_max * 1e-3) / 1e6, "decode_mgs": n_total / (dec_ms_max * 1e-3) / 1e6,
            "encode_ms": enc_ms_max, "decode_ms": dec_ms_max,
            "hbm_gbs_per_gpu": {"encode": enc_gbs, "decode": dec_gbs},
            "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": dom[1], "peak": peak, "unit": "GB/s",
                         "frac": dom[1] / peak, "frac_of_nominal_8TBs": dom[1] / 8000.0, "peak_source": peak_src,
                         "algorithmic_bytes_per_gaussian": alg_bytes, "gaussians_per_launch": n,
                         "avg_launch_ms": dom[2], "traffic": traffic_for(dom[0], n, args.traffic_bytes),
                         "traffic_source": "profiles/traffic.json (ncu --set full dram bytes per gaussian x gaussians per launch)",
                         "encode": {"achieved": enc_gbs, "frac": enc_gbs / peak, "avg_launch_ms": enc_ms},
                         "decode": {"achieved": dec_gbs, "frac": dec_gbs / peak, "avg_launch_ms": dec_ms}},
            "gpu_launches": launches, "clocks": clocks,
        }
        if e2e:
            line["e2e"] = {"value": n_total / e2e["duplex_s"] / 1e6, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                           "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": e2e["duplex_s"] * 1e3, "steps": args.e2e_steps,
                           "api": "spzb200_encode_host + spzb200_decode_host (pinned host planes)",
                           "mode": "full duplex: step i's encode_host and the decode_host of step i-1's stream are issued concurrently "
                                   "from two host threads (one context each), so H2D and D2H of a step's planes overlap on the PCIe link",
                           "encode_call_ms": e2e["duplex_enc_wall_ms"], "decode_call_ms": e2e["duplex_dec_wall_ms"],
                           "sequential": {"value": n_total / e2e["s"] / 1e6, "ms_per_step": e2e["s"] * 1e3,
                                          "mode": "one host thread: encode_host, then decode_host of its result",
                                          "phases_note": "h2d/kernel/d2h_ms are sums over point ranges of per-range stream time; ranges run on 3 streams and overlap, wall_ms is the call",
                                          "encode_phases_ms": e2e["enc"], "decode_phases_ms": e2e["dec"]}}
        if cpu:
            line["cpu_baseline"] = cpu
        if host_zlib:
            line["host_zlib"] = host_zlib
        if latency:
            line["latency_60k"] = latency
        if ply_rows:
            line["ply_rows"] = ply_rows
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
