"""Bare host<->device link ceiling: plain pinned cudaMemcpyAsync, nothing of the codec involved.

    python scripts/link_probe.py [--gpus 1,2,4,8] [--seconds 1.5] [--mb 1024] [--modes hostalloc,registered_thp]

For every N in --gpus, N worker PROCESSES (one per GPU, like bench.py's ranks) copy concurrently:
H2D alone, D2H alone, and both at once (two streams), each for --seconds; the per-GPU and aggregate
GB/s go to stdout as one JSON line per (N, mode).  Modes are how the host buffer is obtained:
  hostalloc       torch pin_memory (cudaHostAlloc)
  registered      malloc'd (numpy) memory + cudaHostRegister
  registered_thp  the same, 2 MiB-aligned and madvise(MADV_HUGEPAGE) before first touch
This is the denominator of bench.py's e2e.frac_of_link; bench.py carries its own copy of the timed
loop (link_ceiling) so the figure in a bench line comes from the same run and the same processes.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import mmap
import os
import sys
import time

import torch
import torch.multiprocessing as mp


def host_buffer(nbytes: int, mode: str):
    """A pinned uint8 tensor of nbytes (and whatever must stay alive with it)."""
    if mode == "hostalloc":
        return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True), None
    m = mmap.mmap(-1, nbytes + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
    off = (-addr) % (2 << 20)
    if mode == "registered_thp":
        libc = ctypes.CDLL(None, use_errno=True)
        libc.madvise(ctypes.c_void_p(addr + off), ctypes.c_size_t(nbytes), 14)  # MADV_HUGEPAGE
    t = torch.frombuffer(m, dtype=torch.uint8, count=nbytes, offset=off)
    t.fill_(1)  # first touch
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), nbytes, 0)
    if int(rc) != 0:
        raise RuntimeError(f"cudaHostRegister failed: {rc}")
    return t, m


def copy_loop(dev, host_in, host_out, d_in, d_out, seconds, h2d, d2h, chunk):
    """Issues chunked async copies on one stream per direction until `seconds` have passed; returns GB/s per direction."""
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    n = host_in.numel()
    torch.cuda.synchronize(dev)
    moved_in = moved_out = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            if h2d:
                with torch.cuda.stream(s_in):
                    d_in[a:b].copy_(host_in[a:b], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out):
                    host_out[a:b].copy_(d_out[a:b], non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()
        moved_in += n if h2d else 0
        moved_out += n if d2h else 0
    dt = time.perf_counter() - t0
    return moved_in / dt / 1e9, moved_out / dt / 1e9


def worker(rank, world, args, mode, barrier, results):
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    nbytes = args.mb << 20
    host_in, keep1 = host_buffer(nbytes, mode)
    host_out, keep2 = host_buffer(nbytes, mode)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    chunk = args.chunk_mb << 20
    copy_loop(dev, host_in, host_out, d_in, d_out, 0.2, True, True, chunk)  # warm
    out = {}
    for name, h2d, d2h in (("h2d_alone", True, False), ("d2h_alone", False, True), ("duplex", True, True)):
        barrier.wait()
        gi, go = copy_loop(dev, host_in, host_out, d_in, d_out, args.seconds, h2d, d2h, chunk)
        out[name] = {"h2d_gbs": gi, "d2h_gbs": go}
    results[rank] = out
    barrier.wait()
    if mode != "hostalloc":
        torch.cuda.cudart().cudaHostUnregister(host_in.data_ptr())
        torch.cuda.cudart().cudaHostUnregister(host_out.data_ptr())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--mb", type=int, default=1024, help="bytes per direction per GPU, MiB")
    ap.add_argument("--chunk-mb", type=int, default=256)
    ap.add_argument("--modes", default="hostalloc")
    args = ap.parse_args()
    avail = torch.cuda.device_count()
    mp.set_start_method("spawn", force=True)
    for mode in args.modes.split(","):
        for n in [int(x) for x in args.gpus.split(",")]:
            if n > avail:
                continue
            mgr = mp.Manager()
            results = mgr.dict()
            barrier = mgr.Barrier(n)
            procs = [mp.Process(target=worker, args=(r, n, args, mode, barrier, results)) for r in range(n)]
            for p in procs:
                p.start()
            for p in procs:
                p.join()
            if len(results) != n:
                print(json.dumps({"gpus": n, "mode": mode, "error": "a worker died"}), flush=True)
                continue
            line = {"gpus": n, "mode": mode, "mb_per_direction": args.mb, "chunk_mb": args.chunk_mb, "seconds": args.seconds}
            for phase in ("h2d_alone", "d2h_alone", "duplex"):
                hi = [results[r][phase]["h2d_gbs"] for r in range(n)]
                ho = [results[r][phase]["d2h_gbs"] for r in range(n)]
                line[phase] = {"h2d_gbs_sum": sum(hi), "d2h_gbs_sum": sum(ho), "h2d_gbs_per_gpu": [round(x, 2) for x in hi],
                               "d2h_gbs_per_gpu": [round(x, 2) for x in ho]}
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    sys.exit(main())
