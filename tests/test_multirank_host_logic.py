"""The N>1 path of bench.py on CPU: two gloo ranks agree on a partition of the cloud, reduce their
timings as max and their counters as sum, and the reference arm runs on rank 0 only.  (The data
path itself has no collective; these are the only cross-rank operations it uses.)"""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, os.environ["REPO"])
import torch, torch.distributed as dist
import bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, deg = 100_000_000, 3
a, b = bench.rank_shard(n, deg, world, rank)
edges = [None] * world
dist.all_gather_object(edges, (a, b))
mx = bench.reduce_scalar(dist, 10.0 + rank, "max", "cpu")
sm = bench.reduce_scalar(dist, 40 + rank, "sum", "cpu")
# every rank generates ITS slice of the one counter-seeded cloud; the shard hashes add up to the hash of the whole
from spz_b200 import codec
from spz_b200.synth import counter_cloud_torch
m, d2 = 7 * 1280 + 77, 2
a2, b2 = bench.rank_shard(m, d2, world, rank)
part = counter_cloud_torch(m, d2, "cpu", a2, b2, seed=1)
fwb = [4 * w for w in codec.float_plane_widths(d2)]
h = bench.allsum_u64(dist, bench.planes_hash(part.planes(), fwb, a2, "cpu"), "cpu")
whole = counter_cloud_torch(m, d2, "cpu", seed=1)
h_whole = bench.planes_hash(whole.planes(), fwb, 0, "cpu")
swapped = [p.clone() for p in whole.planes()]
swapped[5][8:16], swapped[5][16:24] = whole.planes()[5][16:24].clone(), whole.planes()[5][8:16].clone()
h_swapped = bench.planes_hash(swapped, fwb, 0, "cpu")
dist.barrier()
if rank == 0:
    print(json.dumps({"edges": edges, "max": mx, "sum": sm, "hash": h, "hash_whole": h_whole, "hash_swapped": h_swapped}))
dist.destroy_process_group()
"""


def test_two_gloo_ranks_partition_and_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=ROOT, SPZB200_NO_REBUILD="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    (a0, b0), (a1, b1) = out["edges"]
    assert a0 == 0 and b0 == a1 and b1 == 100_000_000
    assert b0 % 1280 == 0 and abs((b0 - a0) - (b1 - a1)) <= 2 * 1280
    assert out["max"] == 11.0 and out["sum"] == 81.0
    # shard hashes are additive over the partition, and sensitive to two words trading places
    assert out["hash"] == out["hash_whole"] != out["hash_swapped"]


def test_python_shard_ranges_equal_the_library():
    import bench
    from spz_b200 import codec
    for deg in range(4):
        for n in (0, 1, 6399, 6400, 12_499_200, 100_000_000, 3_000_000_007):
            for shards in (1, 2, 3, 4, 8):
                for i in range(shards):
                    assert bench.shard_range_py(n, deg, shards, i) == codec.shard_range(n, deg, shards, i), (deg, n, shards, i)


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                        "--cpu-sample-points", "20000"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0
    # the same `config` object as the B200 arm prints for the same arguments
    import argparse
    import bench
    want = bench.bench_config(argparse.Namespace(points=100_000_000, sh_degree=3, from_coord=6, to_coord=6), 2)
    assert line["config"] == want and want["points_per_gpu"] == 49_999_360
