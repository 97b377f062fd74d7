"""Seeded synthetic gaussian clouds.

positions U(-10,10); scales U(-7,1); rotations 4 x U(-1,1) unnormalised (near-zero norms replaced
by the identity); alphas U(-6,6); colors U(-2,2); sh U(-0.5,0.5) with ~1% of values set to exact
rounding ties (2j+1)/256 and ~0.1% pushed outside [-1,1] to exercise the clamps.

Two generators with the same distribution: numpy on the host (tests, CPU baseline samples) and
torch on the device (full-size benchmark clouds; written plane by plane in bounded slices so the
generator itself never needs more than ~1 GB of scratch).
"""
from __future__ import annotations

import numpy as np

from .codec import CloudPlanes, SH_DIM, float_plane_widths


def numpy_cloud(n: int, sh_degree: int, seed: int = 1) -> CloudPlanes:
    rng = np.random.default_rng(seed)
    u = lambda k, lo, hi: rng.uniform(lo, hi, k).astype(np.float32)  # noqa: E731
    pos, scl = u(3 * n, -10, 10), u(3 * n, -7, 1)
    rot = u(4 * n, -1, 1)
    if n:
        r = rot.reshape(n, 4)
        bad = np.linalg.norm(r, axis=1) < 1e-3
        r[bad] = np.array([0, 0, 0, 1], np.float32)
    alp, col = u(n, -6, 6), u(3 * n, -2, 2)
    m = 3 * SH_DIM[sh_degree] * n
    sh = u(m, -0.5, 0.5)
    if m:
        pick = rng.random(m)
        ties = pick < 0.01
        sh[ties] = ((2 * rng.integers(-128, 128, int(ties.sum())) + 1) / 256.0).astype(np.float32)
        wild = pick > 0.999
        sh[wild] = (rng.uniform(1.0, 3.0, int(wild.sum())) * rng.choice([-1.0, 1.0], int(wild.sum()))).astype(np.float32)
    return CloudPlanes(n, sh_degree, pos, scl, rot, alp, col, sh)


def torch_cloud(n: int, sh_degree: int, device, seed: int = 1, slice_elems: int = 1 << 27) -> CloudPlanes:
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ws = float_plane_widths(sh_degree)
    ranges = ((-10.0, 10.0), (-7.0, 1.0), (-1.0, 1.0), (-6.0, 6.0), (-2.0, 2.0), (-0.5, 0.5))
    planes = []
    for w, (lo, hi) in zip(ws, ranges):
        t = torch.empty(n * w, dtype=torch.float32, device=device)
        for a in range(0, t.numel(), slice_elems):
            v = t[a:a + slice_elems]
            v.uniform_(lo, hi, generator=g)
        planes.append(t)
    rot = planes[2]
    if n:
        for a in range(0, n, slice_elems // 4):
            r = rot[4 * a:4 * (a + slice_elems // 4)].view(-1, 4)
            bad = r.norm(dim=1) < 1e-3
            r[bad] = torch.tensor([0.0, 0.0, 0.0, 1.0], device=device)
    sh = planes[5]
    for a in range(0, sh.numel(), slice_elems):
        v = sh[a:a + slice_elems]
        pick = torch.rand(v.numel(), device=device, generator=g)
        j = torch.randint(-128, 128, (v.numel(),), device=device, generator=g, dtype=torch.int32)
        tie = (2 * j + 1).to(torch.float32) / 256.0
        v.copy_(torch.where(pick < 0.01, tie, v))
        wild = torch.empty_like(v).uniform_(1.0, 3.0, generator=g) * torch.where(j >= 0, 1.0, -1.0)
        v.copy_(torch.where(pick > 0.999, wild, v))
        del pick, j, tie, wild
    return CloudPlanes(n, sh_degree, *planes)
