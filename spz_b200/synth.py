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


# -------------------------------------------------------------------------------------------------
# Counter-based generator ("counter-based so CPU and GPU shards can generate
# identical data").  Every float is a pure function of (seed, plane, GLOBAL element index), so rank r
# of an N-rank run generates exactly slice [a_r, b_r) of the one cloud a single-GPU run holds, and the
# numpy and torch versions agree bit for bit (integer hashing, then exactly rounded float32 ops).
# Same distribution as above.
# -------------------------------------------------------------------------------------------------
_M32 = 0xFFFFFFFF
_RANGES = ((-10.0, 10.0), (-7.0, 1.0), (-1.0, 1.0), (-6.0, 6.0), (-2.0, 2.0), (-0.5, 0.5))


def _hash32(idx, key: int):
    """lowbias32-style avalanche of a 64-bit element index and a 32-bit key; int64 arrays/tensors holding
    values in [0, 2^32).  Products wrap in int64, which leaves their low 32 bits intact."""
    x = (idx & _M32) ^ (((idx >> 32) * 0x9E3779B1) & _M32) ^ key
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32
    x = x ^ (x >> 16)
    return x


def _plane_key(seed: int, plane: int, stream: int) -> int:
    return (seed * 0x9E3779B1 + plane * 0x85EBCA6B + stream * 0xC2B2AE35 + 0x27D4EB2F) & _M32


def _counter_values(xp, idx, seed, plane, f32):
    """float32 values of global elements `idx` (int64) of plane `plane`; xp = numpy or torch."""
    lo, hi = _RANGES[plane]
    u = f32(_hash32(idx, _plane_key(seed, plane, 0)) >> 8) * f32(2.0 ** -24)  # [0, 1), exact
    v = u * f32(hi - lo) + f32(lo)  # two roundings, the same in numpy and in torch's separate kernels
    if plane == 5:
        pick = _hash32(idx, _plane_key(seed, plane, 1))
        j = _hash32(idx, _plane_key(seed, plane, 2))
        tie = f32(2 * ((j & 0xFF) - 128) + 1) * f32(1.0 / 256.0)           # exact rounding ties of the 8-bit grid
        wild = (f32((j >> 8) & 0xFFFF) * f32(2.0 / 65536.0) + f32(1.0))    # [1, 3)
        wild = xp.where((j >> 24) & 1 == 1, -wild, wild)
        v = xp.where(pick < int(0.01 * 2 ** 32), tie, v)
        v = xp.where(pick > int(0.999 * 2 ** 32), wild, v)
    return v


def _fix_rotations(xp, r):
    """quaternions whose every component is below 1e-3 in magnitude (probability 1e-12) become the identity"""
    q = r.reshape(-1, 4)
    bad = (abs(q) < 1e-3).all(1) if xp is np else (q.abs() < 1e-3).all(dim=1)
    if bool(bad.any()):
        q[bad] = xp.asarray([0, 0, 0, 1], dtype=q.dtype) if xp is np else xp.tensor([0.0, 0.0, 0.0, 1.0], device=q.device)


def counter_cloud_numpy(n_total: int, sh_degree: int, a: int = 0, b: int | None = None, seed: int = 1) -> CloudPlanes:
    """Gaussians [a, b) of the seeded n_total-point cloud, on the host."""
    b = n_total if b is None else b
    ws = float_plane_widths(sh_degree)
    planes = []
    for plane, w in enumerate(ws):
        idx = np.arange(a * w, b * w, dtype=np.int64)
        planes.append(np.ascontiguousarray(_counter_values(np, idx, seed, plane, lambda x: np.asarray(x).astype(np.float32)), np.float32))
    if b > a:
        _fix_rotations(np, planes[2])
    return CloudPlanes(b - a, sh_degree, *planes)


def counter_cloud_torch(n_total: int, sh_degree: int, device, a: int = 0, b: int | None = None, seed: int = 1,
                        slice_elems: int = 1 << 26, out: CloudPlanes | None = None) -> CloudPlanes:
    """Gaussians [a, b) of the same cloud on `device`, bit-identical to counter_cloud_numpy; generated in
    bounded slices (the int64 scratch of a slice is ~0.5 GB per temporary)."""
    import torch
    b = n_total if b is None else b
    ws = float_plane_widths(sh_degree)
    planes = []
    for plane, w in enumerate(ws):
        t = out.planes()[plane] if out is not None else torch.empty((b - a) * w, dtype=torch.float32, device=device)
        first = a * w
        for s in range(0, t.numel(), slice_elems):
            e = min(t.numel(), s + slice_elems)
            idx = torch.arange(first + s, first + e, dtype=torch.int64, device=device)
            t[s:e] = _counter_values(torch, idx, seed, plane, lambda x: x.to(torch.float32) if hasattr(x, "to") else torch.tensor(float(x), dtype=torch.float32, device=device))
            del idx
        planes.append(t)
    if b > a:
        for s in range(0, b - a, slice_elems // 4):
            _fix_rotations(torch, planes[2][4 * s:4 * min(b - a, s + slice_elems // 4)])
    return out if out is not None else CloudPlanes(b - a, sh_degree, *planes)
