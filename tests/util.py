"""Shared helpers for the test-suite."""
from __future__ import annotations

import numpy as np

from oracle import Cloud, Packed, SH_DIM, bits

PLANES = "positions scales rotations alphas colors sh".split()

SPECIALS = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e30, -1e30, 3e9, -3e9, 2147483520.0,
                     2147483392.0, 1e-40, 524288.0, -524288.0, 2047.99, -2048.0, 0.49999997,
                     0.5, -0.5, 1.5 / 128, 2.5 / 128, -2.5 / 128], np.float32)


def golden_cloud(G, key, n, deg) -> Cloud:
    return Cloud(n, deg, *[G[f"{key}_{p}"].view(np.float32) for p in PLANES])


def golden_packed(G, key, n, deg, fb=12, ver=3) -> Packed:
    return Packed(n, deg, fb, ver, *[G[f"{key}_{p}"] for p in PLANES])


def random_cloud(rng, n, deg, special=False) -> Cloud:
    d = SH_DIM[deg] * 3

    def f(k, lo, hi):
        a = rng.uniform(lo, hi, k).astype(np.float32)
        if special and k:
            idx = rng.integers(0, k, max(1, k // 12))
            a[idx] = rng.choice(SPECIALS, idx.size)
        return a

    return Cloud(n, deg, f(3 * n, -10, 10), f(3 * n, -12, 8), f(4 * n, -1, 1), f(n, -8, 8),
                 f(3 * n, -4, 4), f(d * n, -1.2, 1.2))


def random_stream(rng, n, deg, ver, fb=12) -> Packed:
    d = SH_DIM[deg] * 3
    r = lambda k: rng.integers(0, 256, k).astype(np.uint8)  # noqa: E731
    return Packed(n, deg, fb, ver, r(n * (6 if ver in (1, 4) else 9)), r(n * 3), r(n * (4 if ver >= 3 else 3)),
                  r(n), r(n * 3), r(n * d))


def assert_packed_equal(a, b, what=""):
    for name, x, y in zip(PLANES, a.planes(), b.planes()):
        x, y = np.asarray(x), np.asarray(y)
        assert x.shape == y.shape, f"{what} {name}: shape {x.shape} vs {y.shape}"
        if not np.array_equal(x, y):
            i = np.flatnonzero(x != y)
            raise AssertionError(f"{what} plane {name}: {i.size} byte mismatches, first at {i[:5]}: {x[i[:5]]} vs {y[i[:5]]}")


def assert_cloud_bits_equal(a, b, what=""):
    for name, x, y in zip(PLANES, a.planes(), b.planes()):
        bx = bits(np.asarray(x)) if np.asarray(x).dtype == np.float32 else np.asarray(x)
        by = bits(np.asarray(y)) if np.asarray(y).dtype == np.float32 else np.asarray(y)
        assert bx.shape == by.shape, f"{what} {name}: shape {bx.shape} vs {by.shape}"
        if not np.array_equal(bx, by):
            i = np.flatnonzero(bx != by)
            raise AssertionError(f"{what} plane {name}: {i.size} float-bit mismatches, first at {i[:5]}: "
                                 f"{[hex(v) for v in bx[i[:5]]]} vs {[hex(v) for v in by[i[:5]]]}")


def cloud_to_ply_rows(c, names):
    """[n, len(names)] float32 vertex records of a .ply with the given property order (the layout
    load-spz.cc:858-890 writes, flips aside); properties the cloud does not carry are 0."""
    n = c.n if hasattr(c, "n") else c.num_points
    d = c.sh.size // (3 * n) if n else 0
    cols = {"x": c.positions[0::3], "y": c.positions[1::3], "z": c.positions[2::3],
            "f_dc_0": c.colors[0::3], "f_dc_1": c.colors[1::3], "f_dc_2": c.colors[2::3], "opacity": c.alphas,
            "scale_0": c.scales[0::3], "scale_1": c.scales[1::3], "scale_2": c.scales[2::3],
            "rot_0": c.rotations[3::4], "rot_1": c.rotations[0::4], "rot_2": c.rotations[1::4], "rot_3": c.rotations[2::4]}
    sh = c.sh.reshape(n, d, 3) if d else np.zeros((n, 0, 3), np.float32)
    for ch in range(3):
        for s in range(d):
            cols[f"f_rest_{ch * d + s}"] = sh[:, s, ch]
    zero = np.zeros(n, np.float32)
    return np.ascontiguousarray(np.stack([cols.get(k, zero) for k in names], axis=1), np.float32)


def rotation_guard_stress(rng, n):
    """4n floats of quaternions that sit on and around the boundaries of quant_rotation_smallest3's fast-path
    guard (codec_math.cuh): squared norms near 2^-40 and 2^40, components near 2^-60 and denormal, exact
    and negative zeros, norms whose significand is all ones, near-ties between components, components at
    sqrt(1/2) where the 9-bit magnitude saturates, plus plain random ones at assorted scales."""
    q = rng.normal(size=(n, 4)).astype(np.float32)
    k = n // 10
    scale = np.ones(n, np.float32)
    scale[0 * k:1 * k] = np.float32(2.0) ** rng.uniform(-21, -19, k).astype(np.float32)      # n2 around 2^-40
    scale[1 * k:2 * k] = np.float32(2.0) ** rng.uniform(19, 21, k).astype(np.float32)        # n2 around 2^40
    scale[2 * k:3 * k] = np.float32(2.0) ** rng.integers(-70, 60, k).astype(np.float32)
    q *= scale[:, None]
    tiny = np.float32(2.0) ** rng.uniform(-62, -58, k).astype(np.float32)                     # components around 2^-60
    q[3 * k:4 * k, rng.integers(0, 4, k)] = 0
    q[3 * k:4 * k, 0] = tiny * np.where(rng.random(k) < 0.5, -1, 1).astype(np.float32)
    q[4 * k:5 * k, 1] = np.float32(1e-42) * rng.integers(-5, 6, k).astype(np.float32)         # denormal component
    z = rng.integers(0, 4, k)
    q[np.arange(5 * k, 6 * k), z] = np.where(rng.random(k) < 0.5, -0.0, 0.0).astype(np.float32)
    # norm with an all-ones significand: a single non-zero component of that value
    ones = ((rng.integers(100, 150, k).astype(np.uint32) << 23) | np.uint32(0x7fffff)).view(np.float32)
    q[6 * k:7 * k] = 0
    q[np.arange(6 * k, 7 * k), rng.integers(0, 4, k)] = ones * np.where(rng.random(k) < 0.5, -1, 1).astype(np.float32)
    # near-ties for the largest component and the sqrt(1/2) saturation point
    base = rng.normal(size=k).astype(np.float32)
    q[7 * k:8 * k, 0] = base
    q[7 * k:8 * k, 2] = np.nextafter(base, np.float32(np.inf)) * np.where(rng.random(k) < 0.5, -1, 1).astype(np.float32)
    q[8 * k:9 * k] = 0
    q[8 * k:9 * k, 1] = rng.normal(size=k).astype(np.float32)
    q[8 * k:9 * k, 3] = q[8 * k:9 * k, 1] * np.where(rng.random(k) < 0.5, -1, 1).astype(np.float32)
    return np.ascontiguousarray(q.reshape(-1))
