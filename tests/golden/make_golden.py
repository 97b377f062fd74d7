"""Generates tests/golden/golden.npz by running the UNMODIFIED reference (oracle/_ref/libspz_ref.so,
built in place from /root/reference/src/cc by oracle/Makefile).  Run in the build container:

    make -C oracle ref && python tests/golden/make_golden.py

The .npz is committed; the tests never need /root/reference.  Contents (all produced by the
reference itself, nothing by the oracle restatement or by this repo's kernels):

  kat_*      the reference test-suite's canonical 2-gaussian fixture (tests/python/load_spz_test.py
             :72-100): packed planes for from=UNSPECIFIED and from=RDF, decoded float bits for
             to=UNSPECIFIED and to=RUF, and the gzip container saveSpz() writes for it.
  shedge_*   the reference's only known-answer test, test_sh_encoding_for_zeros_and_edges (:180-207).
  c<deg>_*   seeded 257-point clouds per SH degree (tame and with NaN/Inf/huge specials): inputs,
             reference-packed planes for three `from` systems, reference-decoded bits for three `to`.
  s<ver>_*   random byte streams of version 1/2/3 (any bytes are a valid stream) with several
             fractionalBits values and their reference decodes.
  alpha_*    the alpha step function sampled at and just below each of its 255 thresholds.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import Cloud, Packed, Ref, SH_DIM, bits  # noqa: E402

SPECIALS = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e30, -1e30, 3e9, -3e9, 2147483520.0,
                     2147483392.0, 1e-40, 524288.0, -524288.0, 2047.99, -2048.0, 0.49999997,
                     0.5, -0.5, 1.5 / 128, 2.5 / 128, -2.5 / 128], np.float32)


def kat_cloud() -> Cloud:
    return Cloud(2, 3,
                 np.array([0, .1, -.2, .3, .4, .5], np.float32),
                 np.array([-3, -2, -1.5, -1, 0, .1], np.float32),
                 np.array([-.5, .2, 1, -.2, .1, -.4, -.3, .5], np.float32),
                 np.array([-1, 1], np.float32),
                 np.array([-1, 0, 1, -.5, .5, .1], np.float32),
                 (np.arange(90, dtype=np.float32) / 45.0 - 1.0).astype(np.float32), True)


def random_cloud(rng, n, deg, special) -> Cloud:
    d = SH_DIM[deg] * 3

    def f(k, lo, hi):
        a = rng.uniform(lo, hi, k).astype(np.float32)
        if special and k:
            idx = rng.integers(0, k, max(1, k // 12))
            a[idx] = rng.choice(SPECIALS, idx.size)
        return a

    return Cloud(n, deg, f(3 * n, -10, 10), f(3 * n, -12, 8), f(4 * n, -1, 1), f(n, -8, 8),
                 f(3 * n, -4, 4), f(d * n, -1.2, 1.2))


def put_cloud(out, key, c: Cloud):
    for name, p in zip("positions scales rotations alphas colors sh".split(), c.planes()):
        out[f"{key}_{name}"] = np.ascontiguousarray(p, np.float32).view(np.uint32)  # bit patterns


def put_packed(out, key, p: Packed):
    for name, a in zip("positions scales rotations alphas colors sh".split(), p.planes()):
        out[f"{key}_{name}"] = a


def put_decoded(out, key, c: Cloud):
    for name, p in zip("positions scales rotations alphas colors sh".split(), c.planes()):
        out[f"{key}_{name}"] = bits(p)  # NaNs canonicalised


def main():
    R = Ref()
    rng = np.random.default_rng(20261018)
    out = {}

    k = kat_cloud()
    put_cloud(out, "kat_in", k)
    for frm in (0, 6):
        put_packed(out, f"kat_pack_from{frm}", R.pack(k, frm))
    p0 = R.pack(k, 0)
    for to in (0, 8):
        put_decoded(out, f"kat_unpack_to{to}", R.unpack(p0, to))
    out["kat_spz_gzip"] = np.frombuffer(R.save_spz(k, 0), np.uint8)
    out["kat_container"] = np.frombuffer(R.serialize(k, 0), np.uint8)

    edge = np.array([-.01, 0, .01, -1, -.99, -.95, .95, .99, 1], np.float32)
    ec = Cloud(1, 1, np.zeros(3, np.float32), np.zeros(3, np.float32), np.array([0, 0, 0, 1], np.float32),
               np.zeros(1, np.float32), np.zeros(3, np.float32), edge)
    ep = R.pack(ec, 0)
    out["shedge_in"] = edge
    out["shedge_bytes"] = ep.sh
    out["shedge_decoded"] = bits(R.unpack(ep, 0).sh)

    for deg in range(4):
        for tag, special in (("tame", False), ("wild", True)):
            c = random_cloud(rng, 257, deg, special)
            key = f"c{deg}_{tag}"
            put_cloud(out, key + "_in", c)
            for frm in (0, 6, 7):
                pk = R.pack(c, frm)
                put_packed(out, f"{key}_pack_from{frm}", pk)
            pk = R.pack(c, 0)
            for to in (0, 6, 8):
                put_decoded(out, f"{key}_unpack_to{to}", R.unpack(pk, to))

    for ver in (1, 2, 3):
        for fb in (12, 0, 5, 31, 35):
            n, deg = 131, 3 if ver == 3 else 2
            d = SH_DIM[deg] * 3
            pk = Packed(n, deg, fb, ver,
                        rng.integers(0, 256, n * (6 if ver == 1 else 9)).astype(np.uint8),
                        rng.integers(0, 256, n * 3).astype(np.uint8),
                        rng.integers(0, 256, n * (4 if ver == 3 else 3)).astype(np.uint8),
                        rng.integers(0, 256, n).astype(np.uint8),
                        rng.integers(0, 256, n * 3).astype(np.uint8),
                        rng.integers(0, 256, n * d).astype(np.uint8))
            key = f"s{ver}_fb{fb}"
            put_packed(out, key + "_in", pk)
            for to in (0, 7, 8):
                put_decoded(out, f"{key}_unpack_to{to}", R.unpack(pk, to))

    # alpha step function: for every byte level, the smallest float reaching it and its predecessor
    def alpha_bytes(a):
        n = a.size
        c = Cloud(n, 0, np.zeros(3 * n, np.float32), np.zeros(3 * n, np.float32),
                  np.tile(np.array([0, 0, 0, 1], np.float32), n), a.astype(np.float32),
                  np.zeros(3 * n, np.float32), np.zeros(0, np.float32))
        return R.pack(c, 0).alphas

    def key_of(b):
        b = np.uint32(b)
        return (~b) & np.uint32(0xFFFFFFFF) if b & np.uint32(0x80000000) else b | np.uint32(0x80000000)

    def float_of(k):
        k = np.uint32(k)
        b = (k & np.uint32(0x7FFFFFFF)) if k & np.uint32(0x80000000) else (~k) & np.uint32(0xFFFFFFFF)
        return np.array([b], np.uint32).view(np.float32)[0]

    lo_k, hi_k = int(key_of(0xFF800000)), int(key_of(0x7F800000))
    thr = np.zeros(255, np.float32)
    for level in range(1, 256):
        lo, hi = lo_k, hi_k
        while lo < hi:
            mid = (lo + hi) // 2
            if alpha_bytes(np.array([float_of(mid)], np.float32))[0] >= level:
                hi = mid
            else:
                lo = mid + 1
        thr[level - 1] = float_of(lo)
    below = np.array([float_of(int(key_of(int(np.array([t]).view(np.uint32)[0]))) - 1) for t in thr], np.float32)
    out["alpha_thresholds"] = thr.view(np.uint32)
    out["alpha_at_threshold"] = alpha_bytes(thr)
    out["alpha_below_threshold"] = alpha_bytes(below)
    # decode tables, through the reference's unpack
    n = 256
    pk = Packed(n, 0, 12, 3, np.zeros(9 * n, np.uint8), np.repeat(np.arange(256, dtype=np.uint8), 3)[: 3 * n] * 0,
                np.zeros(4 * n, np.uint8), np.arange(256, dtype=np.uint8), np.zeros(3 * n, np.uint8), np.zeros(0, np.uint8))
    pk.scales = np.repeat(np.arange(256, dtype=np.uint8), 3)
    pk.colors = np.repeat(np.arange(256, dtype=np.uint8), 3)
    g = R.unpack(pk, 0)
    out["table_alpha"] = bits(g.alphas)
    out["table_scale"] = bits(g.scales[::3])
    out["table_color"] = bits(g.colors[::3])

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
