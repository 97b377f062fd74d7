"""Generates tests/golden/golden_unpack_at.npz with the UNMODIFIED reference (oracle/_ref/libspz_ref.so):
PackedGaussians::unpack(i, c) (load-spz.cc:383-463) for every stream flavour and SH degree, with the
identity converter, coordinateConverter(RUB, RDF) and a hand-built converter of arbitrary factors.

    make -C oracle ref && python tests/golden/make_golden_unpack_at.py

The .npz is committed; the tests never need /root/reference.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import Oracle, Ref, bits  # noqa: E402
from util import PLANES, random_stream  # noqa: E402


def main():
    R = Ref()
    flips = Oracle()  # only for the +-1 converter tables, which tests/test_cxx_api.py pins against the reference
    rng = np.random.default_rng(20261019)
    odd = (rng.normal(size=21) * 3).astype(np.float32)
    odd[[2, 9]] = [0.0, -0.0]
    convs = np.stack([flips.converter(0, 0), flips.converter(4, 6), odd])
    out = {"converters": convs}
    n = 64
    for ver in (1, 2, 3, 4):
        for deg in range(4):
            fb = int(rng.choice([0, 5, 12, 20, 35]))
            s = random_stream(rng, n, deg, ver, fb)
            if ver >= 3:
                s.rotations.view("<u4")[::2] &= np.uint32(0xEFFBFEFF)  # every other one a valid unit quaternion
            idx = np.concatenate([[0, n - 1, n - 1], rng.integers(0, n, 29)]).astype(np.int64)
            key = f"v{ver}d{deg}"
            out[f"{key}_fb"] = np.int32(fb)
            out[f"{key}_idx"] = idx
            for name, a in zip(PLANES, s.planes()):
                out[f"{key}_{name}"] = a
            for k, conv in enumerate(convs):
                out[f"{key}_out{k}"] = bits(R.unpack_at(s, idx, conv))
    path = os.path.join(ROOT, "tests", "golden", "golden_unpack_at.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
