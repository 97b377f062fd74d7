"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/spz_b200.h
declares, its host-side helpers agree with the oracle, and every compute entry point fails loudly
(no CPU fallback) when there is no device."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import pytest

from spz_b200 import _native as N
from spz_b200 import codec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "spz_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spzb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 22
    L = N.lib()
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/spz_b200.h but not exported"
    # and the ctypes table covers exactly the header
    assert sorted(N.SIGNATURES) == names


def test_struct_layouts_match_header():
    # 64-bit count first, then 32-bit fields, then six pointers
    assert C.sizeof(N.Cloud) == 8 + 4 + 4 + 6 * 8
    assert C.sizeof(N.Packed) == 8 + 4 * 4 + 6 * 8
    assert C.sizeof(N.Timings) == 4 * 8 + 2 * 8 + 2 * 4 + 8 + 2 * 4
    assert N.lib().spzb200_version() == 200


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under spz_b200/ or include/ may reference it."""
    for base in ("spz_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".hpp")):
                    src = open(os.path.join(dirpath, f), errors="replace").read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
                    assert "spz_oracle" not in src and "libspz_ref" not in src, os.path.join(dirpath, f)


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_shard_ranges_partition_the_cloud(deg):
    tg = codec.tile_gaussians(deg)
    assert tg % 4 == 0 and tg > 0  # float slices 16-byte aligned, packed slices 4-byte aligned
    for n in (0, 1, tg - 1, tg, 7 * tg + 5, 10_000_000, 100_000_000, 3_000_000_007):
        for shards in (1, 2, 3, 4, 8):
            edges = [codec.shard_range(n, deg, shards, i) for i in range(shards)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a0, b0), (a1, b1) in zip(edges, edges[1:]):
                assert b0 == a1 and a0 <= b0
            for a, b in edges[:-1]:
                assert a % tg == 0 and b % tg == 0  # slice starts keep the vector path's alignment
            sizes = [b - a for a, b in edges]
            if n >= shards * tg:
                assert max(sizes) - min(sizes) <= 2 * tg  # balanced to within the granule + remainder
    with pytest.raises(N.CodecError):
        codec.shard_range(10, deg, 0, 0)
    with pytest.raises(N.CodecError):
        codec.shard_range(10, deg, 2, 2)


def test_flip_bits_match_oracle(oracle):
    for frm in range(9):
        for to in range(9):
            p, q, s = codec.flip_bits(frm, to)
            fp, fq, fsh = oracle.flips(frm, to)
            assert [(p >> i) & 1 for i in range(3)] == [int(v < 0) for v in fp]
            assert [(q >> i) & 1 for i in range(3)] == [int(v < 0) for v in fq]
            assert [(s >> i) & 1 for i in range(15)] == [int(v < 0) for v in fsh]


def test_host_built_tables_match_reference(oracle, golden):
    """The two libm-dependent tables a context uploads, against the reference-made golden values."""
    thr, lut = codec.build_tables()
    assert np.array_equal(thr[:255].view(np.uint32), golden["alpha_thresholds"])
    assert np.isposinf(thr[255])
    from oracle import bits
    assert np.array_equal(bits(lut), golden["table_alpha"])
    # first step at -6.23244762 (0xc0c77036), last at 6.23239183 (0x40c76fc1)
    assert thr[:1].view(np.uint32)[0] == 0xc0c77036 and thr[254:255].view(np.uint32)[0] == 0x40c76fc1


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(N.CodecError) as e:
        codec.Context(0)
    assert e.value.code == N.ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)
    # the multi-GPU host entry point reports the same, it does not compute on the host
    c = codec.alloc_cloud(4, 0, numpy_arrays=True)
    for p in c.planes():
        p[...] = 0
    with pytest.raises(N.CodecError) as e:
        codec.encode_host_multi([0], c)
    assert e.value.code == N.ERR_NO_DEVICE
    # a pooled lease fails the same way (and leaves the pool able to try again)
    for _ in range(2):
        with pytest.raises(N.CodecError) as e:
            codec.Context(0, pooled=True)
        assert e.value.code == N.ERR_NO_DEVICE


def test_argument_validation_needs_no_device():
    L = N.lib()
    assert L.spzb200_create(0, None) == N.ERR_INVALID
    assert L.spzb200_encode_device(None, None, 0, None, None) == N.ERR_INVALID
    assert L.spzb200_decode_host(None, None, 0, None, None) == N.ERR_INVALID
    assert b"null context" in L.spzb200_last_error()
    assert L.spzb200_unpack_records_host(None, None, 1, 3, 12, None, None) == N.ERR_INVALID
    assert L.spzb200_unpack_gather_host(None, None, None, 1, None, None) == N.ERR_INVALID
    assert L.spzb200_unpack_gather_device(None, None, None, 1, None, None, None) == N.ERR_INVALID
    assert L.spzb200_encode_host_as(None, None, 0, 2, None, None) == N.ERR_INVALID
    assert L.spzb200_acquire(0, None) == N.ERR_INVALID
    L.spzb200_release(None)  # a no-op, like free(NULL)
    with pytest.raises(TypeError):
        codec._cloud_struct(codec.CloudPlanes(1, 0, *[np.zeros(3, np.float64)] * 6), False)
    with pytest.raises(ValueError):
        codec._cloud_struct(codec.CloudPlanes(2, 0, *[np.zeros(3, np.float32)] * 6), False)
    with pytest.raises(ValueError):
        codec._cloud_struct(codec.CloudPlanes(1, 5, *[np.zeros(3, np.float32)] * 6), False)
