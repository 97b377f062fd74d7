"""The device math header (spz_b200/csrc/codec_math.cuh) compiled for the HOST with its intrinsics
emulated (tests/host_emul/quant_host.cc), swept against the oracle.  This checks the integer
reformulations the kernels rely on (round-half-away via add.rz, AND instead of /b*b, sign-bit XOR
instead of *-1, the alpha threshold search) without a GPU; the -m gpu tests repeat the sweeps on
the real kernels.  Test-only: the product never runs this code on the CPU."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from oracle import Cloud, Packed, bits

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


def _thresholds():
    from spz_b200.codec import build_tables
    return build_tables()[0]


def _sweep(emul, which, first, stride, count, thr=None):
    out = np.zeros(count, np.uint8)
    t = thr.ctypes.data_as(_f32p) if thr is not None else None
    emul.emul_sweep_u8(C.c_int(which), C.c_uint32(first), C.c_uint32(stride), C.c_int64(count), t, out.ctypes.data_as(_u8p))
    return out


# strided sweeps over the whole 2^32 float space (NaN/Inf/denormals included) + dense windows
# around the values where the quantizers actually switch
@pytest.mark.parametrize("which,name", [(1, "scale"), (2, "colour"), (3, "sh b8"), (4, "sh b16")])
def test_u8_quantizers_strided_full_range(emul, oracle, which, name):
    stride, count = 4099, (1 << 32) // 4099
    for first in (0, 1234567):
        a = _sweep(emul, which, first, stride, count)
        b = oracle.sweep_u8(which, first, stride, count)
        assert np.array_equal(a, b), name


@pytest.mark.parametrize("which,lo,hi", [(1, -12.0, 7.0), (2, -3.5, 3.5), (3, -1.1, 1.1), (4, -1.1, 1.1)])
def test_u8_quantizers_dense_windows(emul, oracle, which, lo, hi):
    """Every float in small windows of the active range, both signs (the ties live here)."""
    rng = np.random.default_rng(which)
    for _ in range(6):
        x = np.float32(rng.uniform(lo, hi))
        first = int(np.array([x], np.float32).view(np.uint32)[0])
        count = 1 << 20
        assert np.array_equal(_sweep(emul, which, first, 1, count), oracle.sweep_u8(which, first, 1, count))


def test_sh_flipped_equals_oracle_of_negated(emul, oracle):
    """quant_sh(x, -128) must equal the reference's quantizeSH(-1.0f * x): sign-bit XOR == * -1."""
    count = 1 << 21
    for which_flipped, which in ((5, 3), (6, 4)):
        for first in (0x3a000000, 0xbb000000, 0x3f000000, 0x7f7ffff0, 0x00000000):
            a = _sweep(emul, which_flipped, first, 1, count)
            b = oracle.sweep_u8(which, first ^ 0x80000000, 1, count)
            assert np.array_equal(a, b)


def test_alpha_threshold_search_matches_oracle(emul, oracle):
    thr = _thresholds()
    stride, count = 1021, (1 << 32) // 1021
    assert np.array_equal(_sweep(emul, 0, 0, stride, count, thr), oracle.sweep_u8(0, 0, stride, count))
    # all floats in a window around each of a few thresholds
    for t in thr[[0, 1, 64, 127, 128, 200, 254]]:
        first = int(np.array([t], np.float32).view(np.uint32)[0]) - 4096
        assert np.array_equal(_sweep(emul, 0, first, 1, 8192, thr), oracle.sweep_u8(0, first, 1, 8192))


def test_positions_encode(emul, oracle):
    rng = np.random.default_rng(7)
    vals = np.concatenate([
        rng.uniform(-2100, 2100, 60000).astype(np.float32),
        (rng.integers(-(1 << 24), 1 << 24, 30000) / 8192.0).astype(np.float32),  # exact .5 ties in 1/4096 units
        np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 524288.0, -524288.0, 524287.97, 1e30, -1e30, 2047.9999,
                  -2048.0, 1e-40, 0.00012207031, -0.00012207031], np.float32)])
    n = vals.size // 3
    vals = np.ascontiguousarray(vals[:3 * n])
    c = Cloud(n, 0, vals, np.zeros(3 * n, np.float32), np.tile(np.array([0, 0, 0, 1], np.float32), n),
              np.zeros(n, np.float32), np.zeros(3 * n, np.float32), np.zeros(0, np.float32))
    # from=LDF(5) vs RUB(4): (5-1)^(4-1) = 7 -> all three axes flip; from=0 -> none
    for flip, frm in ((0, 0), (1, 5)):
        p = oracle.pack(c, frm).positions.reshape(-1, 3).astype(np.uint32)
        want = p[:, 0] | (p[:, 1] << 8) | (p[:, 2] << 16)
        assert np.array_equal(_positions(emul, vals.view(np.uint32), flip), want)


def _positions(emul, u32vals, flip):
    out = np.zeros(u32vals.size, np.uint32)
    one = np.zeros(1, np.uint32)
    f = emul.emul_sweep_position
    for i, b in enumerate(u32vals.tolist()):
        f(C.c_uint32(b), C.c_uint32(0), C.c_int64(1), C.c_int(flip), one.ctypes.data_as(_u32p))
        out[i] = one[0]
    return out


def test_positions_encode_bit_pattern_sweep(emul, oracle):
    """Strided sweep over all float bit patterns through the cloud-level oracle."""
    stride = 65537
    count = ((1 << 32) // stride) // 3 * 3
    u = (np.arange(count, dtype=np.uint64) * stride).astype(np.uint32)
    vals = u.view(np.float32)
    n = count // 3
    c = Cloud(n, 0, vals, np.zeros(3 * n, np.float32), np.tile(np.array([0, 0, 0, 1], np.float32), n),
              np.zeros(n, np.float32), np.zeros(3 * n, np.float32), np.zeros(0, np.float32))
    for flip, frm in ((0, 0), (1, 5)):
        p = oracle.pack(c, frm).positions.reshape(-1, 3).astype(np.uint32)
        want = p[:, 0] | (p[:, 1] << 8) | (p[:, 2] << 16)
        got = np.zeros(count, np.uint32)
        emul.emul_sweep_position(C.c_uint32(0), C.c_uint32(stride), C.c_int64(count), C.c_int(flip), got.ctypes.data_as(_u32p))
        assert np.array_equal(got, want)


def _rot_cloud(rot):
    n = rot.size // 4
    z = np.zeros(3 * n, np.float32)
    return Cloud(n, 0, z, z, rot, np.zeros(n, np.float32), z, np.zeros(0, np.float32))


def test_rotations_encode(emul, oracle):
    rng = np.random.default_rng(11)
    n = 300000
    rot = rng.uniform(-1, 1, 4 * n).astype(np.float32)
    # ties between components, zeros, negatives zeros, denormal and huge norms, NaN/Inf, zero quaternion
    rot[:4] = [0.5, 0.5, 0.5, 0.5]
    rot[4:8] = [-0.5, 0.5, -0.5, 0.5]
    rot[8:12] = [0, 0, 0, 0]
    rot[12:16] = [np.nan, 1, 0, 0]
    rot[16:20] = [np.inf, 1, 0, 0]
    rot[20:24] = [1e-30, 1e-30, 1e-30, 1e-30]
    rot[24:28] = [1e25, -1e25, 1e20, 3]
    rot[28:32] = [-0.0, 0.0, -0.0, -1.0]
    rot[32:36] = [0.70710678, 0.70710678, 0, 0]
    rot[36:40] = [1, 0, 0, 0]
    sel = rng.integers(0, n, 2000)
    rot.reshape(-1, 4)[sel, rng.integers(0, 4, 2000)] = 0.0
    for frm in (0, 5, 6, 7, 8):
        from spz_b200.codec import flip_bits
        _, fq, _ = flip_bits(frm, 4)
        got = np.zeros(n, np.uint32)
        emul.emul_rotations(C.c_int64(n), rot.ctypes.data_as(_f32p), C.c_uint32(fq), got.ctypes.data_as(_u32p))
        want = oracle.pack(_rot_cloud(rot), frm).rotations.view("<u4")
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, (frm, bad[:5], rot.reshape(-1, 4)[bad[:5]])


def test_rotations_encode_guard_boundaries(emul, oracle):
    """The fast path of quant_rotation_smallest3 and its fallback meet exactly at the guard: quaternions on
    and around every boundary of it, against the oracle."""
    from spz_b200.codec import flip_bits
    from util import rotation_guard_stress
    rng = np.random.default_rng(12)
    n = 400000
    rot = rotation_guard_stress(rng, n)
    for frm in (0, 7):
        _, fq, _ = flip_bits(frm, 4)
        got = np.zeros(n, np.uint32)
        emul.emul_rotations(C.c_int64(n), rot.ctypes.data_as(_f32p), C.c_uint32(fq), got.ctypes.data_as(_u32p))
        want = oracle.pack(_rot_cloud(rot), frm).rotations.view("<u4")
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, (frm, bad[:5], rot.reshape(-1, 4)[bad[:5]])


def test_rotations_decode_all_streams(emul, oracle):
    """Every 30-bit payload class: all 2^20 combos of two fields with the third random, x 4 index."""
    rng = np.random.default_rng(13)
    n = 1 << 21
    comp = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    comp[:1 << 20] = (np.arange(1 << 20, dtype=np.uint32) << 10) | (comp[:1 << 20] & 0xC00003FF)
    pk = Packed(n, 0, 12, 3, np.zeros(9 * n, np.uint8), np.zeros(3 * n, np.uint8), comp.view(np.uint8).copy(),
                np.zeros(n, np.uint8), np.zeros(3 * n, np.uint8), np.zeros(0, np.uint8))
    from spz_b200.codec import flip_bits
    for to in (0, 6, 7):
        _, fq, _ = flip_bits(4, to)
        got = np.zeros(4 * n, np.float32)
        emul.emul_unrotations_s3(C.c_int64(n), comp.ctypes.data_as(_u32p), C.c_uint32(fq), got.ctypes.data_as(_f32p))
        want = oracle.unpack(pk, to).rotations
        assert np.array_equal(bits(got), bits(want))


def test_rotations_decode_first_three_exhaustive(emul, oracle):
    n = 1 << 24  # every (b0, b1, b2)
    b = np.arange(n, dtype=np.uint32)
    rb = np.stack([b & 255, (b >> 8) & 255, b >> 16], axis=1).astype(np.uint8).reshape(-1)
    pk = Packed(n, 0, 12, 2, np.zeros(9 * n, np.uint8), np.zeros(3 * n, np.uint8), rb,
                np.zeros(n, np.uint8), np.zeros(3 * n, np.uint8), np.zeros(0, np.uint8))
    from spz_b200.codec import flip_bits
    for to in (0, 7):
        _, fq, _ = flip_bits(4, to)
        got = np.zeros(4 * n, np.float32)
        emul.emul_unrotations_f3(C.c_int64(n), rb.ctypes.data_as(_u8p), C.c_uint32(fq), got.ctypes.data_as(_f32p))
        assert np.array_equal(bits(got), bits(oracle.unpack(pk, to).rotations))


def test_dequant_tables_and_half(emul, oracle):
    sc = np.zeros(256, np.float32); co = np.zeros(256, np.float32); sh = np.zeros(256, np.float32)
    shf = np.zeros(256, np.float32); half = np.zeros(65536, np.float32)
    emul.emul_dequant_tables(*[a.ctypes.data_as(_f32p) for a in (sc, co, sh, shf, half)])
    assert np.array_equal(bits(sc), bits(oracle.dequant_table("scale")))
    assert np.array_equal(bits(co), bits(oracle.dequant_table("color")))
    assert np.array_equal(bits(sh), bits(oracle.dequant_table("sh")))
    assert np.array_equal(bits(shf), bits(-oracle.dequant_table("sh")))  # x*-1 incl. 0 -> -0
    want = np.array([oracle.lib.oracle_half_to_float(h) for h in range(65536)], np.float32)
    assert np.array_equal(bits(half), bits(want))


def test_positions_decode_every_24bit_code(emul, oracle):
    n3 = 1 << 24
    codes = np.arange(n3, dtype=np.uint32)
    n = n3 // 3 + 1
    codes3 = np.resize(codes, 3 * n)
    pb = np.stack([codes3 & 255, (codes3 >> 8) & 255, codes3 >> 16], axis=1).astype(np.uint8).reshape(-1)
    for fb, to in ((12, 0), (12, 5), (0, 0), (31, 5), (35, 0), (7, 8)):
        pk = Packed(n, 0, fb, 3, pb, np.zeros(3 * n, np.uint8), np.zeros(4 * n, np.uint8),
                    np.zeros(n, np.uint8), np.zeros(3 * n, np.uint8), np.zeros(0, np.uint8))
        want = oracle.unpack(pk, to).positions
        one = np.int32(np.uint32(1 << (fb & 31)).astype(np.int32))
        scale = np.float32(1.0 / float(one))
        # RUB(4) -> LDF(5) flips all three axes; the emulated kernel folds the sign into the scale
        s = -scale if to == 5 else scale
        got = np.zeros(3 * n, np.float32)
        if to == 8:  # RUF: only z flips
            for ax in range(3):
                sub = np.ascontiguousarray(codes3[ax::3])
                g = np.zeros(sub.size, np.float32)
                emul.emul_positions_decode(C.c_int64(sub.size), sub.ctypes.data_as(_u32p), C.c_float(-scale if ax == 2 else scale), g.ctypes.data_as(_f32p))
                got[ax::3] = g
        else:
            emul.emul_positions_decode(C.c_int64(3 * n), codes3.ctypes.data_as(_u32p), C.c_float(s), got.ctypes.data_as(_f32p))
        assert np.array_equal(bits(got), bits(want)), (fb, to)


def test_flip_bits_all_pairs(emul, oracle):
    for frm in range(9):
        for to in range(9):
            p, q, s = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
            emul.emul_flip_bits(frm, to, C.byref(p), C.byref(q), C.byref(s))
            fp, fq, fsh = oracle.flips(frm, to)
            assert [(p.value >> i) & 1 for i in range(3)] == [int(v < 0) for v in fp]
            assert [(q.value >> i) & 1 for i in range(3)] == [int(v < 0) for v in fq]
            assert [(s.value >> i) & 1 for i in range(15)] == [int(v < 0) for v in fsh]


@pytest.mark.parametrize("nbytes", [3, 6, 9, 24, 45])
def test_record_realignment(emul, nbytes):
    """record_align.cuh: per-gaussian records of 3 / 6 / 9 / 24 / 45 bytes in and out of a packed plane
    with whole-word accesses (funnel shifts + the predecessor's tail), lane after lane on the host.
    128 gaussians = one tile = four warps; the plane must come out byte-identical to the plain
    concatenation of the records, nothing may be written past it, and the don't-care bytes of a
    record's last register word must never reach memory."""
    rng = np.random.default_rng(77 + nbytes)
    n = 128
    records = rng.integers(0, 256, n * nbytes, dtype=np.uint8)
    if nbytes != 6:  # the kernels only ever read half-float positions
        for garbage in (0x00, 0xEE):
            plane = np.full(n * nbytes // 4 + 8, 0xDEADBEEF, np.uint32)
            assert emul.emul_emit_records(C.c_int(nbytes), C.c_int64(n), records.ctypes.data_as(_u8p), plane.ctypes.data_as(_u32p),
                                          C.c_uint8(garbage)) == 0
            assert np.array_equal(plane[:n * nbytes // 4].view(np.uint8), records), (nbytes, garbage)
            assert (plane[n * nbytes // 4:] == 0xDEADBEEF).all(), "wrote past the plane"
    plane = np.concatenate([records, rng.integers(0, 256, 32, dtype=np.uint8)]).view(np.uint32)
    back = np.zeros(n * nbytes, np.uint8)
    assert emul.emul_load_records(C.c_int(nbytes), C.c_int64(n), plane.ctypes.data_as(_u32p), back.ctypes.data_as(_u8p)) == 0
    assert np.array_equal(back, records), nbytes


def test_division_through_the_reciprocal_is_the_ieee_quotient(emul):
    """codec_math.cuh: div_by_rcp (two FMA residual corrections on x * RN(1/b)) against x / b: every
    float in [2^-81, 1.01] and 0 for the constant divisor sqrt1_2 (the magnitudes of the smallest-three
    packer), and 2^27 random pairs from the domain the rotation quantizer's guard admits."""
    bad = np.zeros(2, np.uint32)
    emul.emul_check_div_by_sqrt1_2.restype = C.c_int64
    emul.emul_check_div_by_rcp_random.restype = C.c_int64
    lo, hi = int(np.float32(2.0 ** -81).view(np.uint32)), int(np.float32(1.01).view(np.uint32))
    assert emul.emul_check_div_by_sqrt1_2(C.c_uint32(lo), C.c_uint32(hi), bad.ctypes.data_as(_u32p)) == 0, hex(bad[0])
    assert emul.emul_check_div_by_sqrt1_2(C.c_uint32(0), C.c_uint32(0), bad.ctypes.data_as(_u32p)) == 0
    for seed in (1, 2):
        wrong = emul.emul_check_div_by_rcp_random(C.c_uint64(seed), C.c_int64(1 << 26), bad.ctypes.data_as(_u32p))
        assert wrong == 0, (wrong, hex(bad[0]), hex(bad[1]))
