"""TEST INFRASTRUCTURE ONLY -- ctypes front-ends for the two CPU checkers.

* ``Oracle``  : oracle/_build/libspz_oracle.so, the plain-C restatement (spz_oracle.c).
* ``Ref``     : oracle/_ref/libspz_ref.so, the unmodified reference C++ compiled in place from
                /root/reference/src/cc (ref_shim.cc marshals flat arrays).  Optional: present
                wherever ``make -C oracle ref`` has been run (this container; it ships to the GPU
                box as a built artefact because oracle/_ref/ is git-ignored but not gpurun-ignored).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  The product (spz_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libspz_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libspz_ref.so")
REFERENCE_ROOT = "/root/reference"

SH_DIM = {0: 0, 1: 3, 2: 8, 3: 15}

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)


def build(ref: bool = True) -> None:
    """Compile the checkers (idempotent).  The reference build is attempted only where
    /root/reference exists; elsewhere the prebuilt oracle/_ref/libspz_ref.so is used as shipped."""
    targets = ["all"]
    if ref and os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "cc")):
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def _bp(a: np.ndarray):
    return a.ctypes.data_as(_u8p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32).reshape(-1)


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)


@dataclass
class Cloud:
    """Flat float planes, the reference's GaussianCloud layout (splat-types.h:90-115)."""
    n: int
    sh_degree: int
    positions: np.ndarray
    scales: np.ndarray
    rotations: np.ndarray  # x, y, z, w
    alphas: np.ndarray
    colors: np.ndarray
    sh: np.ndarray
    antialiased: bool = False

    def planes(self):
        return (self.positions, self.scales, self.rotations, self.alphas, self.colors, self.sh)

    def slice(self, a: int, b: int) -> "Cloud":
        d = SH_DIM[self.sh_degree] * 3
        return Cloud(b - a, self.sh_degree, self.positions[3 * a:3 * b], self.scales[3 * a:3 * b],
                     self.rotations[4 * a:4 * b], self.alphas[a:b], self.colors[3 * a:3 * b],
                     self.sh[d * a:d * b], self.antialiased)


@dataclass
class Packed:
    """Byte planes, the reference's PackedGaussians layout (load-spz.h:42-59)."""
    n: int
    sh_degree: int
    fractional_bits: int
    version: int  # 1, 2 or 3 (load-spz.cc:571-572)
    positions: np.ndarray
    scales: np.ndarray
    rotations: np.ndarray
    alphas: np.ndarray
    colors: np.ndarray
    sh: np.ndarray
    antialiased: bool = False

    def planes(self):
        return (self.positions, self.scales, self.rotations, self.alphas, self.colors, self.sh)

    def slice(self, a: int, b: int) -> "Packed":
        d = SH_DIM[self.sh_degree] * 3
        pb = 6 if self.version in (1, 4) else 9
        rb = 4 if self.version >= 3 else 3
        return Packed(b - a, self.sh_degree, self.fractional_bits, self.version,
                      self.positions[pb * a:pb * b], self.scales[3 * a:3 * b],
                      self.rotations[rb * a:rb * b], self.alphas[a:b], self.colors[3 * a:3 * b],
                      self.sh[d * a:d * b], self.antialiased)


def _empty_packed(n: int, deg: int, version: int = 3, fb: int = 12) -> Packed:
    d = SH_DIM[deg] * 3
    return Packed(n, deg, fb, version,
                  np.zeros(n * (6 if version in (1, 4) else 9), np.uint8), np.zeros(n * 3, np.uint8),
                  np.zeros(n * (4 if version >= 3 else 3), np.uint8), np.zeros(n, np.uint8),
                  np.zeros(n * 3, np.uint8), np.zeros(n * d, np.uint8))


def _empty_cloud(n: int, deg: int) -> Cloud:
    d = SH_DIM[deg] * 3
    z = lambda k: np.zeros(k, np.float32)  # noqa: E731
    return Cloud(n, deg, z(n * 3), z(n * 3), z(n * 4), z(n), z(n * 3), z(n * d))


class _Codec:
    """Shared marshalling for the two libraries (same flat signatures modulo the timing arg)."""

    _prefix = ""
    _timed = False

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.path = path

    def pack(self, c: Cloud, frm: int = 0) -> Packed:
        out = _empty_packed(c.n, c.sh_degree)
        out.antialiased = c.antialiased
        ins = [_f32(p) for p in c.planes()]
        fn = getattr(self.lib, self._prefix + "pack")
        fn.restype = C.c_int
        args = [C.c_int64(c.n) if not self._timed else C.c_int32(c.n), C.c_int32(c.sh_degree),
                C.c_int32(frm)] + [_fp(a) for a in ins] + [_bp(a) for a in out.planes()]
        self.last_seconds = None
        if self._timed:
            sec = C.c_double(0)
            args.append(C.byref(sec))
        rc = fn(*args)
        if self._timed:
            self.last_seconds = sec.value
        if rc != 0:
            raise ValueError(f"{self._prefix}pack rejected the input (rc={rc})")
        return out

    def unpack(self, p: Packed, to: int = 0) -> Cloud:
        out = _empty_cloud(p.n, p.sh_degree)
        out.antialiased = p.antialiased
        ins = [_u8(a) for a in p.planes()]
        fn = getattr(self.lib, self._prefix + "unpack")
        fn.restype = C.c_int
        args = [C.c_int64(p.n) if not self._timed else C.c_int32(p.n), C.c_int32(p.sh_degree),
                C.c_int32(p.fractional_bits), C.c_int32(p.version), C.c_int32(to)] + \
               [_bp(a) for a in ins] + [_fp(a) for a in out.planes()]
        self.last_seconds = None
        if self._timed:
            sec = C.c_double(0)
            args.append(C.byref(sec))
        rc = fn(*args)
        if self._timed:
            self.last_seconds = sec.value
        if rc != 0:
            raise ValueError(f"{self._prefix}unpack rejected the input (rc={rc})")
        return out


class Oracle(_Codec):
    _prefix = "oracle_"
    _timed = False

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        super().__init__(ORACLE_SO)
        L = self.lib
        L.oracle_flips.argtypes = [C.c_int32, C.c_int32, _f32p, _f32p, _f32p]
        L.oracle_sweep_u8.argtypes = [C.c_int32, C.c_uint32, C.c_uint32, C.c_int64, _u8p]
        L.oracle_fnv1a64.argtypes = [_u8p, C.c_int64]
        L.oracle_fnv1a64.restype = C.c_uint64
        for name, rt, at in (("oracle_dequant_scale", C.c_float, C.c_uint8),
                             ("oracle_dequant_alpha", C.c_float, C.c_uint8),
                             ("oracle_dequant_color", C.c_float, C.c_uint8),
                             ("oracle_dequant_sh", C.c_float, C.c_uint8),
                             ("oracle_half_to_float", C.c_float, C.c_uint16),
                             ("oracle_quant_alpha", C.c_uint8, C.c_float),
                             ("oracle_quant_scale", C.c_uint8, C.c_float),
                             ("oracle_quant_color", C.c_uint8, C.c_float)):
            f = getattr(L, name)
            f.restype, f.argtypes = rt, [at]

    def pack_v2(self, c: Cloud, frm: int = 0) -> Packed:
        """PARITY UNPINNED restatement of upstream's version-2 encoder (see oracle_pack_v2)."""
        out = _empty_packed(c.n, c.sh_degree, version=2)
        ins = [_f32(p) for p in c.planes()]
        fn = self.lib.oracle_pack_v2
        fn.restype = C.c_int
        rc = fn(C.c_int64(c.n), C.c_int32(c.sh_degree), C.c_int32(frm), *[_fp(a) for a in ins], *[_bp(a) for a in out.planes()])
        if rc != 0:
            raise ValueError(f"oracle_pack_v2 rejected the input (rc={rc})")
        return out

    def unpack_at(self, p: Packed, indices, conv21) -> np.ndarray:
        """[len(indices), 59] float32: PackedGaussians::unpack(i, c) per index (load-spz.cc:383-463)."""
        idx = np.ascontiguousarray(indices, np.int64)
        conv = np.ascontiguousarray(conv21, np.float32)
        assert conv.size == 21
        out = np.zeros((idx.size, 59), np.float32)
        ins = [_u8(a) for a in p.planes()]
        fn = self.lib.oracle_unpack_at
        fn.restype = C.c_int
        rc = fn(C.c_int64(p.n), C.c_int32(p.sh_degree), C.c_int32(p.fractional_bits), C.c_int32(p.version),
                *[_bp(a) for a in ins], idx.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(idx.size), _fp(conv), _fp(out))
        if rc != 0:
            raise ValueError(f"oracle_unpack_at rejected the input (rc={rc})")
        return out

    def converter(self, frm: int, to: int) -> np.ndarray:
        return np.concatenate(self.flips(frm, to)).astype(np.float32)

    def flips(self, frm: int, to: int):
        p, q, s = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(15, np.float32)
        self.lib.oracle_flips(frm, to, _fp(p), _fp(q), _fp(s))
        return p, q, s

    def sweep_u8(self, which: int, first_bits: int, stride: int, count: int) -> np.ndarray:
        out = np.zeros(count, np.uint8)
        self.lib.oracle_sweep_u8(which, first_bits, stride, count, _bp(out))
        return out

    def fnv1a64(self, a: np.ndarray) -> int:
        b = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        return int(self.lib.oracle_fnv1a64(_bp(b), b.size))

    def dequant_table(self, which: str) -> np.ndarray:
        f = getattr(self.lib, "oracle_dequant_" + which)
        return np.array([f(i) for i in range(256)], dtype=np.float32)


class Ref(_Codec):
    """The real reference.  Limits are the reference's own: int32 sizes, <=10M points to load."""
    _prefix = "ref_"
    _timed = True

    def __init__(self):
        if not os.path.exists(REF_SO):
            build(ref=True)
        super().__init__(REF_SO)
        L = self.lib
        L.ref_save_spz.restype = C.c_void_p
        L.ref_serialize.restype = C.c_void_p
        L.ref_load_spz.restype = C.c_void_p
        L.ref_gzip_size.restype = C.c_uint64

    def unpack_at(self, p: Packed, indices, conv21) -> np.ndarray:
        """[len(indices), 59] float32 from the reference's own PackedGaussians::unpack(i, c)."""
        idx = np.ascontiguousarray(indices, np.int64)
        conv = np.ascontiguousarray(conv21, np.float32)
        out = np.zeros((idx.size, 59), np.float32)
        ins = [_u8(a) for a in p.planes()]
        fn = self.lib.ref_unpack_at
        fn.restype = C.c_int
        rc = fn(C.c_int32(p.n), C.c_int32(p.sh_degree), C.c_int32(p.fractional_bits), C.c_int32(p.version),
                *[_bp(a) for a in ins], idx.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(idx.size), _fp(conv), _fp(out))
        if rc != 0:
            raise ValueError(f"ref_unpack_at rejected the input (rc={rc})")
        return out

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO) or os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "cc"))

    def _blob(self, fname: str, c: Cloud, frm: int) -> bytes:
        ins = [_f32(p) for p in c.planes()]
        size = C.c_uint64(0)
        fn = getattr(self.lib, fname)
        ptr = fn(C.c_int32(c.n), C.c_int32(c.sh_degree), C.c_int32(int(c.antialiased)),
                 C.c_int32(frm), *[_fp(a) for a in ins], C.byref(size))
        if not ptr:
            raise ValueError(fname + " failed")
        try:
            return C.string_at(ptr, size.value)
        finally:
            self.lib.ref_free(C.c_void_p(ptr))

    def save_spz(self, c: Cloud, frm: int = 0) -> bytes:
        return self._blob("ref_save_spz", c, frm)

    def serialize(self, c: Cloud, frm: int = 0) -> bytes:
        return self._blob("ref_serialize", c, frm)

    def load_spz(self, blob: bytes, to: int = 0) -> Cloud:
        buf = np.frombuffer(blob, dtype=np.uint8)
        h = self.lib.ref_load_spz(_bp(buf), C.c_int32(buf.size), C.c_int32(to))
        try:
            n, deg, aa = C.c_int32(0), C.c_int32(0), C.c_int32(0)
            self.lib.ref_cloud_info(C.c_void_p(h), C.byref(n), C.byref(deg), C.byref(aa))
            out = _empty_cloud(n.value, deg.value)
            out.antialiased = bool(aa.value)
            self.lib.ref_cloud_copy(C.c_void_p(h), *[_fp(a) for a in out.planes()])
            return out
        finally:
            self.lib.ref_cloud_free(C.c_void_p(h))

    def gzip_size(self, data: bytes):
        buf = np.frombuffer(data, dtype=np.uint8)
        sec = C.c_double(0)
        n = self.lib.ref_gzip_size(_bp(buf), C.c_uint64(buf.size), C.byref(sec))
        return int(n), sec.value


def bits(a: np.ndarray) -> np.ndarray:
    """float32 array -> uint32 bit patterns, with every NaN collapsed to one canonical value
    (the parity rule compares NaNs as a class: x86 and sm_100a mint different payloads)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b
