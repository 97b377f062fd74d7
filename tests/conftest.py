import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# tests/golden/ holds fixtures -- including a byte-identical copy of the reference's own pytest file, which
# tests/test_reference_suite.py runs in a subprocess against this repo's `spz` module -- not tests of this tree
collect_ignore_glob = ["golden/*"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref).  Present in the build container and, as a shipped
    artefact, on the GPU box; tests that need it skip elsewhere."""
    from oracle import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libspz_ref.so not built and /root/reference absent")
    return Ref()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))


@pytest.fixture(scope="session")
def emul():
    """Host build of the device math header (tests/host_emul/quant_host.cc)."""
    import ctypes
    here = os.path.join(ROOT, "tests", "host_emul")
    out_dir = os.path.join(here, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libquant_host.so")
    src = os.path.join(here, "quant_host.cc")
    hdrs = [os.path.join(ROOT, "spz_b200", "csrc", h) for h in ("codec_math.cuh", "record_align.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in [src] + hdrs):
        # -mfma only turns std::fmaf into the hardware instruction (exact either way, 20x faster); contraction stays off
        fma = ["-mfma"] if "fma" in open("/proc/cpuinfo").read().split("flags", 1)[-1].split("\n", 1)[0].split() else []
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-frounding-math"] + fma +
                       ["-Wno-unknown-pragmas", "-shared", "-fPIC", src, "-o", so], check=True)
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from spz_b200.codec import Context
    ctx = Context(0)
    yield ctx
    ctx.close()
