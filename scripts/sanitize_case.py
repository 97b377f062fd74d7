"""Small end-to-end case (written for compute-sanitizer, which this GPU pool keeps closed; now run
by tests/test_gpu_parity.py::test_alternate_launch_shapes under each development knob): every kernel variant once (tile + generic, all SH
degrees, every stream flavour, host pipeline with and without bounce buffers), checked against the
oracle so the run also fails on wrong results."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import oracle as O
from spz_b200 import codec
from util import random_cloud, random_stream, assert_packed_equal, assert_cloud_bits_equal, cloud_to_ply_rows

# enough tiles that persistent CTAs (148 x SPZB200_CTAS_PER_SM of them) each walk several
TILES = 310 if os.environ.get("SPZB200_GRID") == "persistent" else 7
chk = O.Oracle()
rng = np.random.default_rng(9)
with codec.Context(0) as ctx:
    for deg in range(4):
        n = TILES * codec.tile_gaussians(deg) + 37
        c = random_cloud(rng, n, deg, True)
        dev = codec.CloudPlanes(n, deg, *[torch.from_numpy(p).cuda() for p in c.planes()])
        p = ctx.encode_device(dev, 6); torch.cuda.synchronize()
        assert_packed_equal(O.Packed(n, deg, 12, 3, *[a.cpu().numpy() for a in p.planes()]), chk.pack(c, 6), f"enc deg{deg}")
        for ver in (1, 2, 3, 4):
            s = random_stream(rng, n, deg, ver, 12)
            sd = codec.PackedPlanes(n, deg, *[torch.from_numpy(a).cuda() for a in s.planes()], fractional_bits=12, version=ver)
            g = ctx.decode_device(sd, 7); torch.cuda.synchronize()
            assert_cloud_bits_equal(O.Cloud(n, deg, *[a.cpu().numpy() for a in g.planes()]), chk.unpack(s, 7), f"dec deg{deg} v{ver}")
    # fused PLY-rows kernels, canonical property order (compile-time columns unless SPZB200_PLY=mapped)
    for deg in range(4):
        names = codec.ply_property_names(deg)
        n = TILES * 128 + 37
        c = random_cloud(rng, n, deg, True)
        rows = torch.from_numpy(cloud_to_ply_rows(c, names).reshape(-1)).cuda()
        p = ctx.encode_ply_device(rows, n, names, deg, 6); torch.cuda.synchronize()
        assert_packed_equal(O.Packed(n, deg, 12, 3, *[a.cpu().numpy() for a in p.planes()]), chk.pack(c, 6), f"ply enc deg{deg}")
        for ver in (1, 2, 3, 4):
            s = random_stream(rng, n, deg, ver, 12)
            sd = codec.PackedPlanes(n, deg, *[torch.from_numpy(a).cuda() for a in s.planes()], fractional_bits=12, version=ver)
            got = ctx.decode_ply_device(sd, names, 7); torch.cuda.synchronize()
            want = cloud_to_ply_rows(chk.unpack(s, 7), names).reshape(-1)
            assert np.array_equal(O.bits(got.cpu().numpy()), O.bits(want)), f"ply dec deg{deg} v{ver}"
    n, deg = 5 * codec.tile_gaussians(3) + 11, 3
    c = random_cloud(rng, n, deg, False)
    ctx.set_chunk_points(2 * codec.tile_gaussians(3))
    for mode in (0, 2):
        ctx.set_host_staging(mode, 3)
        got, _ = ctx.encode_host(codec.CloudPlanes(n, deg, *c.planes()), 0)
        assert_packed_equal(O.Packed(n, deg, 12, 3, *got.planes()), chk.pack(c, 0), f"host mode {mode}")
        back, _ = ctx.decode_host(got, 8)
        assert_cloud_bits_equal(O.Cloud(n, deg, *back.planes()), chk.unpack(chk.pack(c, 0), 8), f"host dec mode {mode}")
print("sanitize_case ok")
