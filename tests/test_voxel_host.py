"""CPU: host-side behaviour of the voxelization front end (no compute without a GPU)."""
import ctypes as C

import pytest
import torch

from orv_b200 import _lib as L
from orv_b200 import dist as D
from orv_b200 import voxelize as VX


def test_workspace_bytes_is_host_arithmetic():
    lib = L.load()
    sizes = [lib.orvb_voxelize_workspace_bytes(n, 100000) for n in (0, 1, 1000, 2047, 2049, 100000, 2_000_000)]
    assert all(s > 0 and s % 256 == 0 for s in sizes)
    assert sizes == sorted(sizes)
    # ~ 60-100 bytes per point at scale (cell 12 + slot 4 + order 4 + two key/value pairs 16 + hash table 12 x 2..4)
    assert 40 * 2_000_000 < sizes[-1] < 160 * 2_000_000
    assert lib.orvb_voxelize_workspace_bytes(-1, 10) == 0 and lib.orvb_voxelize_workspace_bytes(10, 0) == 0


def test_args_struct_matches_header_layout():
    # field order and types of orvb_voxelize_args (include/orv_b200.h); a mismatch would shift every pointer
    names = [f[0] for f in L.VoxelizeArgs._fields_]
    assert names == ["points", "n", "c", "voxel_size", "coors_range", "max_points", "max_voxels", "voxels", "coors",
                     "num_points_per_voxel", "voxel_num", "voxel_labels", "workspace", "workspace_bytes"]
    # ground truth: gcc's layout of the real header
    import os
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = "#include <stdio.h>\n#include <stddef.h>\n#include \"orv_b200.h\"\nint main(void){printf(\"%zu\", sizeof(orvb_voxelize_args));" \
        + "".join(f'printf(" %zu", offsetof(orvb_voxelize_args, {n}));' for n in names) + "return 0;}\n"
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "layout.c"), os.path.join(td, "layout")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(root, "include"), src, "-o", exe], check=True)
        vals = [int(v) for v in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    assert C.sizeof(L.VoxelizeArgs) == vals[0] == 120
    assert [getattr(L.VoxelizeArgs, n).offset for n in names] == vals[1:]


def test_product_has_no_cpu_path():
    pts = torch.rand(10, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        VX.voxelization(pts, [0.1] * 3, [0, 0, 0, 1, 1, 1], 4, 10)
    with pytest.raises(RuntimeError, match="no CPU path"):
        VX.voxelization(pts, [0.1] * 3, [0, 0, 0, 1, 1, 1], -1, -1)
    with pytest.raises(ValueError):
        VX._geometry([0.1, 0.1], [0, 0, 0, 1, 1, 1])


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-box behaviour")
def test_c_abi_fails_loudly_without_a_device():
    lib = L.load()
    a = L.VoxelizeArgs()
    a.n, a.c, a.max_points, a.max_voxels = 0, 4, 4, 10
    rc = lib.orvb_hard_voxelize(C.byref(a), None)
    assert rc != 0 and lib.orvb_last_error()
    vs = (L.c_float * 3)(0.1, 0.1, 0.1)
    cr = (L.c_float * 6)(0, 0, 0, 1, 1, 1)
    assert lib.orvb_dynamic_voxelize(None, 0, 4, vs, cr, None, None) != 0


def test_frames_shard_like_clips():
    # occupancy frames are independent files: rank r voxelizes frames [r n / N, (r+1) n / N) (remainder to the last
    # rank), the rule the reference uses for clips (evaluation_control_to_video.py:212-222)
    n, world = 37, 8
    spans = [D.shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
