"""TEST INFRASTRUCTURE (oracle/): builds and loads `oracle/_ref` — the reference's OWN CPU voxelizer.

`/root/reference/orv/ops/voxelize/voxelization_cpu.cpp` (+ the headers under `/root/reference/orv/ops/include`) is
compiled where it lies with g++ through `torch.utils.cpp_extension.load` — the reference's own recipe
(`voxelization.py:27-38`), not its build system — together with `oracle/ref_voxelization_glue.cpp` (the two
dispatcher definitions the reference's CPU-only build forgets to link).  Outputs go to
`oracle/_ref/voxelization_cpu/` only (git-ignored, not gpurun-ignored: the prebuilt module travels to the GPU box,
where `/root/reference` does not exist and `load()` just imports the prebuilt file).

Only `tests/`, `__graft_entry__` and `bench.py`'s CPU-baseline legs may import this module.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_OPS = Path("/root/reference/orv/ops")
OUT = HERE / "_ref" / "voxelization_cpu"
NAME = "orv_ref_voxelization_cpu"

_mod = None


def available() -> bool:
    """True when the reference module can actually be loaded (prebuilt file or sources to build it from)."""
    if not ((OUT / f"{NAME}.so").exists() or REF_OPS.exists()):
        return False
    try:
        load()
        return True
    except Exception:  # noqa: BLE001  (a prebuilt module from another torch build, a missing compiler, ...)
        return False


def build(verbose: bool = False):
    """Compiles the reference sources (needs /root/reference; a no-op rebuild when up to date)."""
    from torch.utils.cpp_extension import load
    OUT.mkdir(parents=True, exist_ok=True)
    return load(NAME,
                sources=[str(REF_OPS / "voxelize" / "voxelization_cpu.cpp"), str(HERE / "ref_voxelization_glue.cpp")],
                extra_include_paths=[str(REF_OPS / "include")],
                build_directory=str(OUT), verbose=verbose)


def load():
    """The reference module (`hard_voxelize_forward`, `dynamic_voxelize_forward`; voxelization_cpu.cpp:230-241)."""
    global _mod
    if _mod is not None:
        return _mod
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    so = OUT / f"{NAME}.so"
    if REF_OPS.exists() and not os.environ.get("ORVB_NO_BUILD"):
        _mod = build()
        return _mod
    if not so.exists():
        raise RuntimeError(f"{so} is missing and /root/reference is not present to build it")
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[NAME] = mod
    _mod = mod
    return mod


def voxelization(points, voxel_size, coors_range, max_points: int = 35, max_voxels: int = 20000,
                 deterministic: bool = True):
    """The reference's `_Voxelization.forward` call sequence (voxelization.py:87-119) on the CPU module: same
    allocations, same arguments, same slicing by `voxel_num`."""
    import torch
    op = load()
    if max_points == -1 or max_voxels == -1:
        coors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
        op.dynamic_voxelize_forward(points, torch.tensor(voxel_size, dtype=torch.float),
                                    torch.tensor(coors_range, dtype=torch.float), coors, 3)
        return coors
    voxels = points.new_zeros(size=(max_voxels, max_points, points.size(1)))
    coors = points.new_zeros(size=(max_voxels, 3), dtype=torch.int)
    num_points_per_voxel = points.new_zeros(size=(max_voxels,), dtype=torch.int)
    voxel_num = torch.zeros(size=(), dtype=torch.long)
    op.hard_voxelize_forward(points, torch.tensor(voxel_size, dtype=torch.float),
                             torch.tensor(coors_range, dtype=torch.float), voxels, coors, num_points_per_voxel,
                             voxel_num, max_points, max_voxels, 3, deterministic)
    return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


if __name__ == "__main__":
    m = build(verbose="-v" in sys.argv)
    print(m.__file__)
