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


# ----------------------------------------------------------------------------------------------------------------
# The reference's 3-D Gaussian rasteriser (orv/ops/diff-gaussian-rasterization: Inria's rasteriser plus a 12-channel
# semantic feature, depth and alpha outputs; the occupancy -> depth / semantic-map renderer of orv/dataset/gs_render.py).
# CUDA-only, so it is cross-compiled here with nvcc for sm_100a (glm is vendored under third_party/) and can only RUN on
# the GPU box, where the GPU tests use it as the oracle and tools/bench_gs_render.py as the timed baseline.
# ----------------------------------------------------------------------------------------------------------------
GS_SRC = REF_OPS / "diff-gaussian-rasterization"
GS_OUT = HERE / "_ref" / "diff_gaussian_rasterization"
GS_NAME = "orv_ref_diff_gaussian_rasterization"
_gs_mod = None


def build_rasterizer(verbose: bool = False):
    from torch.utils.cpp_extension import load
    GS_OUT.mkdir(parents=True, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    srcs = ["cuda_rasterizer/rasterizer_impl.cu", "cuda_rasterizer/forward.cu", "cuda_rasterizer/backward.cu",
            "rasterize_points.cu", "ext.cpp"]  # the reference's own source list (setup.py:22-27)
    return load(GS_NAME, sources=[str(GS_SRC / s) for s in srcs],
                extra_include_paths=[str(GS_SRC / "third_party" / "glm")],
                # rasterizer_impl.h uses uint32_t / std::uintptr_t without <cstdint> (fine with the compilers of its day):
                # pre-include the header instead of touching the reference source
                extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "--pre-include", "cstdint"],
                build_directory=str(GS_OUT), verbose=verbose, with_cuda=True)


def rasterizer_available() -> bool:
    return (GS_OUT / f"{GS_NAME}.so").exists()


def load_rasterizer():
    """The reference extension module `_C` (`rasterize_gaussians`, ext.cpp:14-18) — prebuilt file only (GPU box)."""
    global _gs_mod
    if _gs_mod is not None:
        return _gs_mod
    import torch  # noqa: F401
    so = GS_OUT / f"{GS_NAME}.so"
    if not so.exists():
        raise RuntimeError(f"{so} is missing (python -m oracle.build_ref --rasterizer, in the build container)")
    spec = importlib.util.spec_from_file_location(GS_NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[GS_NAME] = mod
    _gs_mod = mod
    return mod


if __name__ == "__main__":
    if "--rasterizer" in sys.argv:
        m = build_rasterizer(verbose="-v" in sys.argv)
    else:
        m = build(verbose="-v" in sys.argv)
    print(m.__file__)
