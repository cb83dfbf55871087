"""TEST INFRASTRUCTURE — writes tests/golden/voxelize_*.pt from the reference's OWN voxelizer.

Runs `oracle/_ref` (the reference's voxelization_cpu.cpp compiled where it lies, see oracle/build_ref.py) through the
reference's call sequence (`_Voxelization.forward`, voxelization.py:87-119) on seeded clouds and stores inputs and
outputs.  Needs /root/reference (this container); the vectors travel, the reference does not.

    python -m oracle.make_voxel_golden
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from . import build_ref as R

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def cloud(seed: int, n: int, c: int, lo: float, hi: float, label_max: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    p = torch.rand((n, c), generator=g) * (hi - lo) + lo
    if label_max > 0:
        p[:, -1] = torch.randint(1, label_max + 1, (n,), generator=g).float()
    return p


def cases():
    # name, points, voxel_size, range, max_points, max_voxels
    yield "hard_small", cloud(0, 600, 4, -0.05, 0.45, 5), [0.05] * 3, [0, 0, 0, 0.4, 0.4, 0.4], 3, 100
    # both caps bite: 8^3 = 512 cells, 4000 points, at most 2 points in at most 200 voxels
    yield "hard_caps", cloud(1, 4000, 4, -0.02, 0.42, 7), [0.05] * 3, [0, 0, 0, 0.4, 0.4, 0.4], 2, 200
    # the occupancy caller's geometry (prepare_dataset.py:956-958): 1 mm cells in [-0.2, 0.2]^2 x [0, 0.4]
    p = cloud(2, 5000, 4, 0.0, 1.0, 12)
    p[:, :3] = p[:, :3] * torch.tensor([0.05, 0.05, 0.05]) + torch.tensor([-0.01, 0.0, 0.2])
    yield "hard_occ_1mm", p, [0.001] * 3, [-0.2, -0.2, 0, 0.2, 0.2, 0.4], 100, 100000
    # 5 features, anisotropic cells, off-origin range
    yield "hard_c5_aniso", cloud(3, 3000, 5, -1.2, 1.3), [0.25, 0.1, 0.5], [-1, -1, -1, 1, 1, 1], 8, 1000
    # dynamic mode
    yield "dyn_small", cloud(4, 2000, 4, -0.1, 0.5), [0.01] * 3, [0, 0, 0, 0.4, 0.4, 0.4], -1, -1
    yield "dyn_c3", cloud(5, 1500, 3, -2.0, 2.0), [0.3, 0.2, 0.1], [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], -1, -1


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    for name, pts, vs, cr, mp, mv in cases():
        res = R.voxelization(pts, vs, cr, mp, mv, True)
        item = {"points": pts, "voxel_size": vs, "coors_range": cr, "max_points": mp, "max_voxels": mv}
        if isinstance(res, tuple):
            item.update(voxels=res[0].clone(), coors=res[1].clone(), num_points_per_voxel=res[2].clone())
            # keep the files small: voxels are re-derivable from (points, kept indices); store a float64 checksum per
            # voxel instead of the dense tensor when it is large
            if item["voxels"].numel() > 200_000:
                item["voxels_rowsum"] = item.pop("voxels").double().sum(dim=(1, 2))
        else:
            item["coors"] = res.clone()
        torch.save(item, OUT / f"voxelize_{name}.pt")
        shape = {k: tuple(v.shape) for k, v in item.items() if isinstance(v, torch.Tensor)}
        print(name, shape)


if __name__ == "__main__":
    main()
