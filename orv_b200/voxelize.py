"""Occupancy voxelization on B200 — the same call surface as the reference extension and its caller.

Mirrors
  * `orv/ops/voxelize/voxelization.py:41-122` — `voxelization(points, voxel_size, coors_range, max_points,
    max_voxels, deterministic)` (dynamic mode when `max_points == -1 or max_voxels == -1`, hard mode otherwise), and
  * `orv/dataset/prepare_dataset.py:137-198` — `points_to_voxels(...)`, the occupancy caller, whose label vote runs
    here in one fused kernel instead of on a `[1e5, 100, 4]` tensor copied to the host.

The arithmetic is in `csrc/voxelize.cu` behind `orvb_dynamic_voxelize` / `orvb_hard_voxelize` (include/orv_b200.h).
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib as L

__all__ = ["voxelization", "points_to_voxels", "hard_voxelize"]


def _geometry(voxel_size, coors_range) -> Tuple[torch.Tensor, torch.Tensor]:
    # the reference builds float32 tensors from whatever the caller passes (voxelization.py:91-92, 105-106)
    vs = torch.tensor([float(v) for v in _as_list(voxel_size, 3)], dtype=torch.float)
    cr = torch.tensor([float(v) for v in _as_list(coors_range, 6)], dtype=torch.float)
    return vs, cr


def _as_list(x, n: int) -> List[float]:
    if isinstance(x, (int, float)):
        return [float(x)] * n
    x = list(x)
    if len(x) != n:
        raise ValueError(f"expected {n} values, got {len(x)}")
    return x


def _check_points(points: torch.Tensor) -> torch.Tensor:
    if not isinstance(points, torch.Tensor) or not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor (liborv_b200 has no CPU path)")
    if points.dim() != 2 or points.size(1) < 3:
        raise RuntimeError(f"points must be [N, >=3], got {tuple(points.shape)}")
    if points.dtype != torch.float32:
        raise RuntimeError(f"points must be float32 (what the occupancy caller passes), got {points.dtype}")
    return points.contiguous()


def hard_voxelize(points: torch.Tensor, voxel_size, coors_range, max_points: int, max_voxels: int, *,
                  want_voxels: bool = True, want_labels: bool = False):
    """One `orvb_hard_voxelize` launch sequence.  Returns a dict of full-size device buffers plus `voxel_num`
    (device int64 scalar); nothing is synchronised here."""
    points = _check_points(points)
    lib = L.load()
    n, c = points.shape
    dev = points.device
    vs, cr = _geometry(voxel_size, coors_range)
    a = L.VoxelizeArgs()
    a.points, a.n, a.c = points.data_ptr(), n, c
    for i in range(3):
        a.voxel_size[i] = float(vs[i])
    for i in range(6):
        a.coors_range[i] = float(cr[i])
    a.max_points, a.max_voxels = int(max_points), int(max_voxels)
    out = {}
    if want_voxels:
        out["voxels"] = points.new_zeros((max_voxels, max_points, c))
        a.voxels = out["voxels"].data_ptr()
    out["coors"] = torch.zeros((max_voxels, 3), dtype=torch.int32, device=dev)
    out["num_points_per_voxel"] = torch.zeros((max_voxels,), dtype=torch.int32, device=dev)
    out["voxel_num"] = torch.zeros((), dtype=torch.int64, device=dev)
    a.coors, a.num_points_per_voxel = out["coors"].data_ptr(), out["num_points_per_voxel"].data_ptr()
    a.voxel_num = out["voxel_num"].data_ptr()
    if want_labels:
        out["voxel_labels"] = torch.zeros((max_voxels, 4), dtype=torch.float64, device=dev)
        a.voxel_labels = out["voxel_labels"].data_ptr()
    ws_bytes = lib.orvb_voxelize_workspace_bytes(n, int(max_voxels))
    ws = torch.empty((max(ws_bytes, 256),), dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
    L.check(lib.orvb_hard_voxelize(a, L.current_stream()), "orvb_hard_voxelize")
    out["_workspace"] = ws  # keeps the scratch alive until the caller has consumed the (async) results
    return out


def voxelization(points: torch.Tensor, voxel_size: Union[Sequence[float], float],
                 coors_range: Union[Sequence[float], float], max_points: int = 35, max_voxels: int = 20000,
                 deterministic: bool = True):
    """Drop-in for the reference's `voxelization` (`_Voxelization.apply`, voxelization.py:122).

    Dynamic mode (`max_points == -1 or max_voxels == -1`) returns `coors [N, 3]` int32 (z, y, x), rows of
    out-of-range points are -1.  Hard mode returns `(voxels [M, max_points, C], coors [M, 3], num_points_per_voxel
    [M])` sliced to the `M` voxels produced.  `deterministic=False` returns the deterministic result too (a valid
    outcome of the reference's atomics-ordered variant)."""
    points = _check_points(points)
    if max_points == -1 or max_voxels == -1:
        lib = L.load()
        vs, cr = _geometry(voxel_size, coors_range)
        coors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
        vs_c = (L.c_float * 3)(*[float(v) for v in vs])
        cr_c = (L.c_float * 6)(*[float(v) for v in cr])
        L.check(lib.orvb_dynamic_voxelize(points.data_ptr(), points.size(0), points.size(1), vs_c, cr_c,
                                          coors.data_ptr(), L.current_stream()), "orvb_dynamic_voxelize")
        return coors
    out = hard_voxelize(points, voxel_size, coors_range, int(max_points), int(max_voxels))
    m = int(out["voxel_num"].item())  # the reference slices by a host-side voxel_num as well (voxelization.py:116-119)
    return out["voxels"][:m], out["coors"][:m], out["num_points_per_voxel"][:m]


def points_to_voxels(points, voxel_size: list = [0.2, 0.2, 0.2], labels=None, max_num_points: int = -1,
                     point_cloud_range=None, device: torch.device = torch.device("cuda"),
                     determinstic: bool = True) -> np.ndarray:
    """Drop-in for `points_to_voxels` (prepare_dataset.py:137-198): `[M, 4]` float64 rows `(x, y, z, label)` with
    the most frequent label of each occupied voxel.  As in the reference, `max_num_points` is overridden by 100 and
    `max_voxels` is 1e5 (:162-163); the spelling of `determinstic` is the reference's."""
    if isinstance(points, np.ndarray):
        points = torch.tensor(points, device=device, dtype=torch.float32)
    if labels is None:
        labels = torch.zeros_like(points[:, 0])
    if isinstance(labels, np.ndarray):
        labels = torch.tensor(labels.astype(np.int32), device=points.device, dtype=torch.float32)
    points = torch.cat([points[:, :3], labels[..., None].float()], dim=1)
    points[:, -1] = points[:, -1] + 1  # empty voxel slots read as 0
    max_voxels = int(1e5)
    max_num_points = int(1e2)
    points = points[~(torch.isnan(points[:, 0]) | torch.isnan(points[:, 1]) | torch.isnan(points[:, 2]))]
    if point_cloud_range is None:
        point_cloud_range = [points[:, 0].min(), points[:, 1].min(), points[:, 2].min(),
                             points[:, 0].max(), points[:, 1].max(), points[:, 2].max()]
    out = hard_voxelize(points.float(), voxel_size, point_cloud_range, max_num_points, max_voxels, want_voxels=False,
                        want_labels=True)
    m = int(out["voxel_num"].item())
    return out["voxel_labels"][:m].cpu().numpy()
