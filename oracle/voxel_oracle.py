"""TEST INFRASTRUCTURE — CPU restatement of the reference's occupancy voxelization.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this module; the product
(`orv_b200/voxelize.py` -> `csrc/voxelize.cu`) never does.

Pinned: `tests/test_voxel_oracle.py` checks every function here against `oracle/_ref` — the reference's own
`voxelization_cpu.cpp` compiled from /root/reference by `oracle/build_ref.py` — on seeded random clouds, and against
the committed vectors `tests/golden/voxelize_*.pt` that `oracle/make_voxel_golden.py` generated from that build.

Follows
  * `orv/ops/voxelize/voxelization_cpu.cpp:6-46`   dynamic_voxelize_forward_cpu_kernel
  * `orv/ops/voxelize/voxelization_cpu.cpp:48-108` hard_voxelize_forward_cpu_kernel
  * `orv/ops/voxelize/voxelization_cpu.cpp:110-173` grid size = round((max - min) / voxel) in float
  * `orv/dataset/prepare_dataset.py:137-198`        points_to_voxels (label vote on the hard voxels)
"""
from __future__ import annotations

import numpy as np


def grid_size(voxel_size, coors_range) -> np.ndarray:
    """voxelization_cpu.cpp:120-123 — float32 arithmetic, round half away from zero."""
    vs = np.asarray(voxel_size, dtype=np.float32)
    cr = np.asarray(coors_range, dtype=np.float32)
    ext = ((cr[3:] - cr[:3]) / vs).astype(np.float32)
    return np.trunc(ext.astype(np.float64) + np.copysign(0.5, ext)).astype(np.int64)


def dynamic_voxelize(points: np.ndarray, voxel_size, coors_range) -> np.ndarray:
    """coors [N, 3] int32 = (z, y, x), or -1 rows for points outside the range (voxelization_cpu.cpp:18-42)."""
    pts = np.asarray(points, dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    cr = np.asarray(coors_range, dtype=np.float32)
    grid = grid_size(voxel_size, coors_range)
    n = pts.shape[0]
    coors = np.full((n, 3), -1, dtype=np.int32)
    if n == 0:
        return coors
    with np.errstate(invalid="ignore", over="ignore"):
        f = np.floor((pts[:, :3] - cr[None, :3]) / vs[None, :])  # float32 throughout, as `T - float / float`
        ok = np.ones(n, dtype=bool)
        for j in range(3):
            ok &= (f[:, j] >= 0) & (f[:, j] < grid[j])  # NaN fails both (the C cast of NaN is negative on x86)
    c = np.where(ok[:, None], f, 0).astype(np.int64)
    coors[ok, 0] = c[ok, 2]
    coors[ok, 1] = c[ok, 1]
    coors[ok, 2] = c[ok, 0]
    return coors


def hard_voxelize_loop(points: np.ndarray, voxel_size, coors_range, max_points: int, max_voxels: int):
    """Line-by-line restatement of hard_voxelize_forward_cpu_kernel (voxelization_cpu.cpp:70-105); small inputs."""
    pts = np.asarray(points, dtype=np.float32)
    coor = dynamic_voxelize(pts, voxel_size, coors_range)
    n, c = pts.shape
    voxels = np.zeros((max_voxels, max_points, c), dtype=np.float32)
    coors = np.zeros((max_voxels, 3), dtype=np.int32)
    num = np.zeros((max_voxels,), dtype=np.int32)
    coor_to_voxelidx = {}
    voxel_num = 0
    for i in range(n):
        if coor[i, 0] == -1:
            continue
        key = (int(coor[i, 0]), int(coor[i, 1]), int(coor[i, 2]))
        voxelidx = coor_to_voxelidx.get(key, -1)
        if voxelidx == -1:
            voxelidx = voxel_num
            if max_voxels != -1 and voxel_num >= max_voxels:
                continue
            voxel_num += 1
            coor_to_voxelidx[key] = voxelidx
            coors[voxelidx] = coor[i]
        k = num[voxelidx]
        if max_points == -1 or k < max_points:
            voxels[voxelidx, k] = pts[i]
            num[voxelidx] += 1
    return voxels[:voxel_num], coors[:voxel_num], num[:voxel_num]


def hard_voxelize(points: np.ndarray, voxel_size, coors_range, max_points: int, max_voxels: int):
    """Vectorised form of the same semantics: voxels numbered by first appearance, points kept in index order."""
    pts = np.asarray(points, dtype=np.float32)
    n, c = pts.shape
    coor = dynamic_voxelize(pts, voxel_size, coors_range)
    grid = grid_size(voxel_size, coors_range)
    valid = np.nonzero(coor[:, 0] != -1)[0]
    if valid.size == 0:
        return (np.zeros((0, max_points, c), np.float32), np.zeros((0, 3), np.int32), np.zeros((0,), np.int32))
    cz, cy, cx = (coor[valid, k].astype(np.int64) for k in range(3))
    key = (cz * grid[1] + cy) * grid[0] + cx
    _, first, inverse = np.unique(key, return_index=True, return_inverse=True)
    rank_of_unique = np.empty_like(first)
    rank_of_unique[np.argsort(first, kind="stable")] = np.arange(first.size)
    vid = rank_of_unique[inverse]                       # voxel number of every valid point
    keep = vid < max_voxels
    valid, vid = valid[keep], vid[keep]
    m = int(min(first.size, max_voxels))
    order = np.argsort(vid, kind="stable")              # points grouped by voxel, index order inside
    svid, sidx = vid[order], valid[order]
    start = np.searchsorted(svid, np.arange(m), side="left")
    end = np.searchsorted(svid, np.arange(m), side="right")
    pos = np.arange(svid.size) - start[svid]
    voxels = np.zeros((m, max_points, c), dtype=np.float32)
    sel = pos < max_points
    voxels[svid[sel], pos[sel]] = pts[sidx[sel]]
    coors = coor[sidx[start]].astype(np.int32)
    num = np.minimum(end - start, max_points).astype(np.int32)
    return voxels, coors, num


def label_vote(voxels: np.ndarray, coors: np.ndarray) -> np.ndarray:
    """The vote of points_to_voxels (prepare_dataset.py:176-196), with the reference's own torch-CPU ops: unique
    labels over all slots (padding reads 0), per-voxel counts, descending argsort, runner-up when the winner is 0,
    minus 1; result rows (x, y, z, label) in numpy's promoted dtype (float64)."""
    import torch
    v = torch.from_numpy(np.ascontiguousarray(voxels))
    labels = v[..., -1]
    unique_labels, mapped = torch.unique(labels, sorted=True, return_inverse=True)
    counts = torch.zeros((len(v), len(unique_labels))).long()
    counts.scatter_add_(1, mapped.long(), torch.ones_like(mapped).long())
    indices = torch.argsort(counts, dim=-1, descending=True)
    top1 = unique_labels[indices[:, 0]]
    if indices.shape[-1] > 1:
        top2 = unique_labels[indices[:, 1]]
        top1 = torch.where(top1 == 0, top2, top1)
    top1 = top1 - 1
    return np.concatenate([np.asarray(coors)[:, [2, 1, 0]], top1.numpy()[..., np.newaxis]], axis=-1)


def label_vote_plain(voxels: np.ndarray, coors: np.ndarray) -> np.ndarray:
    """The same vote without torch's sort: most frequent value per voxel, ties to the smaller label (what a stable
    descending sort of the counts yields), 0 yields to the runner-up.  Cross-checked against `label_vote`."""
    m = voxels.shape[0]
    out = np.zeros((m, 4), dtype=np.float64)
    for v in range(m):
        vals, cnt = np.unique(voxels[v, :, -1], return_counts=True)
        order = sorted(range(len(vals)), key=lambda j: (-cnt[j], vals[j]))
        top = vals[order[0]]
        if top == 0 and len(order) > 1:
            top = vals[order[1]]
        out[v] = (coors[v, 2], coors[v, 1], coors[v, 0], np.float32(top) - np.float32(1))
    return out


def points_to_voxels(points: np.ndarray, voxel_size, labels, point_cloud_range) -> np.ndarray:
    """prepare_dataset.py:137-198 with its fixed caps (max_voxels 1e5, max_num_points 100)."""
    pts = np.asarray(points, dtype=np.float32)
    lab = np.zeros((pts.shape[0],), np.float32) if labels is None else np.asarray(labels).astype(np.int32).astype(np.float32)
    p4 = np.concatenate([pts[:, :3], lab[:, None]], axis=1).astype(np.float32)
    p4[:, -1] = p4[:, -1] + 1
    p4 = p4[~(np.isnan(p4[:, 0]) | np.isnan(p4[:, 1]) | np.isnan(p4[:, 2]))]
    voxels, coors, _ = hard_voxelize(p4, voxel_size, point_cloud_range, 100, 100000)
    return label_vote(voxels, coors)
