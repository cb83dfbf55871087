"""CPU: the voxelization oracle (oracle/voxel_oracle.py) against the reference's own voxelizer.

* golden vectors `tests/golden/voxelize_*.pt` = outputs of the reference's voxelization_cpu.cpp compiled from
  /root/reference (oracle/make_voxel_golden.py);
* live comparison with that build (`oracle/_ref`) on further seeded clouds when it is present (here, and on the GPU
  box, where the prebuilt module travels).
Bit-exact: integer / index work and verbatim float copies.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import build_ref as R
from oracle import voxel_oracle as V

GOLDEN = Path(__file__).parent / "golden"
HARD = ["hard_small", "hard_caps", "hard_occ_1mm", "hard_c5_aniso"]
DYN = ["dyn_small", "dyn_c3"]


def _load(name):
    return torch.load(GOLDEN / f"voxelize_{name}.pt")


@pytest.mark.parametrize("name", DYN)
def test_dynamic_matches_reference_vectors(name):
    g = _load(name)
    coors = V.dynamic_voxelize(g["points"].numpy(), g["voxel_size"], g["coors_range"])
    assert np.array_equal(coors, g["coors"].numpy())
    assert (coors[:, 0] == -1).any() and (coors[:, 0] != -1).any()


@pytest.mark.parametrize("name", HARD)
def test_hard_matches_reference_vectors(name):
    g = _load(name)
    voxels, coors, num = V.hard_voxelize(g["points"].numpy(), g["voxel_size"], g["coors_range"], g["max_points"],
                                         g["max_voxels"])
    assert np.array_equal(coors, g["coors"].numpy())
    assert np.array_equal(num, g["num_points_per_voxel"].numpy())
    if "voxels" in g:
        assert np.array_equal(voxels, g["voxels"].numpy())
    else:
        assert np.array_equal(voxels.astype(np.float64).sum(axis=(1, 2)), g["voxels_rowsum"].numpy())


@pytest.mark.parametrize("name", ["hard_small", "hard_caps"])
def test_loop_and_vectorised_restatements_agree(name):
    g = _load(name)
    a = V.hard_voxelize_loop(g["points"].numpy(), g["voxel_size"], g["coors_range"], g["max_points"], g["max_voxels"])
    b = V.hard_voxelize(g["points"].numpy(), g["voxel_size"], g["coors_range"], g["max_points"], g["max_voxels"])
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_grid_size_rounding():
    # (0.4 - 0) / 0.001 is 400.00003 in float32; 0.3 / 0.1 is 3.0000002; both must round, not truncate up or down
    assert V.grid_size([0.001] * 3, [-0.2, -0.2, 0, 0.2, 0.2, 0.4]).tolist() == [400, 400, 400]
    assert V.grid_size([0.1, 0.25, 0.4], [0, 0, 0, 0.3, 1.0, 1.0]).tolist() == [3, 4, 3]  # 2.5 rounds away from zero


def test_edge_cases():
    vs, cr = [0.1] * 3, [0, 0, 0, 1, 1, 1]
    empty = np.zeros((0, 4), np.float32)
    assert V.dynamic_voxelize(empty, vs, cr).shape == (0, 3)
    v, c, n = V.hard_voxelize(empty, vs, cr, 4, 10)
    assert v.shape == (0, 4, 4) and c.shape == (0, 3) and n.shape == (0,)
    # nothing in range, NaN / inf coordinates, the upper bound itself is outside
    pts = np.array([[2, 2, 2, 1], [np.nan, 0.5, 0.5, 1], [np.inf, 0.5, 0.5, 1], [1.0, 0.5, 0.5, 1], [-1e-9, 0.5, 0.5, 1]],
                   np.float32)
    assert (V.dynamic_voxelize(pts, vs, cr) == -1).all()
    assert V.hard_voxelize(pts, vs, cr, 4, 10)[1].shape == (0, 3)
    # all points in one voxel: max_points caps the count, order = index order
    one = np.tile(np.array([[0.55, 0.55, 0.55, 0]], np.float32), (9, 1))
    one[:, 3] = np.arange(9)
    v, c, n = V.hard_voxelize(one, vs, cr, 4, 10)
    assert c.tolist() == [[5, 5, 5]] and n.tolist() == [4] and v[0, :, 3].tolist() == [0, 1, 2, 3]


@pytest.mark.skipif(not R.available(), reason="oracle/_ref (reference build) not present")
@pytest.mark.parametrize("seed,n,c,mp,mv", [(11, 3000, 4, 5, 300), (12, 2500, 6, 1, 5000), (13, 50000, 4, 100, 100000),
                                            (14, 1, 4, 3, 3), (15, 7000, 3, 2, 64)])
def test_against_reference_build_live(seed, n, c, mp, mv):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand((n, c), generator=g) * 0.6 - 0.1
    vs = [0.02, 0.05, 0.025] if seed != 13 else [0.004] * 3
    cr = [0, 0, 0, 0.4, 0.4, 0.4]
    rv, rc, rn = R.voxelization(pts, vs, cr, mp, mv, True)
    ov, oc, on = V.hard_voxelize(pts.numpy(), vs, cr, mp, mv)
    assert np.array_equal(oc, rc.numpy()) and np.array_equal(on, rn.numpy()) and np.array_equal(ov, rv.numpy())
    rd = R.voxelization(pts, vs, cr, -1, -1)
    assert np.array_equal(V.dynamic_voxelize(pts.numpy(), vs, cr), rd.numpy())


def test_label_vote_ties_and_padding():
    # voxel 0: padding wins (7 zeros) -> runner-up; tie between labels 3 and 5 (one each... ) -> smaller label
    # voxel 1: full, 4 x label 2, 4 x label 9, 2 x label 4 -> tie -> 2
    # voxel 2: 5 x label 6 and 5 empty slots: tie between 0 and 6 -> 0 first -> runner-up 6
    vox = np.zeros((3, 10, 4), np.float32)
    vox[0, :3, 3] = [5, 3, 8]
    vox[1, :, 3] = [2, 9, 2, 9, 4, 9, 2, 4, 9, 2]
    vox[2, :5, 3] = 6
    coors = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], np.int32)
    a = V.label_vote(vox, coors)
    b = V.label_vote_plain(vox, coors)
    assert a.dtype == np.float64 and np.array_equal(a, b)
    assert a[:, 3].tolist() == [2.0, 1.0, 5.0]
    assert a[:, :3].tolist() == [[3, 2, 1], [6, 5, 4], [9, 8, 7]]


def test_label_vote_plain_matches_torch_on_random_voxels():
    g = _load("hard_caps")
    p = g["points"].clone()
    vox, coors, _ = V.hard_voxelize(p.numpy(), [0.1] * 3, g["coors_range"], 12, 64)
    assert np.array_equal(V.label_vote(vox, coors), V.label_vote_plain(vox, coors))


def test_points_to_voxels_oracle_shape():
    g = _load("hard_occ_1mm")
    pts = g["points"].numpy()
    out = V.points_to_voxels(pts[:, :3], [0.001] * 3, pts[:, 3].astype(np.int32) - 1, g["coors_range"])
    assert out.shape == (g["coors"].shape[0], 4) and out.dtype == np.float64
    assert np.array_equal(out[:, :3], g["coors"].numpy()[:, [2, 1, 0]].astype(np.float64))
    assert out[:, 3].min() >= 0
