"""GPU parity of the voxelization kernels (csrc/voxelize.cu) through the C ABI: bit-exact against the reference-run
golden vectors, the numpy oracle and — where the prebuilt module travelled — the reference's own CPU voxelizer."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import build_ref as R
from oracle import voxel_oracle as V

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).parent / "golden"
HARD = ["hard_small", "hard_caps", "hard_occ_1mm", "hard_c5_aniso"]
DYN = ["dyn_small", "dyn_c3"]


def _load(name):
    return torch.load(GOLDEN / f"voxelize_{name}.pt")


def _run_hard(pts, vs, cr, mp, mv):
    from orv_b200.voxelize import voxelization
    v, c, n = voxelization(pts.cuda(), vs, cr, mp, mv, True)
    torch.cuda.synchronize()
    return v.cpu().numpy(), c.cpu().numpy(), n.cpu().numpy()


def _assert_hard_equal(got, want, tag):
    gv, gc, gn = got
    wv, wc, wn = want
    assert gc.shape == wc.shape, f"{tag}: voxel count {gc.shape[0]} != {wc.shape[0]}"
    assert np.array_equal(gc, wc), f"{tag}: coors differ in {(gc != wc).any(axis=1).sum()} of {len(wc)} voxels"
    assert np.array_equal(gn, wn), f"{tag}: num_points_per_voxel differ in {(gn != wn).sum()} voxels"
    assert np.array_equal(gv, wv), f"{tag}: voxel contents differ in {(gv != wv).any(axis=(1, 2)).sum()} voxels"


@pytest.mark.parametrize("name", DYN)
def test_dynamic_golden(name):
    from orv_b200.voxelize import voxelization
    g = _load(name)
    coors = voxelization(g["points"].cuda(), g["voxel_size"], g["coors_range"], -1, -1)
    assert coors.dtype == torch.int32 and np.array_equal(coors.cpu().numpy(), g["coors"].numpy())


@pytest.mark.parametrize("name", HARD)
def test_hard_golden(name):
    g = _load(name)
    got = _run_hard(g["points"], g["voxel_size"], g["coors_range"], g["max_points"], g["max_voxels"])
    assert np.array_equal(got[1], g["coors"].numpy()), f"{name}: coors"
    assert np.array_equal(got[2], g["num_points_per_voxel"].numpy()), f"{name}: num_points_per_voxel"
    if "voxels" in g:
        assert np.array_equal(got[0], g["voxels"].numpy()), f"{name}: voxels"
    else:
        assert np.array_equal(got[0].astype(np.float64).sum(axis=(1, 2)), g["voxels_rowsum"].numpy())


@pytest.mark.parametrize("seed,n,c,mp,mv,cell", [
    (21, 1, 4, 3, 3, 0.05),             # a single point
    (22, 2047, 4, 2, 50, 0.05),         # one tile minus one
    (23, 2049, 3, 5, 100000, 0.01),     # tile + 1, scalar-load path, 12-bit keys (cap = n): two 9-bit sort passes
    (24, 300000, 4, 3, 70000, 0.004),   # many tiles, max_voxels bites, 17-bit keys
    (25, 200000, 5, 100, 255, 0.05),    # one sort pass, long segments, max_points bites
    (26, 100000, 4, 7, 511, 0.05),      # keys 0..511 (sentinel included): 10 bits, two passes
    (27, 100000, 4, 7, 510, 0.05),      # keys 0..510: 9 bits, one pass with every bin in use
])
def test_hard_random_vs_oracle(seed, n, c, mp, mv, cell):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand((n, c), generator=g) * 0.6 - 0.1
    vs, cr = [cell] * 3, [0, 0, 0, 0.4, 0.4, 0.4]
    got = _run_hard(pts, vs, cr, mp, mv)
    _assert_hard_equal(got, V.hard_voxelize(pts.numpy(), vs, cr, mp, mv), f"seed {seed}")


@pytest.mark.skipif(not R.available(), reason="oracle/_ref (reference build) did not travel")
def test_hard_vs_reference_build_and_repeatable():
    g = torch.Generator().manual_seed(31)
    pts = torch.rand((60000, 4), generator=g) * 0.5 - 0.05
    pts[:, 3] = torch.randint(1, 9, (60000,), generator=g).float()
    vs, cr = [0.01] * 3, [0, 0, 0, 0.4, 0.4, 0.4]
    rv, rc, rn = R.voxelization(pts, vs, cr, 6, 20000, True)
    a = _run_hard(pts, vs, cr, 6, 20000)
    _assert_hard_equal(a, (rv.numpy(), rc.numpy(), rn.numpy()), "reference build")
    b = _run_hard(pts, vs, cr, 6, 20000)
    _assert_hard_equal(b, a, "second run")


def test_hard_large_multi_level_scan():
    # 5M points: flag scan and digit-histogram scan both recurse (more than 2048 tiles); checked against the oracle
    n = 5_000_000
    g = torch.Generator().manual_seed(41)
    pts = torch.rand((n, 4), generator=g)
    vs, cr = [0.01] * 3, [0, 0, 0, 1, 1, 1]
    got = _run_hard(pts, vs, cr, 2, 300000)
    _assert_hard_equal(got, V.hard_voxelize(pts.numpy(), vs, cr, 2, 300000), "5M")


def test_points_to_voxels_matches_oracle():
    from orv_b200.voxelize import points_to_voxels
    g = _load("hard_occ_1mm")
    pts = g["points"].numpy()
    labels = pts[:, 3].astype(np.int32) - 1
    out = points_to_voxels(pts[:, :3], voxel_size=[0.001] * 3, labels=labels, point_cloud_range=g["coors_range"])
    want = V.points_to_voxels(pts[:, :3], [0.001] * 3, labels, g["coors_range"])
    assert out.dtype == np.float64 and out.shape == want.shape and np.array_equal(out, want)
    # crowded voxels: more than 100 points each (max_num_points bites, no padding), few labels -> ties
    rng = np.random.default_rng(5)
    p = rng.random((40000, 3), dtype=np.float32) * 0.4
    lab = rng.integers(0, 3, size=(40000,)).astype(np.int32)
    p[:7] = np.nan  # dropped by the caller's NaN filter
    out = points_to_voxels(p, voxel_size=[0.1] * 3, labels=lab, point_cloud_range=[0, 0, 0, 0.4, 0.4, 0.4])
    want = V.points_to_voxels(p, [0.1] * 3, lab, [0, 0, 0, 0.4, 0.4, 0.4])
    assert np.array_equal(out, want)
    # sparse voxels with ties between labels (2-3 points per voxel)
    p = rng.random((3000, 3), dtype=np.float32) * 0.4
    lab = rng.integers(0, 4, size=(3000,)).astype(np.int32)
    out = points_to_voxels(p, voxel_size=[0.04] * 3, labels=lab, point_cloud_range=[0, 0, 0, 0.4, 0.4, 0.4])
    want = V.points_to_voxels(p, [0.04] * 3, lab, [0, 0, 0, 0.4, 0.4, 0.4])
    assert np.array_equal(out, want)


def test_label_vote_beyond_the_staged_slots():
    # max_points = 200 > the 128 labels a warp stages in shared memory: the tail is re-read from global memory
    from orv_b200.voxelize import hard_voxelize
    rng = np.random.default_rng(9)
    p = np.concatenate([rng.random((30000, 3), dtype=np.float32) * 0.4,
                        rng.integers(1, 4, size=(30000, 1)).astype(np.float32)], axis=1)
    vs, cr = [0.1] * 3, [0, 0, 0, 0.4, 0.4, 0.4]
    out = hard_voxelize(torch.from_numpy(p).cuda(), vs, cr, 200, 1000, want_voxels=True, want_labels=True)
    m = int(out["voxel_num"].item())
    vox, coors, _ = V.hard_voxelize(p, vs, cr, 200, 1000)
    assert m == len(coors) and np.array_equal(out["voxels"][:m].cpu().numpy(), vox)
    assert np.array_equal(out["voxel_labels"][:m].cpu().numpy(), V.label_vote(vox, coors))
    assert (out["voxel_labels"][m:] == 0).all()


def test_voxelize_edge_cases_and_errors():
    from orv_b200.voxelize import voxelization
    vs, cr = [0.1] * 3, [0, 0, 0, 1, 1, 1]
    empty = torch.zeros((0, 4), device="cuda")
    assert voxelization(empty, vs, cr, -1, -1).shape == (0, 3)
    v, c, n = voxelization(empty, vs, cr, 4, 10)
    assert v.shape == (0, 4, 4) and c.shape == (0, 3) and n.shape == (0,)
    pts = torch.tensor([[2, 2, 2, 1], [float("nan"), 0.5, 0.5, 1], [float("inf"), 0.5, 0.5, 1], [1.0, 0.5, 0.5, 1],
                        [-1e-9, 0.5, 0.5, 1]], device="cuda")
    assert (voxelization(pts, vs, cr, -1, -1) == -1).all()
    assert voxelization(pts, vs, cr, 4, 10)[1].shape == (0, 3)
    with pytest.raises(RuntimeError):
        voxelization(pts.cpu(), vs, cr, 4, 10)
    with pytest.raises(RuntimeError):
        voxelization(pts.double(), vs, cr, 4, 10)
    with pytest.raises(RuntimeError):
        voxelization(pts, [0.1, 0.0, 0.1], cr, 4, 10)
