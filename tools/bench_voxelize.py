"""Throughput of the voxelization path on one B200, next to the reference's own CPU voxelizer —
supplementary to bench.py (whose line is the denoising path, config 2).

Workload = the occupancy caller's geometry (prepare_dataset.py:956-958: 1 mm cells in [-0.2,0.2]^2 x [0,0.4],
max_points 100, max_voxels 1e5) on a synthetic surface-like cloud of N points [N, 4] fp32.
Reports points/s of `points_to_voxels`-style fused voxelization (no dense voxel tensor) and of the reference-shaped
`voxelization` call (dense [1e5, 100, 4] output), the ALGORITHMIC HBM bytes per point and the fraction of the
measured copy bandwidth, and `oracle/_ref` (the reference's voxelization_cpu.cpp) on a bounded sample.

    python tools/bench_voxelize.py [N=2000000]        (under gpurun)
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from orv_b200 import _lib as L  # noqa: E402
from orv_b200.voxelize import hard_voxelize  # noqa: E402


def cloud(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    uv = torch.rand((n, 2), generator=g) * 0.36 - 0.18
    z = 0.2 + 0.05 * torch.sin(uv[:, 0] * 40) * torch.cos(uv[:, 1] * 31) + 0.002 * torch.randn((n,), generator=g)
    lab = torch.randint(1, 13, (n,), generator=g).float()
    return torch.stack([uv[:, 0], uv[:, 1], z, lab], dim=1).contiguous()


def time_cuda(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    vs, cr, mp, mv = [0.001] * 3, [-0.2, -0.2, 0, 0.2, 0.2, 0.4], 100, 100000
    pts_cpu = cloud(n)
    pts = pts_cpu.cuda()
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
        peak_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        peak, peak_src = 6500.0, "fallback (B200_PROFILING.md)"
    # the allocation of outputs / workspace is part of the call a user makes; time it as such
    ms_fused = time_cuda(lambda: hard_voxelize(pts, vs, cr, mp, mv, want_voxels=False, want_labels=True))
    ms_dense = time_cuda(lambda: hard_voxelize(pts, vs, cr, mp, mv, want_voxels=True, want_labels=False))
    out = hard_voxelize(pts, vs, cr, mp, mv, want_voxels=False, want_labels=True)
    m = int(out["voxel_num"].item())
    # algorithmic bytes: every point read once (16 B); fused output = 32 B per voxel; dense = kept points (16 B each,
    # at most max_points per voxel) + 16 B per voxel of coors / counts
    alg_fused = 16.0 * n + 32.0 * m
    alg_dense = 16.0 * n + 16.0 * min(n, m * mp) + 16.0 * m
    res = {
        "workload": f"occupancy voxelization, {n} points [N,4] fp32, 1 mm cells, max_points {mp}, max_voxels {mv}",
        "voxels": m,
        "fused_points_to_voxels": {"ms": ms_fused, "points_per_s": n / ms_fused * 1e3,
                                   "roofline": {"bound": "hbm", "achieved": alg_fused / ms_fused / 1e6, "peak": peak,
                                                "unit": "GB/s", "frac": alg_fused / ms_fused / 1e6 / peak,
                                                "peak_source": peak_src}},
        "reference_shaped_voxelization": {"ms": ms_dense, "points_per_s": n / ms_dense * 1e3,
                                          "roofline": {"bound": "hbm", "achieved": alg_dense / ms_dense / 1e6,
                                                       "peak": peak, "unit": "GB/s",
                                                       "frac": alg_dense / ms_dense / 1e6 / peak}},
        "workspace_mb": L.load().orvb_voxelize_workspace_bytes(n, mv) / 1e6,
    }
    try:
        from oracle import build_ref as R
        ns = min(n, 1_000_000)
        sample = pts_cpu[:ns]
        t0 = time.perf_counter()
        R.voxelization(sample, vs, cr, mp, mv, True)
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": ns / dt, "unit": "points/s", "cores": 1, "kind": "reference",
                               "sample": f"first {ns} points through the reference's voxelization_cpu.cpp (oracle/_ref), "
                                         "dense output allocation included as in voxelization.py:97-103"}
    except Exception as e:  # the prebuilt reference module did not travel
        res["cpu_baseline"] = {"unavailable": str(e)[:200]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
