"""One fused (label vote) and one dense voxelization of a 2M-point cloud, for an ncu launch list:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/<tag>_voxel_launches.csv python tools/profile_voxelize.py"""
import sys

import torch

sys.path.insert(0, ".")
from tools.bench_voxelize import cloud  # noqa: E402
from orv_b200.voxelize import hard_voxelize  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
pts = cloud(n).cuda()
vs, cr = [0.001] * 3, [-0.2, -0.2, 0, 0.2, 0.2, 0.4]
out = hard_voxelize(pts, vs, cr, 100, 100000, want_voxels=False, want_labels=True)
torch.cuda.synchronize()
out = hard_voxelize(pts, vs, cr, 100, 100000, want_voxels=True, want_labels=False)
torch.cuda.synchronize()
print("voxels", int(out["voxel_num"].item()))
