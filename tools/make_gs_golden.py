"""Writes gpurun_out/gs_render_small.pt: the REFERENCE rasteriser's own output (oracle/_ref build of
orv/ops/diff-gaussian-rasterization, run on the GPU box) for the seeded small scene of oracle/gs_oracle.py.  The file is
then copied to tests/golden/ and committed: it pins the numpy oracle in the CPU-only suite.
    python tools/make_gs_golden.py        (under gpurun)"""
import os
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from _gs_common import run_reference, scene_tensors  # noqa: E402
from oracle import build_ref  # noqa: E402
from oracle import gs_oracle as G  # noqa: E402

scene = G.synthetic_scene(P=400, H=48, W=80, seed=0)
ref = run_reference(build_ref.load_rasterizer(), scene_tensors(scene, "cuda"))
os.makedirs("gpurun_out", exist_ok=True)
torch.save({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in ref.items()}, "gpurun_out/gs_render_small.pt")
print("gs golden:", ref["num_rendered"], "instances,", int((ref["radii"] > 0).sum()), "visible Gaussians")
