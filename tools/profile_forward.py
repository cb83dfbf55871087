"""Runs a few eager (no CUDA graph) config-2 forwards + sampler steps for ncu (see profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, ".")
os.environ["ORVB_CUDA_GRAPH"] = "0"
# The pipeline builds the AdaLN tables once per clip (61 small tensor-core GEMMs, outside the per-step forward); this tool
# calls the model directly, where they would sit inside every forward and shift the ncu launch windows of
# tools/gpu_round.sh.  Keep the single batched CUDA-core launch here: the per-step kernels are the ones being profiled.
os.environ.setdefault("ORVB_MOD_TABLES_TC", "0")
from bench import config2, init_weights_  # noqa: E402
from orv_b200 import CogVideoXDPMScheduler, CogVideoXTransformer3DModelTraj  # noqa: E402

n_fwd = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda")
with torch.device(dev):
    model = CogVideoXTransformer3DModelTraj(**config2())
init_weights_(model, 0)
model = model.to(torch.bfloat16).eval()
model.action_embed.mask = False
hs = torch.randn(1, 5, 32, 40, 60, device=dev).bfloat16()
text = (torch.randn(1, 226, 4096, device=dev) * 0.2).bfloat16()
act = torch.randn(1, 16, 7, device=dev).bfloat16()
lat = torch.randn(1, 5, 16, 40, 60, device=dev).bfloat16()
sched = CogVideoXDPMScheduler(timestep_spacing="trailing")
sched.set_timesteps(50)
old = torch.zeros(1, 5, 16, 40, 60, device=dev)
noise = torch.randn(1, 5, 16, 40, 60, device=dev).bfloat16()
with torch.no_grad():
    for i in range(n_fwd):
        out = model(hs, text, {"actions": act}, torch.tensor([979 - 20 * i], device=dev), return_dict=False)[0]
        sched.fused_step(out, old, i > 0, 979 - 20 * i, 999 - 20 * i, lat, noise, 1, 1.0, hs)
torch.cuda.synchronize()
print("done", model.last_launch_count)
