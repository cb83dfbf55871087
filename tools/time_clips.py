"""Per-clip timing of back-to-back pipeline calls (device events + host wall clock), to spot intermittent stalls.
    python tools/time_clips.py [config=2] [n=8]        (under gpurun)"""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
run, model, info = bench.build_runner(cid, 0)
for i in range(3):
    run(40 + i)
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
host = []
evs[0].record()
for i in range(n):
    t0 = time.perf_counter()
    run(100 + i)
    host.append(time.perf_counter() - t0)
    evs[i + 1].record()
torch.cuda.synchronize()
dev = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
print("device ms per clip:", " ".join(f"{d:.0f}" for d in dev))
print("host   ms per clip:", " ".join(f"{h * 1e3:.0f}" for h in host))
