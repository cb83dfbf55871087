"""Occupancy -> depth / semantic-map rendering: orvb_gs_rasterize next to the reference's own CUDA
extension (oracle/_ref) on the same B200, same inputs.  One JSON line.
    python tools/bench_gs_render.py [P=200000] [H=320] [W=480]        (under gpurun)"""
import json
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from _gs_common import occupancy_scene, run_ours, run_reference, scene_tensors  # noqa: E402
from oracle import build_ref  # noqa: E402

_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
P = int(_pos[0]) if len(_pos) > 0 else 200000
H = int(_pos[1]) if len(_pos) > 1 else 320
W = int(_pos[2]) if len(_pos) > 2 else 480
s = scene_tensors(occupancy_scene(P, H, W), "cuda")


def timed(fn, n=30):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


ms, out = timed(lambda: run_ours(s))
line = {"workload": f"{P} voxel Gaussians (0.012 m isotropic, opacity 1, 12 one-hot semantic channels) into a {H}x{W} frame",
        "instances": out["num_rendered"], "ours_ms_per_frame": round(ms, 4), "ours_frames_per_s": round(1e3 / ms, 1),
        "ours_gaussians_per_s": round(P / ms * 1e3)}
if build_ref.rasterizer_available():
    ms_ref, ref = timed(lambda: run_reference(build_ref.load_rasterizer(), s))
    line.update(reference_cuda_ms_per_frame=round(ms_ref, 4), speedup_vs_reference_cuda=round(ms_ref / ms, 2),
                max_abs_diff_color=float((out["color"] - ref["color"]).abs().max()))
print(json.dumps(line))
if "--kernels" in sys.argv:  # device-side activity records of one call of each implementation
    from torch.autograd import DeviceType
    from torch.profiler import ProfilerActivity, profile
    for name, fn in (("ours", lambda: run_ours(s)), ("reference", lambda: run_reference(build_ref.load_rasterizer(), s))):
        if name == "reference" and not build_ref.rasterizer_available():
            continue
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        evs = sorted((e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == DeviceType.CUDA)
        agg = {}
        for a, b, n in evs:
            k = n.replace("void ", "").replace("orvb::", "").replace("(anonymous namespace)::", "")[:48]
            agg.setdefault(k, [0.0, 0])
            agg[k][0] += b - a
            agg[k][1] += 1
        print(f"--- {name}: {len(evs)} device activities, span {evs[-1][1] - evs[0][0]:.0f} us, busy {sum(b - a for a, b, _ in evs):.0f} us",
              file=sys.stderr)
        for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
            print(f"   {t:8.1f} us x{c:3d}  {k}", file=sys.stderr)
