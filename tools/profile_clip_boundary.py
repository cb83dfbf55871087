"""Two pipeline calls back to back under CUPTI + CPU tracing: does the host prologue of the second clip overlap the GPU
tail of the first?  Prints, from the moment the first pipe() returns: the host-side runtime calls longer than 150 us
(in order), and the GPU's idle gaps longer than 150 us.

    python tools/profile_clip_boundary.py [config=2] [clips=0]        (under gpurun)
"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402


def main():
    cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    clips = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    run, model, _ = bench.build_runner(cid, clips)
    for i in range(3):
        run(40 + i)
    torch.cuda.synchronize()
    from torch.autograd import DeviceType
    from torch.profiler import ProfilerActivity, profile, record_function
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        with record_function("CLIP_A"):
            run(300)
        with record_function("CLIP_B"):
            run(301)
        torch.cuda.synchronize()
    evs = list(prof.events())
    marks = {e.name: e.time_range for e in evs if e.name in ("CLIP_A", "CLIP_B")}
    t0 = marks["CLIP_B"].start
    print(f"CLIP_A host {marks['CLIP_A'].end - marks['CLIP_A'].start:.0f} us, CLIP_B host {marks['CLIP_B'].end - t0:.0f} us")
    dev = sorted((e.time_range.start, e.time_range.end, e.name) for e in evs if e.device_type == DeviceType.CUDA)
    samp = [i for i, e in enumerate(dev) if "sampler_step_kernel" in e[2]]
    print(f"last sampler step of clip A ends at {dev[samp[49]][1] - t0:.0f} us after clip B's pipe() started; "
          f"clip B's first patchify at {next(s for s, e, n in dev if 'patchify' in n and s > dev[samp[49]][1]) - t0:.0f} us")
    print("host calls > 150 us from the start of clip B's pipe() to +40 ms:")
    for e in sorted((e for e in evs if e.device_type == DeviceType.CPU), key=lambda e: e.time_range.start):
        d = e.time_range.end - e.time_range.start
        if d > 150 and -1000 <= e.time_range.start - t0 <= 40000 and not e.name.startswith("CLIP"):
            print(f"  {e.time_range.start - t0:9.0f}  {d:8.0f}  {e.name[:90]}")
    print("GPU idle gaps > 150 us between -50 ms and +60 ms:")
    prev_end = dev[0][1]
    for s, e, n in dev[1:]:
        if s - prev_end > 150 and -50000 <= s - t0 <= 60000:
            print(f"  idle {prev_end - t0:9.0f} .. {s - t0:9.0f} ({s - prev_end:7.0f} us) before {n[:70]}")
        prev_end = max(prev_end, e)


if __name__ == "__main__":
    main()
