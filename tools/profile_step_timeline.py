"""What the GPU does between two forwards of the denoise loop (host gaps, the table select copies, the noise upload,
the sampler step): device-side activity records (CUPTI through torch.profiler) of one pipeline call, printed for one
iteration boundary, plus the idle time per iteration.

    python tools/profile_step_timeline.py [config=2] [clips=0]        (under gpurun)
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
    import time
    from torch.autograd import DeviceType
    from torch.profiler import ProfilerActivity, profile
    # host view of one clip without a profiler: enqueue time of the call, then the wait for the GPU to drain
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(200 + rep)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"clip {rep}: pipe() returned after {1e3 * (t1 - t0):.1f} ms (host), GPU drained {1e3 * (t2 - t1):.1f} ms later, "
              f"total {1e3 * (t2 - t0):.1f} ms")
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        run(300)
        torch.cuda.synchronize()
    evs = sorted((e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == DeviceType.CUDA)
    samp = [i for i, e in enumerate(evs) if "sampler_step_kernel" in e[2]]
    print(f"{len(evs)} device activities, {len(samp)} sampler steps")
    first_fwd = next(i for i, e in enumerate(evs) if "patchify_kernel" in e[2])
    print(f"clip prologue on the device: first activity -> first forward kernel {evs[first_fwd][0] - evs[0][0]:.0f} us "
          f"({first_fwd} activities, busy {sum(e - s for s, e, _ in evs[:first_fwd]):.0f} us); denoise loop "
          f"{evs[samp[-1]][1] - evs[first_fwd][0]:.0f} us; after the last sampler step {evs[-1][1] - evs[samp[-1]][1]:.0f} us")
    cpu = sorted(((e.time_range.end - e.time_range.start, e.name) for e in prof.events()
                  if e.device_type == DeviceType.CPU and e.time_range.end - e.time_range.start > 200), reverse=True)[:14]
    print("longest host-side ops of the clip (us):")
    for d, n in cpu:
        print(f"  {d:9.0f}  {n[:80]}")
    busy = sum(e - s for s, e, _ in evs)
    span = evs[-1][1] - evs[0][0]
    print(f"span {span / 1e3:.2f} ms, sum of activity durations {busy / 1e3:.2f} ms")
    # idle gaps > 5 us, aggregated by the activity that follows them
    gaps = {}
    prev_end = evs[0][1]
    for s, e, n in evs[1:]:
        g = s - prev_end
        if g > 5:
            k = n[:60]
            gaps.setdefault(k, [0.0, 0])
            gaps[k][0] += g
            gaps[k][1] += 1
        prev_end = max(prev_end, e)
    print("idle before (total us, count):")
    for k, (t, c) in sorted(gaps.items(), key=lambda kv: -kv[1][0])[:12]:
        print(f"  {t:10.1f} us  x{c:4d}  {k}")
    if len(samp) > 12:
        a, b = samp[10], samp[11]
        # the boundary after sampler step 10: last 3 kernels of the forward before it ... first 6 of the next forward
        lo = max(a - 3, 0)
        t0 = evs[lo][0]
        print("one iteration boundary (us from the first line):")
        for s, e, n in evs[lo:a + 12]:
            print(f"  {s - t0:9.1f} .. {e - t0:9.1f}  ({e - s:7.1f})  {n[:90]}")
        print(f"iteration length (sampler to sampler): {evs[b][1] - evs[a][1]:.1f} us")


if __name__ == "__main__":
    main()
