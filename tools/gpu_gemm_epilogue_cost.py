"""Where do the block GEMMs lose time inside the forward?  (run under gpurun)
Times the four GEMMs of a CogVideoX-2B block in isolation with (a) plain bias epilogue vs the fused epilogue the
forward uses, (b) weights L2-resident (same matrix every launch) vs streamed from HBM (rotating over enough matrices
to exceed the 126 MB L2), each over a ~0.5 s loop so clocks settle under the power cap."""
import sys

import torch

sys.path.insert(0, ".")
from orv_b200 import ops, _lib as L  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
S, St, D, FF, H = 3226, 226, 1920, 7680, 30
rm = ops.rowmap(S, St, 600, 6)


import pynvml  # noqa: E402

pynvml.nvmlInit()
_h = pynvml.nvmlDeviceGetHandleByIndex(0)


def timeit(fn, n_w, iters=6000):
    """~0.5 s loop; returns (us per launch, SM MHz and power W sampled while the tail of the loop is still running)."""
    for i in range(5):
        fn(i % n_w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_w)
    e1.record()
    mhz = pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)   # host is ahead of the GPU: the queue is still draining
    watts = pynvml.nvmlDeviceGetPowerUsage(_h) / 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3, mhz, watts


def bench(name, M, N, K, make_call):
    n_rot = max(2, int(400e6 // (N * K * 2)) + 1)
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    ws = [(torch.randn(N, K, device=dev) * 0.05).bfloat16() for _ in range(n_rot)]
    b = torch.randn(N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    plain = lambda i: ops.gemm(a, ws[i], b, out=out)  # noqa: E731
    fused = make_call(a, ws, b, out)
    fl = 2.0 * M * N * K
    res = {}
    for label, fn, nw in (("plain/L2", plain, 1), ("plain/HBM", plain, n_rot), ("fused/L2", fused, 1), ("fused/HBM", fused, n_rot)):
        res[label] = timeit(fn, nw)
    cub = timeit(lambda i: torch.nn.functional.linear(a, ws[i], b), n_rot)
    print(f"{name:4s} M={M} N={N} K={K}:\n   " + "\n   ".join(f"{k:10s} {v[0]:6.1f}us {fl / v[0] / 1e6:5.0f} TF  {v[1]} MHz {v[2]:.0f} W" for k, v in res.items())
          + f"\n   cuBLAS/HBM {cub[0]:6.1f}us {fl / cub[0] / 1e6:5.0f} TF  {cub[1]} MHz {cub[2]:.0f} W", flush=True)


def qkv_call(a, ws, b, out):
    qn = (torch.ones(64, device=dev).bfloat16(), torch.zeros(64, device=dev).bfloat16())
    return lambda i: ops.gemm(a, ws[i], b, out=out, epilogue=L.EPI_QKV, qk_dim=D, q_norm=qn, k_norm=qn, qk_eps=1e-6, rm=rm)


def gelu_call(a, ws, b, out):
    return lambda i: ops.gemm(a, ws[i], b, out=out, epilogue=L.EPI_GELU)


def gate_call(a, ws, b, out):
    gate = torch.randn(6, 6 * D, device=dev)
    x = torch.randn(S, D, device=dev).bfloat16()
    return lambda i: ops.gemm(a, ws[i], b, out=x, epilogue=L.EPI_GATE_RESID, resid=x, gate=gate, gate_text_off=5 * D,
                              gate_video_off=2 * D, rm=rm)


print("device:", torch.cuda.get_device_name(0), flush=True)
bench("qkv", S, 3 * D, D, qkv_call)
bench("out", S, D, D, gate_call)
bench("ff1", S, FF, D, gelu_call)
bench("ff2", S, D, FF, gate_call)
