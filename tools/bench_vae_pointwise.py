"""Isolated bandwidth of the memory-bound VAE kernels on the decoder's shapes (under gpurun)."""
import sys
import torch
sys.path.insert(0, ".")
from orv_b200 import ops
dev = "cuda"
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (T, H, W, C, sh) in [(9, 240, 360, 128, 3), (9, 120, 180, 256, 2), (5, 60, 90, 256, 1), (3, 30, 45, 512, 0), (9, 240, 360, 256, 3)]:
    x = torch.randn(T, H, W, C, device=dev).bfloat16()
    h, w = H >> sh, W >> sh
    table = torch.randn(3 * h * w, 2 * C, device=dev).bfloat16()
    gamma, beta = torch.ones(C, device=dev).bfloat16(), torch.zeros(C, device=dev).bfloat16()
    tsrc = torch.tensor([min(i * 3 // T, 2) for i in range(T)], dtype=torch.int32, device=dev)
    out = torch.empty_like(x)
    us_s = t(lambda: ops.gn_stats_cl(x, 32, 1e-6))
    st = ops.gn_stats_cl(x, 32, 1e-6)
    us_n = t(lambda: ops.spatial_norm_cl(x, st, gamma, beta, table, 0, C, tsrc, (h, w), sh, out=out))
    us_c = t(lambda: out.copy_(x))
    mb = x.numel() * 2 / 1e6
    print(f"[{T}x{H}x{W}x{C}] {mb:.0f} MB: gn_stats {us_s:.1f} us = {mb / us_s / 1e3 * 1e3:.0f} GB/s | spatial_norm {us_n:.1f} us = "
          f"{2 * mb / us_n:.0f} GB/s (r+w) | torch copy {us_c:.1f} us = {2 * mb / us_c:.0f} GB/s", flush=True)
