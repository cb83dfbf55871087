"""Attention kernel alone: us per launch on the shapes of the BASELINE configs (200-launch loops, CUDA events, L2-hot
qkv), next to torch SDPA on the same inputs.     python tools/bench_attention.py        (under gpurun)"""
import sys

import torch

sys.path.insert(0, ".")
from orv_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
sms = torch.cuda.get_device_properties(0).multi_processor_count
for name, (B, S, H, q0, nq) in {"config 2 / 3 (B=1, S=3226, 30 heads)": (1, 3226, 30, 0, 0),
                                "config 2, 2 clips": (2, 3226, 30, 0, 0),
                                "config 4 (CFG pair, S=2026, 48 heads)": (2, 2026, 48, 0, 0),
                                "config 5 temporal (6 sequences, S=2146)": (6, 2146, 30, 0, 0),
                                "config 5 view blocks (10 x 1830 keys, 1152 queries)": (10, 1830, 30, 678, 1152)}.items():
    qkv = torch.randn(B * S, 3 * H * 64, device=dev).bfloat16()
    pairs = -(-(nq or S) // 256)
    res = []
    for rep in range(2):
        out = ops.attention(qkv, B, S, H, 0.125, q_row0=q0, q_rows=nq)
        for _ in range(5):
            ops.attention(qkv, B, S, H, 0.125, out=out, q_row0=q0, q_rows=nq)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            ops.attention(qkv, B, S, H, 0.125, out=out, q_row0=q0, q_rows=nq)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 200 * 1e3)
    q, k, v = qkv.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    if nq:
        q = q[:, :, q0:q0 + nq]
    for _ in range(5):
        torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=0.125)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=0.125)
    e1.record()
    torch.cuda.synchronize()
    sdpa = e0.elapsed_time(e1) / 100 * 1e3
    flop = 4.0 * B * (nq or S) * S * H * 64
    print(f"{name}: {pairs * H * B} CTAs = {pairs * H * B / sms:.2f} waves: {res[0]:.1f} / {res[1]:.1f} us "
          f"({flop / res[1] / 1e6:.0f} TFLOP/s); torch SDPA (strided q/k/v views of the same buffer) {sdpa:.1f} us", flush=True)
