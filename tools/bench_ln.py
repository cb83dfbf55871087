"""LayerNorm + AdaLN (folded A/B tables) alone on config 2's shape: us per launch inside a CUDA graph of 60 launches
(what the forward does), next to a plain copy of the same tensor.  (under gpurun)"""
import sys
import torch
sys.path.insert(0, ".")
from orv_b200 import ops
dev = "cuda"
M, D = 3226, 1920
x = torch.randn(M, D, device=dev).bfloat16()
ab = torch.randn(6, 4 * D, device=dev).bfloat16()
rm = ops.rowmap(seq_len=M, text_len=226, tokens_per_group=600, groups_per_batch=6)
out = torch.empty_like(x)
def chain(n):
    for _ in range(n):
        ops.ln_modulate(x, None, None, 1e-5, rm=rm, ab=ab, out=out)
def chain_copy(n):
    for _ in range(n):
        out.copy_(x)
for name, fn in (("ln_ab", chain), ("copy", chain_copy)):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(3)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn(60)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20 / 60 * 1e3:.2f} us per launch (graph of 60, {M}x{D} bf16 = {M * D * 2 / 1e6:.1f} MB in + out)")
