import sys
import torch
sys.path.insert(0, ".")
from orv_b200 import ops
B, S, H = 1, 3226, 30
qkv = torch.randn(B * S, 3 * H * 64, device="cuda").bfloat16()
out = ops.attention(qkv, B, S, H, 0.125)
for _ in range(3):
    ops.attention(qkv, B, S, H, 0.125, out=out)
torch.cuda.synchronize()
print("ok")
