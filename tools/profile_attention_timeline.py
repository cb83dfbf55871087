"""Prints the clock64 timeline of CTA (0,0,0) of the attention kernel (run under gpurun).
Needs a library built with the stamps compiled in:  ORVB_EXTRA_NVCC_FLAGS=-DORVB_ATT_TIMELINE python -m orv_b200.build --force"""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from orv_b200 import ops, _lib as L
B, S, H = 1, 3226, 30
qkv = torch.randn(B * S, 3 * H * 64, device="cuda").bfloat16()
out = ops.attention(qkv, B, S, H, 0.125)
for _ in range(2):
    ops.attention(qkv, B, S, H, 0.125, out=out)
dbg = torch.zeros(18 * 128, dtype=torch.int64, device="cuda")
lib = L.load()
lib.orvb_attention_set_debug.argtypes = [C.c_void_p]
lib.orvb_attention_set_debug(dbg.data_ptr())
ops.attention(qkv, B, S, H, 0.125, out=out)
torch.cuda.synchronize()
lib.orvb_attention_set_debug(None)
d = dbg.cpu().view(18, 32, 4)
t0 = int(d[d > 0].min())
def rel(x):
    return int(x) - t0 if x > 0 else -1
per_iter = (int(d[0, 24, 1]) - int(d[0, 4, 1])) / 20.0
print("clk per iteration (warp 0, j=4..24):", per_iter)
print("per softmax warp, mean over j=4..23: [S copy wait + s_free arrive | max + exp + P stores | s_full wait + prefetch issue | fence + p_full arrive]")
for w in range(16):
    a = sum(int(d[w, j, 2]) - int(d[w, j, 1]) for j in range(4, 24)) / 20.0
    b = sum(int(d[w, j, 3]) - int(d[w, j, 2]) for j in range(4, 24)) / 20.0
    c = sum(int(d[w, j, 0]) - int(d[w, j, 3]) for j in range(4, 24)) / 20.0
    e = sum(int(d[w, j + 1, 1]) - int(d[w, j, 0]) for j in range(4, 24)) / 20.0
    print(f"warp {w:2d}: {a:7.1f} {b:7.1f} {c:7.1f} {e:7.1f}   start of j=10: {rel(d[w, 10, 1])}")
for t in (0, 1):
    a = sum(int(d[16 + t, j, 1]) - int(d[16 + t, j, 0]) for j in range(4, 24)) / 20.0
    b = sum(int(d[16 + t, j, 3]) - int(d[16 + t, j, 2]) for j in range(4, 24)) / 20.0
    print(f"MMA tile {t}: QK issue {a:7.1f} clk, PV issue (incl. v_full wait) {b:7.1f} clk")
