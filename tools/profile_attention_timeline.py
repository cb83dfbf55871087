"""clock64 timeline of CTA (0,0,0) of the attention kernel (run under gpurun with a -DORVB_ATT_TIMELINE variant library:
ORVB_BUILD_VARIANT=tl ORVB_EXTRA_NVCC_FLAGS=-DORVB_ATT_TIMELINE python -m orv_b200.build, then ORVB_LIB_PATH=orv_b200/liborv_b200_tl.so)."""
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
n = 26
print("per softmax warp, mean over j=4..23: [s_full wait | S copy + s_free | max + exps | o_full wait + P st + arrive] period")
for w in range(8):
    a = sum(int(d[w, j, 1]) - int(d[w, j, 0]) for j in range(4, 24)) / 20.0
    b = sum(int(d[w, j, 2]) - int(d[w, j, 1]) for j in range(4, 24)) / 20.0
    c = sum(int(d[w, j, 3]) - int(d[w, j, 2]) for j in range(4, 24)) / 20.0
    e = sum(int(d[w, j + 1, 0]) - int(d[w, j, 3]) for j in range(4, 24)) / 20.0
    per = (int(d[w, 24, 0]) - int(d[w, 4, 0])) / 20.0
    print(f"warp {w}: {a:7.1f} {b:7.1f} {c:7.1f} {e:7.1f}   period {per:7.1f}")
for t in (0, 1):
    a = sum(int(d[16 + t, j, 1]) - int(d[16 + t, j, 0]) for j in range(4, 24)) / 20.0
    b = sum(int(d[16 + t, j, 3]) - int(d[16 + t, j, 2]) for j in range(4, 24)) / 20.0
    print(f"MMA tile {t}: QK issue block {a:7.1f} clk, PV issue block {b:7.1f} clk")
# issuer view of one steady-state iteration
j = 10
base = int(d[16, j, 0])
print("issuer j=10 (relative clk): QK0 start/end, QK1 start/end, PV0 start/end, PV1 start/end, next QK0 start:",
      [int(d[16, j, 0]) - base, int(d[16, j, 1]) - base, int(d[17, j, 0]) - base, int(d[17, j, 1]) - base,
       int(d[16, j, 2]) - base, int(d[16, j, 3]) - base, int(d[17, j, 2]) - base, int(d[17, j, 3]) - base,
       int(d[16, j + 1, 0]) - base])
for w in (0, 4):
    print(f"warp {w} j=10 (relative to issuer QK0 start): top {int(d[w, j, 0]) - base}, s_full {int(d[w, j, 1]) - base}, "
          f"copied {int(d[w, j, 2]) - base}, exps done {int(d[w, j, 3]) - base}, next top {int(d[w, j + 1, 0]) - base}")
w = 0
start, end = int(d[w, 30, 0]), int(d[w, 30, 1])
print(f"CTA(0,0,0) warp 0: total {end - start} clk; prologue {int(d[w, 0, 1]) - start}; main loop {int(d[w, n - 1, 3]) - int(d[w, 0, 1])} "
      f"= {(int(d[w, n - 1, 3]) - int(d[w, 0, 1])) / n:.0f} per key step; epilogue {end - int(d[w, n - 1, 3])}")
