"""One process per (stagger, emu) setting: time the attention kernel on the BASELINE shapes and check it against SDPA.
Needs a measurement build:  ORVB_BUILD_VARIANT=exp ORVB_EXTRA_NVCC_FLAGS=-DORVB_EXPERIMENTAL python -m orv_b200.build
(the product library has the shipped setting compiled in and ignores the knobs)."""
import os
import subprocess
import sys

CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from orv_b200 import ops
torch.manual_seed(0)
res = []
for (B, S, H) in [(1, 3226, 30), (2, 2026, 48)]:
    qkv = torch.randn(B * S, 3 * H * 64, device="cuda").bfloat16()
    qkv[:, : H * 64] *= 2.0
    out = ops.attention(qkv, B, S, H, 0.125)
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=0.125).permute(0, 2, 1, 3).reshape(B * S, H * 64)
    err = (out.float() - ref).abs().max().item()
    for _ in range(5):
        ops.attention(qkv, B, S, H, 0.125, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        ops.attention(qkv, B, S, H, 0.125, out=out)
    e1.record()
    torch.cuda.synchronize()
    res.append(f"B={B} S={S} H={H}: {e0.elapsed_time(e1) * 10:.1f} us maxerr {err:.2e}")
print(" | ".join(res))
'''
settings = [("v5", s, e) for e in (0, 1, 2, 3) for s in (0, 600, 1200, 1800)]
if len(sys.argv) > 1:
    settings = [("v5", int(a.split(",")[0]), int(a.split(",")[1])) for a in sys.argv[1:]]
for kind, stagger, emu in settings:
    env = dict(os.environ)
    if kind == "v5":
        env.update(ORVB_LIB_PATH=os.environ.get("ORVB_SWEEP_LIB", "orv_b200/liborv_b200_exp.so"), ORVB_ATT_STAGGER=str(stagger),
                   ORVB_ATT_EMU=str(emu))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    print(f"{kind} stagger={stagger} emu={emu}/8: {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
