"""Times the four block GEMMs of config 2 (and two VAE convolutions) for every library named on the command line
(ORVB_LIB_PATH variants built with different -DORVB_G2_* settings), one subprocess each, same inputs.
    python tools/bench_gemm_variants.py orv_b200/liborv_b200.so orv_b200/liborv_b200_a.so ..."""
import os
import subprocess
import sys

CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from orv_b200 import _lib as L, ops
dev = "cuda"
torch.manual_seed(0)
def t(call, n=50):
    for _ in range(5): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
M, D = 3226, 1920
resid = torch.randn(M, D, device=dev).bfloat16()
gate = torch.randn(6, 6 * D, device=dev)
rm = ops.rowmap(seq_len=M, text_len=226, tokens_per_group=600, groups_per_batch=6)
res = []
def gemm(name, N, K, epi, **kw):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev).bfloat16(); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    us = t(lambda: ops.gemm(a, w, b, epilogue=epi, out=out, **kw))
    ref = None
    res.append(f"{name} {us:.1f}")
    return out
o1 = gemm("out", D, D, L.EPI_GATE_RESID, resid=resid, gate=gate, gate_text_off=0, gate_video_off=0, rm=rm)
gemm("ff1", 4 * D, D, L.EPI_GELU)
gemm("ff2", D, 4 * D, L.EPI_GATE_RESID, resid=resid, gate=gate, gate_text_off=0, gate_video_off=0, rm=rm)
gemm("qkv", 3 * D, D, L.EPI_BIAS)
for (T, H, W, ci, co) in [(9, 240, 360, 128, 128), (9, 120, 180, 256, 256)]:
    x = (torch.randn(T, H, W, ci, device=dev) * 0.5).bfloat16(); w = (torch.randn(co, 27 * ci, device=dev) * 0.02).bfloat16()
    b = torch.randn(co, device=dev).bfloat16(); r = torch.randn(T, H, W, co, device=dev).bfloat16()
    out = torch.empty(T, H, W, co, device=dev, dtype=torch.bfloat16)
    us = t(lambda: ops.conv_cl(x, w, b, (3, 3, 3), resid=r, out=out), 10)
    fl = 2.0 * T * H * W * 27 * ci * co
    res.append(f"conv{ci} {us:.0f}us={fl / us / 1e6:.0f}TF")
print(" | ".join(res), "| checksum", float(o1.float().abs().sum()))
'''
for lib in sys.argv[1:]:
    env = dict(os.environ, ORVB_LIB_PATH=lib, ORVB_NO_BUILD="1")
    for rep in range(1 if os.environ.get("NCU_ONCE") else 2):
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
        print(f"{os.path.basename(lib)}: {r.stdout.strip() or r.stderr.strip()[-400:]}", flush=True)
