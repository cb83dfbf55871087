"""GPU bring-up + tile sweep for the CTA-pair GEMM (tcgen05 cta_group::2).  Run under gpurun:
    python tools/gpu_check_gemm2.py smoke      one tiny case (run first, under `timeout`: a protocol bug hangs)
    python tools/gpu_check_gemm2.py            correctness over shapes / widths / epilogues, then the timing sweep
"""
import sys

import torch

sys.path.insert(0, ".")
from orv_b200 import ops, _lib as L  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def check(name, got, ref, tol=2e-2):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    rel = err / (ref.abs().max().item() + 1e-9)
    ok = rel < tol and torch.isfinite(got).all().item()
    print(f"[{'OK' if ok else 'FAIL'}] {name}: max_abs_err={err:.4g} rel={rel:.4g}", flush=True)
    if not ok:
        d = (got - ref).abs()
        M, N = d.shape
        for i in range(0, min(M, 512), 64):
            print(" ".join(f"{d[i:i+64, j:j+32].max().item():8.3g}" for j in range(0, min(N, 320), 32)), flush=True)
    return ok


def plain(M, N, K, bn):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    out = ops.gemm(a, w, b, bn=bn)
    torch.cuda.synchronize()
    return check(f"plain M={M} N={N} K={K} bn={bn}", out, a.float() @ w.float().T + b.float())


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print("device:", torch.cuda.get_device_name(0), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "smoke":
    ok = plain(256, 256, 64, -256)
    ok &= plain(256, 256, 256, -256)
    print("SMOKE_OK" if ok else "SMOKE_FAILED", flush=True)
    sys.exit(0 if ok else 1)

ok = True
for (M, N, K, bn) in [(256, 128, 64, -128), (256, 256, 512, -256), (384, 512, 512, -192), (200, 136, 72, -64),
                      (1000, 1920, 256, -176), (1000, 1920, 256, -240), (1000, 1920, 256, -208), (3226, 1920, 1920, -176),
                      (3226, 5760, 1920, -192), (3226, 7680, 1920, -240), (3226, 1920, 7680, -256), (3226, 64, 1920, -64),
                      (3226, 1920, 1920, 0)]:
    try:
        ok &= plain(M, N, K, bn)
    except Exception as e:  # noqa: BLE001
        print(f"[EXC] {M} {N} {K} {bn}: {e}", flush=True)
        ok = False
        break

if ok:  # epilogues through the pair kernel (auto selection at M > 128)
    M, D = 3226, 1920
    a = (torch.randn(M, D, device=dev) * 0.5).bfloat16()
    w = (torch.randn(4 * D, D, device=dev) * 0.05).bfloat16()
    b = torch.randn(4 * D, device=dev).bfloat16()
    out = ops.gemm(a, w, b, epilogue=L.EPI_GELU)
    ref = torch.nn.functional.gelu(a.float() @ w.float().T + b.float(), approximate="tanh")
    ok &= check("gelu (auto)", out, ref)
    out2 = ops.gemm(a, w, b, epilogue=L.EPI_GELU, bn=256)
    ok &= check("gelu pair == 1-CTA", out, out2, tol=1e-6)

if ok:
    shapes = {"qkv": (3226, 5760, 1920), "out": (3226, 1920, 1920), "ff1": (3226, 7680, 1920), "ff2": (3226, 1920, 7680),
              "big": (8192, 8192, 8192)}
    for name, (M, N, K) in shapes.items():
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        res = []
        for bn in (128, 192, 256, -128, -160, -176, -192, -208, -224, -240, -256):
            ms = timeit(lambda: ops.gemm(a, w, b, bn=bn, out=out))
            res.append((bn, ms))
        ms_c = timeit(lambda: torch.nn.functional.linear(a, w, b))
        fl = 2.0 * M * N * K
        print(f"{name} M={M} N={N} K={K}: " + "  ".join(f"{bn}:{ms*1e3:.1f}us/{fl/ms/1e9:.0f}TF" for bn, ms in res)
              + f"  cuBLAS:{ms_c*1e3:.1f}us/{fl/ms_c/1e9:.0f}TF", flush=True)
print("ALL_OK" if ok else "SOME_FAILED", flush=True)
sys.exit(0 if ok else 1)
