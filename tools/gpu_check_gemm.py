"""GPU bring-up check for the tcgen05 GEMM (run under gpurun). Prints per-case error stats and a coarse
error map when a case fails, so descriptor / swizzle bugs can be diagnosed from one run."""
import sys
import time

import torch

sys.path.insert(0, ".")
from orv_b200 import ops, _lib as L  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def errmap(got, ref, bs=32):
    d = (got.float() - ref.float()).abs()
    M, N = d.shape
    rows = []
    for i in range(0, min(M, 256), bs):
        rows.append(" ".join(f"{d[i:i+bs, j:j+bs].max().item():8.3g}" for j in range(0, min(N, 256), bs)))
    return "\n".join(rows)


def check(name, got, ref, tol=2e-2):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs().max().item()
    rel = err / (ref.abs().max().item() + 1e-9)
    ok = rel < tol and torch.isfinite(got).all().item()
    print(f"[{'OK' if ok else 'FAIL'}] {name}: max_abs_err={err:.4g} rel={rel:.4g} ref_max={ref.abs().max().item():.4g}",
          flush=True)
    if not ok:
        print(errmap(got, ref), flush=True)
    return ok


def run_plain(M, N, K, bn):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    out = ops.gemm(a, w, b, bn=bn)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T + b.float()
    return check(f"plain M={M} N={N} K={K} bn={bn}", out, ref)


all_ok = True
print("device:", torch.cuda.get_device_name(0), "lib version", L.load().orvb_version(), flush=True)
for (M, N, K, bn) in [(128, 128, 64, 128), (128, 128, 128, 128), (128, 64, 64, 64), (128, 256, 64, 256),
                      (256, 256, 256, 128), (384, 512, 512, 256), (200, 136, 72, 128), (3226, 1920, 1920, 128),
                      (3226, 5760, 1920, 256), (3226, 64, 1920, 64)]:
    try:
        all_ok &= run_plain(M, N, K, bn)
    except Exception as e:  # noqa: BLE001
        print(f"[EXC] plain {M} {N} {K} {bn}: {e}", flush=True)
        all_ok = False
        break

# ---- epilogues ------------------------------------------------------------------------------------
if all_ok:
    M, N, K = 3226, 7680, 1920
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    out = ops.gemm(a, w, b, epilogue=L.EPI_GELU)
    ref = torch.nn.functional.gelu(a.float() @ w.float().T + b.float(), approximate="tanh")
    all_ok &= check("gelu", out, ref)

    # gated residual with the joint-sequence row map (B=1, S=3226, text 226, 600 tokens/frame, 6 groups)
    S, St, tpf, G = 3226, 226, 600, 6
    D = 1920
    a = (torch.randn(M, 7680, device=dev) * 0.5).bfloat16()
    w = (torch.randn(D, 7680, device=dev) * 0.02).bfloat16()
    b = torch.randn(D, device=dev).bfloat16()
    resid = torch.randn(M, D, device=dev).bfloat16()
    gate = torch.randn(G, 6 * D, device=dev)
    rm = ops.rowmap(S, St, tpf, G)
    out = ops.gemm(a, w, b, epilogue=L.EPI_GATE_RESID, resid=resid, gate=gate, gate_text_off=5 * D,
                   gate_video_off=2 * D, rm=rm)
    lin = a.float() @ w.float().T + b.float()
    grp = torch.zeros(M, dtype=torch.long, device=dev)
    grp[St:] = 1 + (torch.arange(M - St, device=dev) // tpf)
    gvec = torch.where((grp == 0)[:, None], gate[grp][:, 5 * D:6 * D], gate[grp][:, 2 * D:3 * D])
    ref = resid.float() + gvec * lin
    all_ok &= check("gate_resid", out, ref)

    # in-place variant (resid aliases out)
    x = resid.clone()
    ops.gemm(a, w, b, epilogue=L.EPI_GATE_RESID, resid=x, gate=gate, gate_text_off=5 * D, gate_video_off=2 * D,
             rm=rm, out=x)
    all_ok &= check("gate_resid in-place", x, ref)

    # QKV with per-head LayerNorm and RoPE
    a = (torch.randn(M, D, device=dev) * 0.5).bfloat16()
    w = (torch.randn(3 * D, D, device=dev) * 0.05).bfloat16()
    b = torch.randn(3 * D, device=dev).bfloat16()
    qn = ((1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16())
    kn = ((1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16())
    ang = torch.rand(S - St, 32, device=dev) * 6.28
    cos = torch.cos(ang).repeat_interleave(2, dim=1).contiguous()
    sin = torch.sin(ang).repeat_interleave(2, dim=1).contiguous()
    for rope in (None, (cos, sin)):
        out = ops.gemm(a, w, b, epilogue=L.EPI_QKV, qk_dim=D, q_norm=qn, k_norm=kn, qk_eps=1e-6, rm=rm, rope=rope)
        lin = a.float() @ w.float().T + b.float()
        q, k, v = lin.split(D, dim=1)

        def hn(t, p):
            t = t.view(M, D // 64, 64)
            t = torch.nn.functional.layer_norm(t, (64,), p[0].float(), p[1].float(), 1e-6)
            if rope is not None:
                tv = t[St:]
                xr = torch.stack([-tv[..., 1::2], tv[..., 0::2]], -1).flatten(-2)
                t = torch.cat([t[:St], tv * cos[:, None, :] + xr * sin[:, None, :]], 0)
            return t.reshape(M, D)

        ref = torch.cat([hn(q, qn), hn(k, kn), v], 1)
        all_ok &= check(f"qkv rope={'yes' if rope else 'no'}", out, ref)

    # row remap + positional add (patch-embed style): per-batch blocks of 3000 rows into a 3226-row sequence
    Bt, Sv = 2, 3000
    a = (torch.randn(Bt * Sv, 128, device=dev)).bfloat16()
    w = (torch.randn(D, 128, device=dev) * 0.1).bfloat16()
    b = torch.randn(D, device=dev).bfloat16()
    pos = torch.randn(Sv, D, device=dev).bfloat16()
    out = torch.zeros(Bt * S, D, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, b, epilogue=L.EPI_GATE_RESID, resid=pos, resid_mod=Sv, out=out, row_remap=(Sv, S, St))
    lin = (a.float() @ w.float().T + b.float()).view(Bt, Sv, D) + pos.float()
    ref = torch.zeros(Bt, S, D, device=dev)
    ref[:, St:] = lin
    all_ok &= check("row-remap + pos add", out, ref.view(Bt * S, D))

# ---- timing ----------------------------------------------------------------------------------------
if all_ok:
    for (M, N, K) in [(3226, 5760, 1920), (3226, 1920, 1920), (3226, 7680, 1920), (3226, 1920, 7680),
                      (8192, 8192, 8192)]:
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        for bn in (64, 128, 256):
            for _ in range(3):
                out = ops.gemm(a, w, b, bn=bn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 20
            e0.record()
            for _ in range(iters):
                ops.gemm(a, w, b, bn=bn, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print(f"time M={M} N={N} K={K} bn={bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        for _ in range(3):
            ref = torch.nn.functional.linear(a, w, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ref = torch.nn.functional.linear(a, w, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"time M={M} N={N} K={K} cuBLAS: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)

print("ALL_OK" if all_ok else "SOME_FAILED", flush=True)
sys.exit(0 if all_ok else 1)
