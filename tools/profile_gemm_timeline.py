"""Clock-stamp timeline of the CTA-pair GEMM (first and last cluster) on the four block shapes of config 2 — where the
~20 us per launch go in which the tensor pipe is idle (open lead 2).  Run under gpurun with a library built
with the stamps compiled in:

    ORVB_EXTRA_NVCC_FLAGS=-DORVB_GEMM_TIMELINE python -m orv_b200.build --force      (here, before gpurun)
    python tools/profile_gemm_timeline.py                                            (on the box)
    python -m orv_b200.build --force                                                 (back to the default build)

Slots per (cluster, CTA rank): 0 entry, 1 prologue done, 2 after griddepcontrol.wait, 3 first stage requested (TMA
warp), 4 last stage requested, 5 first operands landed (MMA warp, leader), 6 first tile's MMAs issued, 7 all MMAs
issued, 8 first accumulator complete (epilogue warp 4), 9 first tile's epilogue done, 10 last accumulator complete,
11 last tile's epilogue done, 12 output stores complete, 13 teardown barrier passed, 14 = tiles processed."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from orv_b200 import _lib as L, ops  # noqa: E402

lib = L.load()
if not hasattr(lib, "orvb_gemm_set_debug"):
    sys.exit("liborv_b200.so was built without -DORVB_GEMM_TIMELINE (see the docstring)")
lib.orvb_gemm_set_debug.argtypes = [C.c_void_p]
lib.orvb_gemm_set_debug.restype = C.c_int32
dev = "cuda"
torch.manual_seed(0)
NAMES = ["entry", "prologue", "pdl_wait", "tma_first", "tma_last", "ops_landed", "mma_tile0", "mma_all", "acc_first",
         "epi_first", "acc_last", "epi_last", "stores", "teardown"]


def run(name, M, N, K, epilogue=L.EPI_BIAS, inplace=False, **kw):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    # inplace: written over its own residual, as the forward does (-> the TMA-prefetched epilogue unless
    # ORVB_GEMM_FAST_RESID=0)
    out = kw["resid"] if inplace else torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    call = lambda: ops.gemm(a, w, b, epilogue=epilogue, out=out, **kw)  # noqa: E731
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    dbg = torch.zeros(128, dtype=torch.int64, device=dev)
    L.check(lib.orvb_gemm_set_debug(dbg.data_ptr()))
    call()
    torch.cuda.synchronize()
    L.check(lib.orvb_gemm_set_debug(None))
    d = dbg.cpu().view(2, 2, 32)  # [first | last cluster][rank][slot]
    print(f"== {name}: M={M} N={N} K={K} tile width {lib.orvb_gemm_tile_width(M, N, epilogue)}: {us:.1f} us per launch "
          f"(L2-hot loop of 20)")
    for ci, cname in enumerate(("first cluster", "last cluster")):
        for r in (0, 1):
            row = d[ci, r]
            t0 = int(row[0])
            if t0 == 0:
                continue
            rel = {n: (int(row[i]) - t0 if int(row[i]) > 0 else None) for i, n in enumerate(NAMES) if n}
            print(f"  {cname} CTA {r} ({int(row[14])} tiles), clk from entry: "
                  + "  ".join(f"{n}={v}" for n, v in rel.items() if v is not None))
    return us


M, D = 3226, 1920
resid = (torch.randn(M, D, device=dev)).bfloat16()
gate = torch.randn(6, 6 * D, device=dev)
rm = ops.rowmap(seq_len=M, text_len=226, tokens_per_group=600, groups_per_batch=6)
run("attn-out (gate*x + residual)", M, D, D, L.EPI_GATE_RESID, resid=resid, gate=gate, gate_text_off=0, gate_video_off=0, rm=rm)
gate_s = gate * 0.01  # (in place the residual stream is accumulated over the timing loop: keep it bounded)
run("attn-out IN PLACE (gate*x + residual)", M, D, D, L.EPI_GATE_RESID, inplace=True, resid=resid.clone(), gate=gate_s,
    gate_text_off=0, gate_video_off=0, rm=rm)
run("FF2 IN PLACE (gate*x + residual)", M, D, 4 * D, L.EPI_GATE_RESID, inplace=True, resid=resid.clone(), gate=gate_s,
    gate_text_off=0, gate_video_off=0, rm=rm)
run("FF1 (GELU)", M, 4 * D, D, L.EPI_GELU)
run("FF2 (gate*x + residual)", M, D, 4 * D, L.EPI_GATE_RESID, resid=resid, gate=gate, gate_text_off=0, gate_video_off=0, rm=rm)
run("QKV-shaped (bias only)", M, 3 * D, D, L.EPI_BIAS)
run("attn-out shape, bias only", M, D, D, L.EPI_BIAS)
