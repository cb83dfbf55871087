"""The CTA-pair GEMM's gated-residual epilogue with TMA-prefetched residual tiles (`fast_resid`, orv_b200/csrc/gemm.cu):
used when `out` is written in place over `resid` (attn-out and FF2 of every block, reference
orv/models/cogvideox_control.py:419-421, :442-443).

It must produce the SAME BITS as the generic epilogue (ORVB_GEMM_FAST_RESID=0) and as the single-CTA kernel, and match
a float64 torch evaluation within one bf16 rounding of the output (max |err| / max |ref| <= 1e-2 like the other bf16
operator tests; the fp32-output tight test of the same arithmetic lives in tests/test_gpu_tight.py).
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from orv_b200 import ops as _ops
    return _ops


def _case(ops, M, N, K, S, St, tpf, G, bn, gate_on=True, bias_on=True, seed=0):
    from orv_b200 import _lib as L
    torch.manual_seed(seed)
    B = M // S if S > 0 else 1
    a = (torch.randn(M, K, device=DEV) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=DEV) * 0.1).bfloat16()
    b = torch.randn(N, device=DEV).bfloat16() if bias_on else None
    x0 = torch.randn(M, N, device=DEV).bfloat16()
    gate = torch.randn(max(B * G, 1), 6 * N, device=DEV) if gate_on else None
    rm = ops.rowmap(S, St, tpf, G) if S > 0 else None
    kw = dict(epilogue=L.EPI_GATE_RESID, rm=rm)
    if gate_on:
        kw.update(gate=gate, gate_text_off=5 * N, gate_video_off=2 * N)

    def run(fast, bn_):
        os.environ["ORVB_GEMM_FAST_RESID"] = "1" if fast else "0"
        x = x0.clone()
        ops.gemm(a, w, b, resid=x, out=x, bn=bn_, **kw)
        torch.cuda.synchronize()
        return x

    try:
        fast, slow = run(True, bn), run(False, bn)
    finally:
        os.environ.pop("ORVB_GEMM_FAST_RESID", None)
    # float64 evaluation
    lin = a.double() @ w.double().T + (b.double() if bias_on else 0.0)
    if gate_on:
        r = torch.arange(M, device=DEV)
        if S > 0:
            s, bi = r % S, r // S
            grp = bi * G + torch.where((s < St) | (tpf <= 0), torch.zeros_like(s), 1 + (s - St) // max(tpf, 1))
            is_text = s < St
        else:
            grp, is_text = torch.zeros_like(r), torch.zeros_like(r, dtype=torch.bool)
        gv = torch.where(is_text[:, None], gate[grp][:, 5 * N:6 * N], gate[grp][:, 2 * N:3 * N]).double()
        lin = lin * gv
    ref = x0.double() + lin
    return fast, slow, ref


SHAPES = [
    # M, N, K, S, St, tpf, G, bn                      (bn < 0: CTA-pair kernel, tile width |bn|)
    (3226, 1920, 1920, 3226, 226, 600, 6, 0),         # config 2 attn-out (picks 176: two full units + a 48-wide tail)
    (3226, 1920, 512, 3226, 226, 600, 6, -192),       # three full units, no tail
    (3226, 1920, 512, 3226, 226, 600, 6, -128),       # two units: one per warp of a pair
    (3226, 1920, 512, 3226, 226, 600, 6, -64),        # one unit: warps 8-11 idle
    (3226, 1920, 512, 3226, 226, 600, 6, -80),        # 64 + 16-wide tail
    (2 * 1013, 1536, 512, 1013, 113, 300, 4, -176),   # two sequences, ragged M (last 256-row tile partly empty)
    (700, 200, 64, 70, 10, 20, 4, -176),              # tiny groups: > 2 gate rows inside 32 output rows (global fallback), N % 64 != 0
    (520, 1920, 256, 0, 0, 0, 1, -176),               # no row map: one gate row
    (300, 96, 64, 0, 0, 0, 1, -48),                   # tile narrower than a unit: tail only
]


@pytest.mark.parametrize("M,N,K,S,St,tpf,G,bn", SHAPES)
def test_fast_resid_bit_identical_to_generic(ops, M, N, K, S, St, tpf, G, bn):
    fast, slow, ref = _case(ops, M, N, K, S, St, tpf, G, bn)
    assert torch.isfinite(fast.float()).all()
    assert torch.equal(fast, slow), f"{(fast != slow).sum().item()} elements differ"
    err = ((fast.double() - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-2, err  # one bf16 rounding of the output


@pytest.mark.parametrize("gate_on,bias_on", [(False, True), (True, False), (False, False)])
def test_fast_resid_optional_operands(ops, gate_on, bias_on):
    fast, slow, ref = _case(ops, 1000, 1920, 256, 500, 100, 100, 5, -176, gate_on=gate_on, bias_on=bias_on, seed=3)
    assert torch.equal(fast, slow)
    assert ((fast.double() - ref).abs().max() / ref.abs().max()).item() < 1e-2


def test_fast_resid_matches_single_cta_kernel(ops):
    """Same K order, same epilogue arithmetic: the 128-row single-CTA kernel must give the same bits."""
    from orv_b200 import _lib as L
    torch.manual_seed(5)
    M, N, K, S, St, tpf, G = 1000, 1920, 320, 500, 100, 100, 5
    a = (torch.randn(M, K, device=DEV) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=DEV) * 0.1).bfloat16()
    b = torch.randn(N, device=DEV).bfloat16()
    x0 = torch.randn(M, N, device=DEV).bfloat16()
    gate = torch.randn(2 * G, 6 * N, device=DEV)
    rm = ops.rowmap(S, St, tpf, G)
    outs = []
    for bn in (-176, 128):
        x = x0.clone()
        ops.gemm(a, w, b, epilogue=L.EPI_GATE_RESID, resid=x, out=x, gate=gate, gate_text_off=5 * N,
                 gate_video_off=2 * N, rm=rm, bn=bn)
        outs.append(x)
    assert torch.equal(outs[0], outs[1])


def test_fast_resid_repeatable_and_back_to_back(ops):
    """Several launches in a row on one stream (staging tiles and barriers start clean each launch) give the same bits."""
    ref = None
    for _ in range(3):
        fast, _, _ = _case(ops, 3226, 1920, 1920, 3226, 226, 600, 6, 0, seed=7)
        if ref is None:
            ref = fast
        assert torch.equal(fast, ref)
