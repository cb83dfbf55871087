"""CTA-pair GEMM with a mixed tile list (orv_b200/csrc/gemm_common.cuh g2_tile): N cut into n / width full tiles plus one
narrower tile per 256-row block, the narrow tiles placed on the SM pairs that got one full tile less.  Every output
element is still one K-ordered tcgen05 accumulation followed by the same epilogue, so the result must carry the SAME
BITS as the uniform tile list (ORVB_GEMM_MIXED_TILES=0) — QKV projection of a block, reference
orv/models/cogvideox_control.py:232-234 with the QK LayerNorm of :243-254 in the epilogue."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from orv_b200 import ops as _ops
    return _ops


def _both(fn):
    outs = []
    try:
        for flag in ("1", "0"):
            os.environ["ORVB_GEMM_MIXED_TILES"] = flag
            outs.append(fn())
            torch.cuda.synchronize()
    finally:
        os.environ.pop("ORVB_GEMM_MIXED_TILES", None)
    return outs


@pytest.mark.parametrize("M,K", [(3226, 1920), (3226, 256), (6452, 320)])
def test_qkv_mixed_tiles_bit_identical(ops, M, K):
    from orv_b200 import _lib as L
    lib = L.load()
    D = 1920
    assert lib.orvb_gemm_tile_remainder(M, 3 * D, L.EPI_QKV) > 0, "this shape is expected to use the mixed tile list"
    torch.manual_seed(M + K)
    a = (torch.randn(M, K, device=DEV) * 0.5).bfloat16()
    w = (torch.randn(3 * D, K, device=DEV) * 0.1).bfloat16()
    b = torch.randn(3 * D, device=DEV).bfloat16()
    qn = ((1 + 0.1 * torch.randn(64, device=DEV)).bfloat16(), (0.1 * torch.randn(64, device=DEV)).bfloat16())
    kn = ((1 + 0.1 * torch.randn(64, device=DEV)).bfloat16(), (0.1 * torch.randn(64, device=DEV)).bfloat16())
    rm = ops.rowmap(3226, 226, 600, 6)
    mixed, uniform = _both(lambda: ops.gemm(a, w, b, epilogue=L.EPI_QKV, qk_dim=D, q_norm=qn, k_norm=kn, rm=rm))
    assert torch.isfinite(mixed.float()).all()
    assert torch.equal(mixed, uniform), f"{(mixed != uniform).sum().item()} elements differ"
    # and against fp32 torch (bf16 output: one rounding of the largest value)
    lin = a.float() @ w.float().T + b.float()
    q, k, v = lin.split(D, dim=1)

    def hn(t, p):
        return torch.nn.functional.layer_norm(t.view(M, D // 64, 64), (64,), p[0].float(), p[1].float(), 1e-6).reshape(M, D)

    ref = torch.cat([hn(q, qn), hn(k, kn), v], 1)
    assert ((mixed.float() - ref).abs().max() / ref.abs().max()).item() < 1e-2


def test_bias_gemm_shapes_with_a_remainder_tile(ops):
    """Bias epilogue over shapes for which the picker chooses a remainder tile (whatever it picks, both lists must agree)."""
    from orv_b200 import _lib as L
    lib = L.load()
    hit = 0
    for (M, N, K) in [(3226, 5760, 128), (3226, 2176, 128), (4052, 3200, 192), (3226, 9344, 64), (5000, 1408, 64)]:
        hit += lib.orvb_gemm_tile_remainder(M, N, L.EPI_BIAS) > 0
        torch.manual_seed(N)
        a = (torch.randn(M, K, device=DEV) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=DEV) * 0.1).bfloat16()
        b = torch.randn(N, device=DEV).bfloat16()
        mixed, uniform = _both(lambda: ops.gemm(a, w, b))
        assert torch.equal(mixed, uniform), (M, N, K)
        ref = a.float() @ w.float().T + b.float()
        assert ((mixed.float() - ref).abs().max() / ref.abs().max()).item() < 1e-2
    assert hit >= 1
