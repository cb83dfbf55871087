"""GPU tests of EXPERIMENTAL, opt-in kernels that have not run on hardware yet.  They are skipped unless
ORVB_TEST_EXPERIMENTAL=1 (the default `pytest -m gpu` run covers only what ships on the default path):

    ORVB_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_zz_gpu_experimental.py -m gpu -q
"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("ORVB_TEST_EXPERIMENTAL") != "1",
                                 reason="experimental kernels: set ORVB_TEST_EXPERIMENTAL=1")]


def _ff_pair(M, D, seed):
    from orv_b200 import _lib as L, ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s, k=1.0: (torch.randn(*s, generator=g, device="cuda") * k)  # noqa: E731
    xn = r(M, D, k=0.5).bfloat16()
    w1, b1 = r(4 * D, D, k=0.05).bfloat16(), r(4 * D).bfloat16()
    w2, b2 = r(D, 4 * D, k=0.05).bfloat16(), r(D).bfloat16()
    x = r(M, D).bfloat16()
    gate = r(6, 6 * D)
    rm = ops.rowmap(seq_len=M, text_len=min(226, M // 4), tokens_per_group=max(8, (M - min(226, M // 4)) // 5 + 1),
                    groups_per_batch=6)
    kw = dict(resid=None, gate=gate, gate_text_off=5 * D, gate_video_off=2 * D, rm=rm)
    return L, ops, xn, w1, b1, w2, b2, x, kw


@pytest.mark.parametrize("M,D", [(300, 256), (1000, 512), (3226, 1920)])
def test_ff_chain_is_bit_identical_to_two_launches(M, D):
    L, ops, xn, w1, b1, w2, b2, x, kw = _ff_pair(M, D, seed=M)
    # reference: the shipped path, two launches (FF2 accumulates into a copy of x)
    ffh_ref = ops.gemm(xn, w1, b1, epilogue=L.EPI_GELU)
    x_ref = x.clone()
    ops.gemm(ffh_ref, w2, b2, epilogue=L.EPI_GATE_RESID, out=x_ref, **dict(kw, resid=x_ref))
    # chained launch
    ffh = torch.empty_like(ffh_ref)
    x_out = x.clone()
    first = ops.gemm(xn, w1, b1, epilogue=L.EPI_GELU, out=ffh, launch=False)
    second = ops.gemm(ffh, w2, b2, epilogue=L.EPI_GATE_RESID, out=x_out, launch=False, **dict(kw, resid=x_out))
    for _ in range(3):  # repeated launches: the counters are re-zeroed by every call
        x_out.copy_(x)
        ops.gemm_chain(first, second)
    torch.cuda.synchronize()
    assert torch.equal(ffh, ffh_ref)
    assert torch.equal(x_out, x_ref)


def test_forward_with_ff_chain_matches_default(monkeypatch):
    """Whole small forward with ORVB_FF_CHAIN=1 in a child process against the default path (bit-identical)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = (
        "import sys, torch; sys.path.insert(0, %r)\n"
        "from oracle import flat_oracle as O\n"
        "from orv_b200 import CogVideoXTransformer3DModelTraj\n"
        "cfg = O.default_config(num_attention_heads=4, attention_head_dim=64, num_layers=2, sample_width=24, "
        "sample_height=16, sample_frames=9, text_embed_dim=128, max_text_seq_length=16, modulate_encoder_hidden_states=True)\n"
        "sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)\n"
        "inp = O.synthetic_inputs(cfg, 1, 3, 16, 24, seed=1, n_actions=8)\n"
        "m = CogVideoXTransformer3DModelTraj(**cfg); m.load_state_dict(sd, strict=False); m.action_embed.mask = False\n"
        "m = m.to('cuda', torch.bfloat16).eval()\n"
        "c = lambda t: t.to('cuda', torch.bfloat16)\n"
        "with torch.no_grad():\n"
        "    y = m(c(inp['hidden_states']), c(inp['text']), {'actions': c(inp['actions'])}, torch.tensor([499], device='cuda'), "
        "return_dict=False)[0]\n"
        "torch.save(y.cpu(), sys.argv[1])\n") % root
    outs = []
    for flag in ("0", "1"):
        path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"orvb_ff_chain_{flag}.pt")
        env = dict(os.environ, ORVB_FF_CHAIN=flag)
        subprocess.run([sys.executable, "-c", prog, path], check=True, env=env, timeout=300)
        outs.append(torch.load(path))
    assert torch.isfinite(outs[1].float()).all()
    assert torch.equal(outs[0], outs[1])
