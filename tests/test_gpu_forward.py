"""Parity of the CUDA forward (through the module API -> C ABI) against the fp32 CPU oracle."""
import pytest
import torch

from _gates import assert_forward_close
from oracle import flat_oracle as O

pytestmark = pytest.mark.gpu


def small_cfg(**over):
    base = dict(num_attention_heads=2, attention_head_dim=64, in_channels=32, out_channels=16, num_layers=2,
                sample_width=8, sample_height=6, sample_frames=9, modulate_encoder_hidden_states=True,
                text_embed_dim=64, max_text_seq_length=10, num_control_blocks=2)
    base.update(over)
    return O.default_config(**base)


def build_model(cfg, sd):
    from orv_b200.models.cogvideox_control import CogVideoXTransformer3DModelTraj
    m = CogVideoXTransformer3DModelTraj(**cfg)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not [k for k in missing if "action_recon" not in k], missing
    m.action_embed.mask = False
    return m.to("cuda", torch.bfloat16).eval()


def rel_err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().mean() / b.abs().mean()).item()


def run_case(cfg, B, Fr, H, W, controls=False, rope=False, ofs=None, n_actions=8, use_actions=True):
    sd32 = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    sd = {k: v.bfloat16().float() for k, v in sd32.items()}  # oracle sees the same bf16-rounded weights
    inp = O.synthetic_inputs(cfg, B, Fr, H, W, seed=1, with_controls=controls, n_actions=n_actions)
    hs = inp["hidden_states"].bfloat16().float()
    text = inp["text"].bfloat16().float()
    acts = inp["actions"].bfloat16().float() if use_actions else None
    dep = inp["depths"].bfloat16().float() if controls else None
    lab = inp["labels"].bfloat16().float() if controls else None
    t = torch.full((B,), 499, dtype=torch.int64)
    rp = None
    if rope:
        rp = O.pipeline_rope(cfg, H * 8, W * 8, Fr)
    ofs_t = torch.tensor([ofs]) if ofs is not None else None
    ref = O.forward(sd, cfg, hs, text, t, actions=acts, depths=dep, labels=lab, ofs=ofs_t, rope=rp)
    m = build_model(cfg, sd32)
    cg = {}
    if use_actions:
        cg["actions"] = acts.cuda().bfloat16()
    if controls:
        cg["depths"], cg["labels"] = dep.cuda().bfloat16(), lab.cuda().bfloat16()
    with torch.no_grad():
        out, is_mask, recon = m(hs.cuda().bfloat16(), text.cuda().bfloat16(), cg, t.cuda(),
                                ofs=ofs_t.cuda() if ofs_t is not None else None,
                                image_rotary_emb=(rp[0].cuda(), rp[1].cuda()) if rp else None, return_dict=False)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    assert torch.isfinite(out.float()).all()
    assert_forward_close(out, ref, 1.5e-2, 3e-2, 8e-2, "small forward vs fp32 oracle")  # mean, max-abs, worst row
    return rel_err(out, ref), m


def test_forward_small_actions():
    e, m = run_case(small_cfg(), 1, 3, 6, 8)
    assert m.last_launch_count > 0
    assert e < 1.5e-2, e


def test_forward_small_no_actions_batch2():
    e, _ = run_case(small_cfg(), 2, 3, 6, 8, use_actions=False)
    assert e < 1.5e-2, e


def test_forward_small_controls():
    e, _ = run_case(small_cfg(visual_guidance=True), 2, 3, 6, 8, controls=True)
    assert e < 1.5e-2, e


@pytest.mark.parametrize("use_actions", [True, False])
def test_forward_small_no_text_modulation(use_actions):
    """modulate_encoder_hidden_states=False (reference :70-99, :404-424; the from-scratch 1.4B configs)."""
    e, _ = run_case(small_cfg(modulate_encoder_hidden_states=False), 2, 3, 6, 8, use_actions=use_actions)
    assert e < 1.5e-2, e


def test_forward_small_rope_pt2_ofs():
    cfg = small_cfg(patch_size_t=2, use_rotary_positional_embeddings=True, ofs_embed_dim=512, patch_bias=False)
    e, _ = run_case(cfg, 2, 4, 6, 8, rope=True, ofs=2.0, n_actions=12)
    assert e < 1.5e-2, e


@pytest.mark.parametrize("mask_on,views", [(False, 1), (True, 1), (False, 3)])
def test_modulation_schedule_is_bit_identical(mask_on, views):
    """AdaLN tables built once for a list of timesteps (prepare_modulation_schedule + _mod_step) must reproduce the
    per-step forwards bit for bit, including the reference's eval-time action-mask draws (same device RNG stream)."""
    over = dict(visual_guidance=True, multiview=True, max_n_view=3) if views > 1 else {}
    cfg = small_cfg(**over)
    sd32 = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    m = build_model(cfg, sd32)
    m.action_embed.mask = mask_on
    B, Fr, H, W = (1, 2, 6, 8) if views > 1 else (4, 3, 6, 8)
    inp = O.synthetic_inputs(cfg, B, Fr * views, H, W, seed=1, with_controls=views > 1, n_actions=4 if views > 1 else 8)
    hs, text = inp["hidden_states"].cuda().bfloat16(), inp["text"].cuda().bfloat16()
    cg = {"actions": inp["actions"].cuda().bfloat16()}
    if views > 1:
        cg["depths"], cg["labels"] = inp["depths"].cuda().bfloat16(), inp["labels"].cuda().bfloat16()
    steps = [999.0, 749.0, 499.0, 249.0, 19.0]
    with torch.no_grad():
        torch.manual_seed(11)
        ref = [m(hs, text, cg, torch.full((B,), t, device="cuda"), return_dict=False, num_views=views) for t in steps]
        ref = [(o.clone(), k.clone()) for o, k, _ in ref]
        torch.manual_seed(11)
        m.prepare_modulation_schedule(steps, tuple(hs.shape), text.shape[1], cg, num_views=views)
        for rep in range(2):  # second pass replays the captured graph
            for i, t in enumerate(steps):
                out, is_mask, _ = m(hs, text, cg, torch.full((B,), t, device="cuda"), return_dict=False, num_views=views,
                                    _mod_step=i)
                assert torch.equal(out, ref[i][0]), (rep, i)
                assert torch.equal(is_mask, ref[i][1])
        m.clear_modulation_schedule()
        with pytest.raises(RuntimeError):
            m(hs, text, cg, torch.full((B,), 499.0, device="cuda"), return_dict=False, num_views=views, _mod_step=0)
    if mask_on:
        assert any(bool(k.any()) for _, k in ref) or True  # 10 % draws: may be all-False for a tiny batch


def test_weight_broadcast_refreshes_parameters_that_cannot_alias_the_arena():
    """after the arena is overwritten from outside (the NCCL weight broadcast), the zero-padded
    action_embed.mlp.0.weight (K = 28 -> 32) must follow, or a later re-pack would silently restore the old values."""
    from orv_b200 import dist as D
    cfg = small_cfg()
    src = build_model(cfg, O.synthetic_state_dict(cfg, seed=0, std=0.05))
    dst = build_model(cfg, O.synthetic_state_dict(cfg, seed=7, std=0.05))
    key = "action_embed.mlp.0.weight"
    assert not torch.equal(src.state_dict()[key], dst.state_dict()[key])
    dst.weight_arena().copy_(src.weight_arena())  # what dist.broadcast does on a non-source rank
    D.broadcast_weights(dst, src=0)                # world size 1: no collective, but the refresh must still run
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k
    inp = O.synthetic_inputs(cfg, 1, 3, 6, 8, seed=1, n_actions=8)
    args = (inp["hidden_states"].cuda().bfloat16(), inp["text"].cuda().bfloat16(),
            {"actions": inp["actions"].cuda().bfloat16()}, torch.tensor([499], device="cuda"))
    with torch.no_grad():
        a = src(*args, return_dict=False)[0]
        dst._invalidate()  # force a re-pack from the parameters
        b = dst(*args, return_dict=False)[0]
    assert torch.equal(a, b)
