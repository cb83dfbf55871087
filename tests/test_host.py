"""CPU: host-side logic of the drop-in (no kernels are launched)."""
import os

import pytest
import torch

from oracle import flat_oracle as O
from oracle import make_golden as G


def small_cfg(**over):
    return O.default_config(**dict(G.BASE, **over))


def make_model(**over):
    from orv_b200 import CogVideoXTransformer3DModelTraj
    return CogVideoXTransformer3DModelTraj(**small_cfg(**over))


@pytest.mark.parametrize("over", [dict(), dict(visual_guidance=True), dict(multiview=True, visual_guidance=True),
                                  dict(patch_size_t=2, use_rotary_positional_embeddings=True, ofs_embed_dim=64)])
def test_state_dict_keys_and_shapes_match_reference_layout(over):
    m = make_model(**over)
    shapes = O.param_shapes(small_cfg(**over))
    sd = m.state_dict()
    assert set(sd) == set(shapes)
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k


def test_config_defaults_and_attributes():
    from orv_b200 import CogVideoXTransformer3DModelTraj
    m = CogVideoXTransformer3DModelTraj()
    ref = O.default_config()
    for k, v in ref.items():
        assert m.config[k] == v, k
    assert m.config.patch_size_t is None and m.config.ofs_embed_dim is None
    assert len(m.transformer_blocks) == 30 and m.gradient_checkpointing is False
    assert m.action_embed.mask is True  # reference quirk: mask=self.training evaluated at construction
    assert dict(m.config)["num_attention_heads"] == 30


def test_zero_init_of_control_and_multiview_layers():
    m = make_model(visual_guidance=True, multiview=True)
    assert m.initial_combine_linear.weight.abs().max() == 0
    assert m.mv_blocks[0].proj_out.weight.abs().max() == 0
    assert all(not p.requires_grad for p in m.transformer_blocks.parameters())
    assert all(p.requires_grad for p in m.mv_blocks.parameters())


def test_reference_constructor_errors():
    from orv_b200 import CogVideoXTransformer3DModelTraj
    with pytest.raises(ValueError, match="num_tracking_blocks"):
        CogVideoXTransformer3DModelTraj(**small_cfg(visual_guidance=True, num_control_blocks=5))
    with pytest.raises(RuntimeError, match="modulate_encoder_hidden_states"):
        CogVideoXTransformer3DModelTraj(**small_cfg(modulate_encoder_hidden_states=False,
                                                    loaded_pretrained_model_name_or_path="THUDM/CogVideoX-2b"))


def test_forward_has_no_cpu_fallback():
    m = make_model().eval()
    inp = O.synthetic_inputs(small_cfg(), 1, 3, 6, 8, n_actions=8)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        m(inp["hidden_states"], inp["text"], {"actions": inp["actions"]}, torch.tensor([499]))


def test_action_embed_prepare_matches_reference_bookkeeping():
    m = make_model()
    acts = torch.arange(2 * 11 * 7, dtype=torch.float32).reshape(2, 11, 7)  # already (n+1)%4==0
    x, is_mask, apply = m.action_embed.prepare(acts)
    assert x.shape == (2, 3, 28) and is_mask.shape == (2,) and apply.dtype == torch.uint8
    ref = torch.cat([torch.zeros(2, 1, 7), acts], 1).reshape(2, 3, 28)
    assert torch.equal(x, ref)
    with pytest.raises(ValueError, match="mismatched"):
        m.action_embed.prepare(torch.zeros(1, 3, 6))
    # the random draw happens on every call and follows the global RNG like the reference
    torch.manual_seed(5)
    _, a, _ = m.action_embed.prepare(torch.zeros(16, 11, 7))
    torch.manual_seed(5)
    assert torch.equal(a, torch.rand(16) < 0.1)


def test_save_and_from_pretrained_roundtrip(tmp_path):
    from orv_b200 import CogVideoXTransformer3DModelTraj
    cfg = small_cfg(visual_guidance=True)
    m = CogVideoXTransformer3DModelTraj(**cfg)
    m.load_state_dict(O.synthetic_state_dict(cfg, seed=3), strict=False)
    d = tmp_path / "ckpt" / "transformer"
    m.save_pretrained(str(d))
    import json
    conf = json.load(open(d / "config.json"))
    assert conf["_class_name"] == "CogVideoXTransformer3DModelTraj" and conf["visual_guidance"] is True
    m2 = CogVideoXTransformer3DModelTraj.from_pretrained(str(tmp_path / "ckpt"), subfolder="transformer",
                                                         torch_dtype=torch.bfloat16)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert b.dtype == torch.bfloat16 and torch.equal(a.bfloat16(), b), k


def test_from_pretrained_widens_t2v_checkpoint(tmp_path):
    """Reference :1014-1030: a 16-channel THUDM CogVideoX-2b checkpoint becomes a 32-channel model whose extra input
    channels are zero."""
    import json
    from safetensors.torch import save_file
    from orv_b200 import CogVideoXTransformer3DModelTraj
    cfg = small_cfg(in_channels=16)
    sd = {k: v for k, v in O.synthetic_state_dict(cfg, seed=4).items() if not k.startswith("action_embed.")}
    d = tmp_path / "THUDM" / "CogVideoX-2b" / "transformer"
    os.makedirs(d)
    save_file(sd, str(d / "diffusion_pytorch_model.safetensors"))
    json.dump(dict({k: v for k, v in cfg.items()}, _class_name="CogVideoXTransformer3DModel"),
              open(d / "config.json", "w"))
    m = CogVideoXTransformer3DModelTraj.from_pretrained(str(tmp_path / "THUDM" / "CogVideoX-2b"), subfolder="transformer")
    assert m.config.in_channels == 32 and m.config.from_t2v is True
    w = m.patch_embed.proj.weight
    assert w.shape[1] == 32 and torch.equal(w[:, :16], sd["patch_embed.proj.weight"]) and w[:, 16:].abs().max() == 0


# ---- schedulers -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [2, 5, 50])
def test_scheduler_timesteps_and_alphas_match_oracle(n):
    from orv_b200 import CogVideoXDDIMScheduler, CogVideoXDPMScheduler
    o = O.Scheduler("dpm")
    o.set_timesteps(n)
    for cls in (CogVideoXDDIMScheduler, CogVideoXDPMScheduler):
        s = cls(timestep_spacing="trailing")
        s.set_timesteps(n)
        assert s.timesteps.tolist() == o.timesteps.tolist()
        assert torch.equal(s.alphas_cumprod, o.alphas_cumprod)
        assert s.order == 1 and s.init_noise_sigma == 1.0


def test_scheduler_steps_match_oracle_fp32():
    from orv_b200 import CogVideoXDDIMScheduler, CogVideoXDPMScheduler
    torch.manual_seed(0)
    n = 5
    x0 = torch.randn(2, 3, 4)
    vs = [torch.randn(2, 3, 4) for _ in range(n)]
    s, o = CogVideoXDDIMScheduler(timestep_spacing="trailing"), O.Scheduler("ddim")
    s.set_timesteps(n); o.set_timesteps(n)
    a = b = x0
    for i, t in enumerate(o.timesteps.tolist()):
        a = s.step(vs[i], t, a, return_dict=False)[0]
        b = o.step_ddim(vs[i], t, b)
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-6)
    s, o = CogVideoXDPMScheduler(timestep_spacing="trailing"), O.Scheduler("dpm")
    s.set_timesteps(n); o.set_timesteps(n)
    g1, g2 = torch.Generator().manual_seed(1), torch.Generator().manual_seed(1)
    a = b = x0
    olda = oldb = None
    ts = o.timesteps.tolist()
    assert s.noise_draws(n) == [1, 2, 2, 2, 1]
    for i, t in enumerate(ts):
        a, olda = s.step(vs[i], olda, t, ts[i - 1] if i > 0 else None, a, generator=g1, return_dict=False)
        b, oldb = o.step_dpm(vs[i], oldb, t, ts[i - 1] if i > 0 else None, b, g2)
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-6)


def test_fused_step_refuses_cpu_tensors():
    from orv_b200 import CogVideoXDDIMScheduler
    s = CogVideoXDDIMScheduler(timestep_spacing="trailing")
    s.set_timesteps(2)
    with pytest.raises(RuntimeError, match="CUDA"):
        s.fused_step(torch.zeros(4, dtype=torch.bfloat16), 999, torch.zeros(4, dtype=torch.bfloat16))


# ---- pipeline host logic --------------------------------------------------------------------------------------
def _pipe():
    from orv_b200 import CogVideoXDPMScheduler, CogVideoXImageToVideoPipelineTraj
    from orv_b200.models.pipeline_control import default_vae_config
    return CogVideoXImageToVideoPipelineTraj(None, None, default_vae_config(), make_model(),
                                             CogVideoXDPMScheduler(timestep_spacing="trailing"))


def test_pipeline_rejects_foreign_transformer():
    from orv_b200 import CogVideoXDPMScheduler, CogVideoXImageToVideoPipelineTraj
    with pytest.raises(ValueError, match="CogVideoXTransformer3DModelTraj"):
        CogVideoXImageToVideoPipelineTraj(None, None, None, torch.nn.Linear(1, 1), CogVideoXDPMScheduler())


def test_pipeline_check_inputs_positional_quirk():
    """Reference :1261-1270 passes prompt_embeds positionally into the `latents` slot (SURVEY probe P7)."""
    p = _pipe()
    img = torch.zeros(1, 32, 1, 6, 8)
    emb = torch.zeros(1, 10, 32)
    with pytest.raises(ValueError, match="Provide either"):
        p(image=img, prompt_embeds=emb, height=48, width=64, num_frames=9, output_type="latent")
    with pytest.raises(ValueError, match="divisible by 8"):
        p(image=img, prompt="", prompt_embeds=emb, height=50, width=64, num_frames=9, output_type="latent")


def test_prepare_latents_matches_oracle():
    p = _pipe()
    g = torch.Generator().manual_seed(7)
    moments = torch.randn(1, 32, 1, 6, 8, generator=g)
    lat, img = p.prepare_latents(moments, 1, 16, 9, 1, 48, 64, torch.float32, torch.device("cpu"),
                                 torch.Generator().manual_seed(42), None)
    gen = torch.Generator().manual_seed(42)
    mean, logvar = moments.chunk(2, dim=1)
    ref_img = (1.15258426 * (mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * torch.randn(mean.shape, generator=gen)))
    ref_img = ref_img.permute(0, 2, 1, 3, 4)
    assert img.shape == (1, 3, 16, 6, 8) and torch.equal(img[:, :1], ref_img) and img[:, 1:].abs().max() == 0
    assert torch.equal(lat, torch.randn((1, 3, 16, 6, 8), generator=gen))
    with pytest.raises(RuntimeError, match="Invalid input channels"):
        p.prepare_latents(torch.zeros(1, 5, 1, 6, 8), 1, 16, 9, 1, 48, 64, torch.float32, "cpu", None, None)


def test_rope_tables_match_oracle():
    from orv_b200.models.embeddings import get_3d_rotary_pos_embed, get_resize_crop_region_for_grid, sincos_pos_embed_3d
    cfg = small_cfg(patch_size_t=2, use_rotary_positional_embeddings=True)
    cos, sin = get_3d_rotary_pos_embed(64, None, (3, 4), 2, grid_type="slice", max_size=(3, 4))
    oc, osn = O.pipeline_rope(cfg, 48, 64, 4)
    assert torch.equal(cos, oc) and torch.equal(sin, osn)
    crops = get_resize_crop_region_for_grid((20, 30), 30, 20)
    cos, sin = get_3d_rotary_pos_embed(64, crops, (20, 30), 5)
    oc, osn = O.rope_3d(64, O.resize_crop_region((20, 30), 30, 20), (20, 30), 5)
    assert torch.equal(cos, oc) and torch.equal(sin, osn)
    pos = sincos_pos_embed_3d(1920, 30, 20, 5, 1.875, 1.0)
    ref = O.sincos_3d(1920, (30, 20), 5, 1.875, 1.0).flatten(0, 1)
    assert torch.equal(pos, ref)


def test_shard_range_is_the_reference_partition():
    from orv_b200.dist import shard_range
    for n, w in [(10, 4), (8, 8), (3, 8), (17, 2)]:
        got = [shard_range(n, r, w) for r in range(w)]
        per = n // w
        assert got == [(r * per, (r + 1) * per if r != w - 1 else n) for r in range(w)]
        covered = [i for a, b in got for i in range(a, b)]
        assert covered == list(range(n))


def test_gemm_tile_choice_for_the_config2_block():
    """Host-side tile selection (no GPU needed): the four GEMMs of a CogVideoX-2B block at S = 3226 tokens use the
    CTA-pair kernel with the widths that fill the last wave of the 74 SM pairs (profiles/r01_gemm_notes.md), the QKV
    epilogue only ever gets whole 64-wide heads, and single-tile problems stay on the single-CTA kernel."""
    from orv_b200 import _lib as L
    lib = L.load()
    S, D, FF = 3226, 1920, 7680
    # QKV: 22 tiles of 256 columns + one of 128 per 256-row block = at most 4 x 256 columns per SM pair (6 x 192 uniform)
    assert lib.orvb_gemm_tile_width(S, 3 * D, L.EPI_QKV) == -256
    assert lib.orvb_gemm_tile_remainder(S, 3 * D, L.EPI_QKV) == 128
    assert lib.orvb_gemm_tile_remainder(S, FF, L.EPI_GELU) == 0
    assert lib.orvb_gemm_tile_width(S, D, L.EPI_GATE_RESID) == -176
    assert lib.orvb_gemm_tile_width(S, FF, L.EPI_GELU) == -240
    for n in (64, 128, 1920, 3072, 5760, 9216, 12288):
        w = lib.orvb_gemm_tile_width(S, n, L.EPI_QKV)
        assert w < 0 and (-w) % 64 == 0 and 64 <= -w <= 256
        w = lib.orvb_gemm_tile_width(S, n, L.EPI_BIAS)
        assert w < 0 and (-w) % 16 == 0 and 64 <= -w <= 256
    assert lib.orvb_gemm_tile_width(128, 512, L.EPI_BIAS) > 0
    assert lib.orvb_gemm_tile_width(6, 512, L.EPI_BIAS) in (64, 128, 192, 256)
    assert lib.orvb_gemm_tile_width(0, 512, L.EPI_BIAS) == 0


def _tile_list(lib, m, n, k, epi, in_place=0):
    import ctypes as C
    cap = 8192
    buf = (C.c_int32 * (4 * cap))()
    cnt = lib.orvb_gemm_tile_list(m, n, k, epi, in_place, buf, cap)
    assert 0 <= cnt <= cap, cnt
    return [tuple(buf[4 * i: 4 * i + 4]) for i in range(cnt)]


def test_gemm_tile_lists_cover_the_output_exactly_once():
    """Host-side check of the CTA-pair kernel's tile list (the same g2_tile() the kernel decodes its virtual tile indices
    with, orv_b200/csrc/gemm_common.cuh): for uniform AND mixed lists (n / width full tiles + one narrower tile per 256-row
    block, placed on the SM pairs with one full tile less) every output element belongs to exactly one tile, widths respect
    the epilogue's granularity, and the busiest SM pair of a mixed list has no more columns than with uniform tiles."""
    import numpy as np
    from orv_b200 import _lib as L
    lib = L.load()
    shapes = [(3226, 5760, 1920, L.EPI_QKV, 0), (3226, 5760, 128, L.EPI_BIAS, 0), (6452, 5760, 1920, L.EPI_QKV, 0),
              (4052, 3200, 192, L.EPI_BIAS, 0), (3226, 1920, 1920, L.EPI_GATE_RESID, 1), (3226, 1920, 7680, L.EPI_GATE_RESID, 1),
              (3226, 7680, 1920, L.EPI_GELU, 0), (12876, 1920, 1920, L.EPI_GATE_RESID, 1), (4052, 3072, 3072, L.EPI_GATE_RESID, 1),
              (4052, 9216, 3072, L.EPI_QKV, 0), (300, 11520, 1024, L.EPI_BIAS, 0), (129, 40, 8, L.EPI_BIAS, 0),
              (18944, 1344, 64, L.EPI_BIAS, 0), (19000, 2112, 64, L.EPI_QKV, 0), (2000, 200, 64, L.EPI_GATE_RESID, 1)]
    mixed_seen = 0
    for (m, n, k, epi, inplace) in shapes:
        tiles = _tile_list(lib, m, n, k, epi, inplace)
        assert tiles, (m, n)
        cover = np.zeros(((m + 255) // 256, n), dtype=np.int32)
        load = {}
        widths = set()
        for (c, r0, n0, w) in tiles:
            assert r0 % 256 == 0 and 0 <= r0 < m and n0 < n and 16 <= w <= 256 and w % 16 == 0, (m, n, c, r0, n0, w)
            if epi == L.EPI_QKV:
                assert w % 64 == 0 and n0 % 64 == 0
            cover[r0 // 256, n0:min(n0 + w, n)] += 1
            load[c] = load.get(c, 0) + w
            widths.add(w)
        assert (cover == 1).all(), f"{(m, n)}: {(cover != 1).sum()} output columns not covered exactly once"
        assert len(load) <= 74 and len(widths) <= 2
        rem = lib.orvb_gemm_tile_remainder(m, n, epi) if not inplace else 0
        if len(widths) == 2:
            mixed_seen += 1
            bn = max(widths)
            uniform_rounds = -(-(((m + 255) // 256) * (-(-n // bn))) // 74)
            assert max(load.values()) <= uniform_rounds * bn
            assert rem == min(widths)
        if inplace and n % 64 == 0 and k <= 4096:
            assert max(widths) <= 192  # room for the prefetching gated-residual epilogue
    assert mixed_seen >= 3
    # the headline case: QKV of a 2B block, at most 4 x 256 columns per SM pair (6 x 192 with uniform tiles)
    tiles = _tile_list(lib, 3226, 5760, 1920, L.EPI_QKV)
    load = {}
    for (c, _, _, w) in tiles:
        load[c] = load.get(c, 0) + w
    assert len(tiles) == 13 * 23 and max(load.values()) == 1024
