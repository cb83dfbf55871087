"""CPU: the flat oracle against the golden vectors produced by the reference's OWN code (oracle/make_golden.py
runs /root/reference/orv/models/cogvideox_control.py unmodified on oracle/shim).  fp32 on both sides: the bound is
the north-star's rtol=1e-3 / atol=1e-4 (observed differences are ~1e-6, pure summation-order noise)."""
import hashlib
import os

import pytest
import torch

from oracle import flat_oracle as O
from oracle import make_golden as G

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL, ATOL = 1e-3, 1e-4


def load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


@pytest.mark.parametrize("name", sorted(G.FORWARD_CASES))
def test_forward_matches_reference_golden(name):
    blob = load(name)
    cfg, sd, inp, rope, ofs, t, V = G.build_case(name)
    assert G.digest(sd) == blob["weights_sha256"], "seeded weights differ from the ones the golden was made with"
    assert G.digest({k: v for k, v in inp.items() if v is not None}) == blob["inputs_sha256"]
    with torch.no_grad():
        out = O.forward(sd, cfg, inp["hidden_states"], inp["text"], t, actions=inp["actions"],
                        depths=inp.get("depths"), labels=inp.get("labels"), ofs=ofs, rope=rope, num_views=V)
    torch.testing.assert_close(out, blob["output"], rtol=RTOL, atol=ATOL)


def test_action_embed_mask_matches_reference_golden():
    """Bug-compatibility vector: the reference masks ~10 % of samples even in eval."""
    blob = load("action_embed_mask")
    assert blob["mask_attr"] is True
    cfg = O.default_config(**G.BASE)
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    acts = O.synthetic_inputs(cfg, 16, 3, 6, 8, seed=1, n_actions=8)["actions"]
    torch.manual_seed(5)
    is_mask = torch.rand(16) < 0.1
    assert torch.equal(is_mask, blob["is_mask"])
    emb = O.action_embed(sd, cfg, acts, is_mask)
    torch.testing.assert_close(emb, blob["emb"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("kind,steps,guidance", [("ddim", 2, 1.0), ("dpm", 4, 1.0), ("dpm", 3, 6.0)])
def test_sampler_matches_reference_pipeline_golden(kind, steps, guidance):
    blob = load(f"sampler_{kind}_{steps}steps_g{int(guidance)}")
    cfg = O.default_config(**G.BASE)
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    assert G.digest(sd) == blob["weights_sha256"]
    inp = O.synthetic_inputs(cfg, 1, 3, 6, 8, seed=1, n_actions=8)
    gen = torch.Generator().manual_seed(42)
    with torch.no_grad():
        lat = O.pipeline_call(sd, cfg, kind, blob["moments"], inp["text"], 9, 48, 64, steps, guidance, gen,
                              actions=None if guidance > 1.0 else inp["actions"],
                              negative_prompt_embeds=torch.zeros_like(inp["text"]))
    torch.testing.assert_close(lat, blob["latents"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", sorted(G.PIPELINE_CASES))
def test_pipeline_controls_multiview_match_reference_golden(name):
    """Control latents through the pipeline (moments -> sample with the global RNG -> scale -> channel dup,
    cogvideox_control.py:1331-1364) and the 3-view path, against the reference pipeline's own output."""
    opt = G.PIPELINE_CASES[name]
    blob = load(name)
    cfg = O.default_config(**dict(G.BASE, **opt["over"]))
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    assert G.digest(sd) == blob["weights_sha256"]
    inp, moments, cg = G.pipeline_case_inputs(cfg, opt["controls"], opt["views"])
    assert torch.equal(moments, blob["moments"])
    torch.manual_seed(G.CONTROL_SEED)
    with torch.no_grad():
        lat = O.pipeline_call(sd, cfg, opt["kind"], moments, inp["text"], 9, 48, 64, opt["steps"], 1.0,
                              torch.Generator().manual_seed(42), actions=cg["actions"], num_views=opt["views"],
                              control_moments={k: cg.get(k) for k in ("depths", "labels")})
    assert lat.shape == blob["latents"].shape
    torch.testing.assert_close(lat, blob["latents"], rtol=RTOL, atol=ATOL)


def test_patchify_closed_form_matches_reshape_chain():
    """Integer index maps: bit-exact."""
    for (Fr, C, H, W, pt) in [(3, 4, 6, 8, None), (4, 4, 6, 8, 2), (5, 32, 40, 60, None)]:
        x = torch.arange(Fr * C * H * W, dtype=torch.int64).reshape(1, Fr, C, H, W)
        idx = O.patchify_index_map(Fr, C, H, W, 2, pt)
        if pt is None:
            # conv2d(k=2,s=2) as unfold: K index = c*4 + kh*2 + kw, tokens (f, i, j)
            u = x.reshape(Fr, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(Fr * (H // 2) * (W // 2), C * 4)
        else:
            u = x.permute(0, 1, 3, 4, 2).reshape(1, Fr // pt, pt, H // 2, 2, W // 2, 2, C)
            u = u.permute(0, 1, 3, 5, 7, 2, 4, 6).flatten(4, 7).flatten(1, 3)[0]
        assert torch.equal(x.flatten()[idx], u)


def test_scheduler_timesteps_and_terminal_snr():
    s = O.Scheduler("dpm")
    s.set_timesteps(50)
    assert s.timesteps.tolist() == list(range(999, 0, -20))
    assert float(s.alphas_cumprod[-1]) == 0.0          # zero terminal SNR
    a0 = 1 - 0.00085
    assert abs(float(s.alphas_cumprod[0]) - a0 / (3.0 - 2.0 * a0)) < 1e-12  # snr_shift_scale = 3
    s.set_timesteps(2)
    assert s.timesteps.tolist() == [999, 499]


def test_golden_files_are_small():
    total = sum(os.path.getsize(os.path.join(GOLDEN, f)) for f in os.listdir(GOLDEN))
    assert total < 2 * 1024 * 1024


def test_timestep_embedding_matches_independent_port(tmp_path):
    """SURVEY App. A.4 is a restatement from memory of diffusers' get_timestep_embedding (diffusers is not installable
    here).  The image does ship one independent port of that very function — TVM's relax frontend
    (tvm/relax/frontend/nn/op.py:get_timestep_embedding, same flip_sin_to_cos / downscale_freq_shift / max_period
    arguments, vendored by tilelang).  Build it with TVM's C backend and compare with the oracle's restatement."""
    import numpy as np
    try:
        import tilelang  # noqa: F401  (puts its vendored tvm on sys.path)
        import tvm
        from tvm.relax.frontend import nn
        from tvm.relax.frontend.nn import op, spec
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"tvm not importable: {e}")

    def port(dim, flip, shift, t):
        class M(nn.Module):
            def forward(self, x: nn.Tensor):
                return op.get_timestep_embedding(x, dim, flip_sin_to_cos=flip, downscale_freq_shift=shift)
        try:
            mod, _ = M().export_tvm(spec={"forward": {"x": spec.Tensor([len(t)], "float32")}})
            ex = tvm.relax.build(mod, target="c")
            so = str(tmp_path / f"ts_{dim}_{int(flip)}.so")
            ex.export_library(so)
            vm = tvm.relax.VirtualMachine(tvm.runtime.load_module(so), tvm.cpu())
            mk = getattr(tvm.runtime, "tensor", None) or tvm.runtime.ndarray.array
            return vm["forward"](mk(np.asarray(t, dtype="float32"))).numpy()
        except Exception as e:  # noqa: BLE001
            pytest.skip(f"tvm C backend unavailable: {e}")

    for dim, flip, shift in [(1920, True, 0.0), (512, True, 0.0), (64, False, 1.0)]:
        # small arguments: agreement to fp32 rounding of exp and the product; sampler timesteps (up to 999): sin/cos of ~1e3 in fp32 differ
        # by the argument's own rounding (ulp(1000) = 6e-5) between libm implementations
        for t, tol in [([0.0, 1.0, 2.0, 19.0], 1e-5), ([999.0, 979.0, 499.0, 2.0], 5e-4)]:
            got = port(dim, flip, shift, t)
            ref = O.timestep_sinusoid(torch.tensor(t), dim, flip, shift).numpy()
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() < tol, (dim, flip, shift, t, float(np.abs(got - ref).max()))


def test_sincos_tables_match_the_mae_port_in_transformers():
    """diffusers' sin-cos position tables are the MAE functions (get_2d_sincos_pos_embed_from_grid /
    get_1d_sincos_pos_embed_from_grid); diffusers is absent, but `transformers` ships its own copy of the MAE code
    (models/vit_mae/modeling_vit_mae.py).  That copy pins the oracle's restatement of (a) the frequency ladder and the
    [sin | cos] layout, (b) the w-first meshgrid / first-half-encodes-grid[0] convention of the spatial table, and
    (c) the temporal | spatial split of the 3-D table (App. A.2)."""
    import numpy as np
    from transformers.models.vit_mae.modeling_vit_mae import (get_1d_sincos_pos_embed_from_grid,
                                                              get_2d_sincos_pos_embed)
    pos = torch.tensor([0.0, 1.0, 2.5, 17.0, 299.0])
    for d in (16, 480, 720):
        want = get_1d_sincos_pos_embed_from_grid(d, pos.numpy().astype(np.float64))
        assert np.allclose(O.sincos_1d(d, pos).numpy(), want, rtol=0, atol=1e-12)
    D, g, T = 64, 6, 3
    table = O.sincos_3d(D, (g, g), T, 1.0, 1.0)  # [T, g*g, D], temporal quarter first
    spatial = get_2d_sincos_pos_embed(3 * D // 4, g)  # [g*g, 3D/4]
    temporal = get_1d_sincos_pos_embed_from_grid(D // 4, np.arange(T, dtype=np.float64))
    for t in range(T):
        assert np.allclose(table[t, :, D // 4:].numpy(), spatial, rtol=0, atol=1e-6)
        assert np.allclose(table[t, :, :D // 4].numpy(), np.broadcast_to(temporal[t], (g * g, D // 4)), rtol=0, atol=1e-6)


def test_rope_matches_independent_ports():
    """diffusers' apply_rotary_emb(use_real=True, use_real_unbind_dim=-1) rotates interleaved (even, odd) pairs by the
    angle pos * theta^(-2i/d).  Two independent implementations of that rotation ship in the image: flash-attn's torch
    reference `apply_rotary_emb_torch(interleaved=True)` and the complex-multiplication form in transformers' Llama-4
    (`apply_rotary_emb` with freqs_cis = polar(1, angle)).  Both pin `O.rope_1d` + `O.apply_rope`."""
    from flash_attn.layers.rotary import apply_rotary_emb_torch
    from transformers.models.llama4.modeling_llama4 import apply_rotary_emb as llama4_rope
    g = torch.Generator().manual_seed(0)
    B, H, S, d = 2, 3, 11, 64
    x = torch.randn(B, H, S, d, generator=g)
    pos = torch.arange(S).float() * 1.5
    cos, sin = O.rope_1d(d, pos)                      # [S, d], each frequency repeated for the pair
    got = O.apply_rope(x, cos, sin)
    # the frequency ladder as flash-attn's RotaryEmbedding / Llama define it
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, d, 2).float() / d))
    ang = torch.outer(pos, inv_freq)                  # [S, d/2]
    assert torch.equal(cos[:, 0::2], ang.cos()) and torch.equal(cos[:, 1::2], ang.cos())
    assert torch.equal(sin[:, 0::2], ang.sin()) and torch.equal(sin[:, 1::2], ang.sin())
    want_fa = apply_rotary_emb_torch(x.transpose(1, 2), ang.cos(), ang.sin(), interleaved=True).transpose(1, 2)
    assert torch.allclose(got, want_fa, rtol=0, atol=1e-6)
    freqs_cis = torch.polar(torch.ones_like(ang), ang)[None]  # [1, S, d/2]
    want_l4, _ = llama4_rope(x.transpose(1, 2), x.transpose(1, 2), freqs_cis)
    assert torch.allclose(got, want_l4.transpose(1, 2), rtol=0, atol=1e-5)


def _exact_prediction(sched, t, seed=0, shape=(2, 3, 4, 5)):
    """A denoiser that knows the truth: x_t = alpha x0 + sigma eps and the v-target v = alpha eps - sigma x0
    (Salimans & Ho 2022), in fp32."""
    g = torch.Generator().manual_seed(seed)
    x0, eps = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    a = sched.alphas_cumprod[t]
    alpha, sigma = float(a ** 0.5), float((1 - a) ** 0.5)
    return x0, eps, alpha * x0 + sigma * eps, alpha * eps - sigma * x0


def test_ddim_step_is_the_published_update():
    """DDIM with eta = 0 (Song et al. 2021, eq. 12) must map the exact marginal at t onto the exact marginal at t_prev
    when the network is exact: x_prev = alpha_prev x0 + sigma_prev eps.  Pins the oracle's coefficient algebra
    (v -> x0, the x / x0 recombination, the `final_alpha_cumprod = 1` last step) against the published algorithm,
    independently of my recollection of diffusers' code."""
    s = O.Scheduler("ddim")
    s.set_timesteps(10)
    for t in s.timesteps.tolist():
        x0, eps, x_t, v = _exact_prediction(s, t, seed=t)
        a_t, a_prev, prev = s._alphas(t)
        want = float(a_prev ** 0.5) * x0 + float((1 - a_prev) ** 0.5) * eps
        got = s.step_ddim(v, t, x_t)
        assert torch.allclose(got, want, rtol=0, atol=2e-5), (t, (got - want).abs().max())
        if prev < 0:
            assert torch.allclose(got, x0, rtol=0, atol=2e-5)  # the last step lands on the clean sample


def test_dpm_step_is_sde_dpm_solver_pp():
    """CogVideoXDPMScheduler is SDE-DPM-Solver++ (Lu et al. 2022, data prediction): first order
        x_prev = (sigma_prev / sigma_t) e^{-h} x_t + alpha_prev (1 - e^{-2h}) x0 + sigma_prev sqrt(1 - e^{-2h}) z,
    second order (2M) with D = (1 + 1/(2r)) x0_t - 1/(2r) x0_back, r = h_back / h, h = lambda_prev - lambda_t,
    lambda = log(alpha / sigma).  With an exact denoiser the update must keep the marginal: mean alpha_prev x0 +
    sigma_prev e^{-h} eps, total noise variance sigma_prev^2."""
    import math
    s = O.Scheduler("dpm")
    s.set_timesteps(10)
    ts = s.timesteps.tolist()
    lam = lambda a: -math.inf if float(a) == 0.0 else math.log(float(a ** 0.5) / float((1 - a) ** 0.5))  # noqa: E731
    for i, t in enumerate(ts[:-1]):  # the last step (prev < 0) has sigma_prev = 0; covered by the goldens
        x0, eps, x_t, v = _exact_prediction(s, t, seed=100 + t)
        a_t, a_prev, prev = s._alphas(t)
        al_p, sg_p, sg_t = float(a_prev ** 0.5), float((1 - a_prev) ** 0.5), float((1 - a_t) ** 0.5)
        h = lam(a_prev) - lam(a_t)
        noises = []
        # first-order step (no history)
        got, x0_hat = s.step_dpm(v, None, t, None, x_t, torch.Generator().manual_seed(7), noises)
        assert torch.allclose(x0_hat, x0, rtol=0, atol=2e-5)
        want = (sg_p / sg_t) * math.exp(-h) * x_t + al_p * (1 - math.exp(-2 * h)) * x0 \
            + sg_p * math.sqrt(1 - math.exp(-2 * h)) * noises[-1]
        assert torch.allclose(got, want, rtol=0, atol=5e-5), (t, (got - want).abs().max())
        marginal = al_p * x0 + sg_p * math.exp(-h) * eps + sg_p * math.sqrt(1 - math.exp(-2 * h)) * noises[-1]
        assert torch.allclose(got, marginal, rtol=0, atol=5e-5)
        assert abs((sg_p * math.exp(-h)) ** 2 + sg_p ** 2 * (1 - math.exp(-2 * h)) - sg_p ** 2) < 1e-12
        # second-order step: history = an (inexact) previous data prediction at the previous, noisier timestep
        if i > 0:
            t_back = ts[i - 1]
            old = x0 + 0.1 * eps
            r = (lam(a_t) - lam(s.alphas_cumprod[t_back])) / h
            d = (1 + 1 / (2 * r)) * x0 - (1 / (2 * r)) * old
            noises = []
            got2, _ = s.step_dpm(v, old, t, t_back, x_t, torch.Generator().manual_seed(8), noises)
            want2 = (sg_p / sg_t) * math.exp(-h) * x_t + al_p * (1 - math.exp(-2 * h)) * d \
                + sg_p * math.sqrt(1 - math.exp(-2 * h)) * noises[-1]
            assert torch.allclose(got2, want2, rtol=0, atol=5e-5), (t, (got2 - want2).abs().max())


def test_noise_schedule_follows_the_published_transforms():
    """alphas_cumprod = scaled-linear betas (LDM), SNR divided by snr_shift_scale, then the zero-terminal-SNR rescale
    of Lin et al. 2023 (Algorithm 1: shift sqrt(alpha_bar) so the last value is 0, scale so the first is unchanged)."""
    base = O.Scheduler("ddim", snr_shift_scale=1.0, rescale_betas_zero_snr=False).alphas_cumprod
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    assert torch.allclose(base, torch.cumprod(1 - betas, 0), rtol=0, atol=1e-15)
    shifted = O.Scheduler("ddim", snr_shift_scale=3.0, rescale_betas_zero_snr=False).alphas_cumprod
    snr = lambda a: a / (1 - a)  # noqa: E731
    assert torch.allclose(snr(shifted), snr(base) / 3.0, rtol=1e-12, atol=0)
    final = O.Scheduler("ddim").alphas_cumprod
    s_in, s_out = shifted.sqrt(), final.sqrt()
    assert float(s_out[-1]) == 0.0 and abs(float(s_out[0] - s_in[0])) < 1e-15
    ratio = (s_out[:-1] / (s_in[:-1] - s_in[-1]))
    assert float(ratio.max() - ratio.min()) < 1e-12  # one affine map of sqrt(alpha_bar)
