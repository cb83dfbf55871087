"""B200: the CUDA path against (1) the golden vectors the reference's own code produced and (2) the fp32 oracle at
BASELINE.json's full config-2 size, next to torch-bf16 (the reference's deployed precision) as yardstick."""
import os

import pytest
import torch

from _gates import assert_forward_close, forward_errors
from oracle import flat_oracle as O
from oracle import make_golden as G

pytestmark = pytest.mark.gpu
# Gates of a bf16 forward against an fp32 reference (2-layer goldens).  Yardstick: the reference's own deployed
# precision (eager torch bf16) sits at mean 7.5e-3, max-abs/max 1.2e-2, worst row 5e-2 on these cases.
MEAN_TOL, MAX_TOL, ROW_TOL = 1.5e-2, 3e-2, 8e-2
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _model(cfg, sd):
    from orv_b200 import CogVideoXTransformer3DModelTraj
    m = CogVideoXTransformer3DModelTraj(**cfg)
    m.load_state_dict(sd, strict=False)
    m.action_embed.mask = False
    return m.to("cuda", torch.bfloat16).eval()


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().mean() / b.abs().mean()).item()


@pytest.mark.parametrize("name", sorted(G.FORWARD_CASES))
def test_forward_vs_reference_golden(name):
    """bf16 kernels + bf16-rounded weights/inputs vs the reference's fp32 run (2 layers): mean, max-abs and worst-row
    gates.  Includes the modulate_encoder_hidden_states=False cases (reference :70-99, :404-424)."""
    blob = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    cfg, sd, inp, rope, ofs, t, V = G.build_case(name)
    m = _model(cfg, sd)
    cg = {}
    if inp["actions"] is not None:
        cg["actions"] = inp["actions"].cuda().bfloat16()
    if "depths" in inp:
        cg["depths"], cg["labels"] = inp["depths"].cuda().bfloat16(), inp["labels"].cuda().bfloat16()
    with torch.no_grad():
        out = m(inp["hidden_states"].cuda().bfloat16(), inp["text"].cuda().bfloat16(), cg, t.cuda(),
                ofs=ofs.cuda() if ofs is not None else None,
                image_rotary_emb=(rope[0].cuda(), rope[1].cuda()) if rope else None, return_dict=False,
                num_views=V)[0]
    assert out.shape == blob["output"].shape
    assert_forward_close(out, blob["output"], MEAN_TOL, MAX_TOL, ROW_TOL, name)


def test_eval_time_action_mask_bug_compat():
    """Default module (mask=True): masked samples must produce the output obtained with mask_embed as action
    embedding, and is_action_mask must be the torch.rand(B) < 0.1 draw (reference components.py:66-69)."""
    cfg = O.default_config(**G.BASE)
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    sdr = {k: v.bfloat16().float() for k, v in sd.items()}
    B = 16
    inp = O.synthetic_inputs(cfg, B, 3, 6, 8, seed=1, n_actions=8)
    from orv_b200 import CogVideoXTransformer3DModelTraj
    m = CogVideoXTransformer3DModelTraj(**cfg)
    m.load_state_dict(sd, strict=False)
    m = m.to("cuda", torch.bfloat16).eval()
    assert m.action_embed.mask is True
    t = torch.full((B,), 499)
    torch.manual_seed(5)
    with torch.no_grad():
        out, is_mask, _ = m(inp["hidden_states"].cuda().bfloat16(), inp["text"].cuda().bfloat16(),
                            {"actions": inp["actions"].cuda().bfloat16()}, t.cuda(), return_dict=False)
    torch.manual_seed(5)
    expect = (torch.rand(B, device="cuda") < 0.1).cpu()
    assert torch.equal(is_mask.cpu(), expect) and int(expect.sum()) > 0
    r = lambda x: x.bfloat16().float()  # noqa: E731
    ref = O.forward(sdr, cfg, r(inp["hidden_states"]), r(inp["text"]), t, actions=r(inp["actions"]), action_mask=expect)
    assert _rel(out, ref) < 1.5e-2
    ref_nomask = O.forward(sdr, cfg, r(inp["hidden_states"]), r(inp["text"]), t, actions=r(inp["actions"]))
    assert _rel(out[expect], ref_nomask[expect]) > 5 * _rel(out[expect], ref[expect])


def test_pipeline_ddim_vs_reference_golden():
    """2 DDIM steps through the public pipeline (config-1 plumbing at small size) vs the reference pipeline golden.
    fp32 latents/prompt dtype as in the golden run (CPU-generator fp32 draws), so the RNG streams coincide; this
    exercises the non-fused (torch-op) scheduler path of the pipeline."""
    from orv_b200 import CogVideoXDDIMScheduler, CogVideoXImageToVideoPipelineTraj
    from orv_b200.models.pipeline_control import default_vae_config
    blob = torch.load(os.path.join(GOLDEN, "sampler_ddim_2steps_g1.pt"), weights_only=False)
    cfg = O.default_config(**G.BASE)
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    m = _model(cfg, sd)
    pipe = CogVideoXImageToVideoPipelineTraj(None, None, default_vae_config(), m,
                                             CogVideoXDDIMScheduler(timestep_spacing="trailing"))
    inp = O.synthetic_inputs(cfg, 1, 3, 6, 8, seed=1, n_actions=8)
    out = pipe(image=blob["moments"], prompt="", prompt_embeds=inp["text"].cuda(), height=48,
               width=64, num_frames=9, num_inference_steps=2, guidance_scale=1.0,
               generator=torch.Generator().manual_seed(42), controls_or_guidances={"actions": inp["actions"]},
               output_type="latent", return_dict=False)[0]
    assert out.shape == blob["latents"].shape and out.dtype == torch.float32
    assert_forward_close(out, blob["latents"], 2e-2, 4e-2, 1e-1, "pipeline ddim golden")


@pytest.mark.parametrize("kind,steps,guidance", [("dpm", 4, 1.0), ("ddim", 3, 1.0)])
def test_pipeline_fused_bf16_vs_oracle_pipeline(kind, steps, guidance):
    """The fused bf16 path (CUDA-graph replay + orvb_sampler_step) against the oracle's pipeline run in bf16 with
    the same CPU generator: identical RNG stream, so the only difference is bf16 forward noise (< 3e-2 relative)."""
    from orv_b200 import CogVideoXDDIMScheduler, CogVideoXDPMScheduler, CogVideoXImageToVideoPipelineTraj
    from orv_b200.models.pipeline_control import default_vae_config
    cfg = O.default_config(**G.BASE)
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    m = _model(cfg, sd)
    sch = (CogVideoXDDIMScheduler if kind == "ddim" else CogVideoXDPMScheduler)(timestep_spacing="trailing")
    pipe = CogVideoXImageToVideoPipelineTraj(None, None, default_vae_config(), m, sch)
    inp = O.synthetic_inputs(cfg, 1, 3, 6, 8, seed=1, n_actions=8)
    g = torch.Generator().manual_seed(7)
    moments = torch.randn(1, 32, 1, 6, 8, generator=g).bfloat16()
    out = pipe(image=moments, prompt="", prompt_embeds=inp["text"].cuda().bfloat16(), height=48, width=64,
               num_frames=9, num_inference_steps=steps, guidance_scale=guidance,
               generator=torch.Generator().manual_seed(42), controls_or_guidances={"actions": inp["actions"]},
               output_type="latent", return_dict=False)[0]
    sdb = {k: v.bfloat16() for k, v in sd.items()}
    with torch.no_grad():
        ref = O.pipeline_call(sdb, cfg, kind, moments, inp["text"].bfloat16(), 9, 48, 64, steps, guidance,
                              torch.Generator().manual_seed(42), actions=inp["actions"].bfloat16())
    assert out.dtype == torch.bfloat16
    assert_forward_close(out, ref, 3e-2, 6e-2, 1.5e-1, f"fused {kind} pipeline vs oracle pipeline (both bf16)")


def _pipe(cfg, sd, kind="dpm"):
    from orv_b200 import CogVideoXDDIMScheduler, CogVideoXDPMScheduler, CogVideoXImageToVideoPipelineTraj
    from orv_b200.models.pipeline_control import default_vae_config
    sch = (CogVideoXDDIMScheduler if kind == "ddim" else CogVideoXDPMScheduler)(timestep_spacing="trailing")
    return CogVideoXImageToVideoPipelineTraj(None, None, default_vae_config(), _model(cfg, sd), sch)


@pytest.mark.parametrize("fused", [False, True])
def test_pipeline_cfg_guidance6_vs_reference_golden(fused):
    """Classifier-free guidance through `pipe()` (guidance 6: negative | positive prompt batch, fp32 CFG combine,
    cogvideox_control.py:1409-1437) against the reference pipeline's own 3-step DPM run.  fused=False: fp32 latents as
    in the golden run (torch-op scheduler path); fused=True: bf16 latents, `orvb_sampler_step` doing the CFG combine
    + DPM update, CUDA-graph replay of the batch-2 forward.  The reference cannot run CFG together with actions (P5),
    so the golden has none."""
    blob = torch.load(os.path.join(GOLDEN, "sampler_dpm_3steps_g6.pt"), weights_only=False)
    cfg = O.default_config(**G.BASE)
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    pipe = _pipe(cfg, sd)
    inp = O.synthetic_inputs(cfg, 1, 3, 6, 8, seed=1, n_actions=8)
    dt = torch.bfloat16 if fused else torch.float32
    text = inp["text"].cuda().to(dt)
    out = pipe(image=blob["moments"].to(dt), prompt=None, prompt_embeds=text, negative_prompt_embeds=torch.zeros_like(text),
               height=48, width=64, num_frames=9, num_inference_steps=3, guidance_scale=6.0,
               generator=torch.Generator().manual_seed(42), controls_or_guidances={}, output_type="latent",
               return_dict=False)[0]
    assert out.shape == blob["latents"].shape and out.dtype == dt
    if fused:  # bf16 noise draws differ from the golden's fp32 ones: compare with the oracle pipeline in bf16 instead
        sdb = {k: v.bfloat16() for k, v in sd.items()}
        with torch.no_grad():
            ref = O.pipeline_call(sdb, cfg, "dpm", blob["moments"].bfloat16(), inp["text"].bfloat16(), 9, 48, 64, 3, 6.0,
                                  torch.Generator().manual_seed(42), actions=None,
                                  negative_prompt_embeds=torch.zeros_like(inp["text"]).bfloat16())
        assert_forward_close(out, ref, 4e-2, 8e-2, 2e-1, "CFG g=6 fused bf16 vs oracle pipeline bf16")
    else:
        # guidance 6 amplifies the bf16 forward noise of (cond - uncond) sixfold
        assert_forward_close(out, blob["latents"], 4e-2, 8e-2, 2e-1, "CFG g=6 vs reference golden")


@pytest.mark.parametrize("name", sorted(G.PIPELINE_CASES))
def test_pipeline_controls_multiview_vs_reference_golden(name):
    """Depth / label VAE moments through `pipe()` — sampled with the global RNG, scaled, duplicated on the channel axis
    (cogvideox_control.py:1331-1364) — and the 3-view path (num_views=3, config 5) against the reference pipeline's
    own output (fp32 golden, CPU RNG streams reproduced exactly: the draws happen on the host in both)."""
    opt = G.PIPELINE_CASES[name]
    blob = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    cfg = O.default_config(**dict(G.BASE, **opt["over"]))
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    pipe = _pipe(cfg, sd, opt["kind"])
    inp, moments, cg = G.pipeline_case_inputs(cfg, opt["controls"], opt["views"])
    torch.manual_seed(G.CONTROL_SEED)
    out = pipe(image=moments, prompt="", prompt_embeds=inp["text"].cuda(), height=48, width=64, num_frames=9,
               num_inference_steps=opt["steps"], guidance_scale=1.0, generator=torch.Generator().manual_seed(42),
               controls_or_guidances=cg, output_type="latent", return_dict=False, num_views=opt["views"])[0]
    assert out.shape == blob["latents"].shape
    assert_forward_close(out, blob["latents"], 2e-2, 4e-2, 1e-1, name)


def test_pipeline_static_cache_and_schedule_are_bit_identical(monkeypatch):
    """The step-invariant cache (text projection + control embeddings computed by the first iteration only) and the
    per-clip modulation schedule must not change a single bit of a fused bf16 run with controls."""
    opt = G.PIPELINE_CASES["sampler_dpm_3steps_controls"]
    cfg = O.default_config(**dict(G.BASE, **opt["over"]))
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    inp, moments, cg = G.pipeline_case_inputs(cfg, True, 1)
    outs = []
    for static, sched in (("1", "1"), ("0", "1"), ("0", "0")):
        monkeypatch.setenv("ORVB_STATIC_CACHE", static)
        monkeypatch.setenv("ORVB_MOD_SCHEDULE", sched)
        pipe = _pipe(cfg, sd)
        for rep in range(2):  # the second call replays captured graphs
            torch.manual_seed(G.CONTROL_SEED)
            outs.append(pipe(image=moments.bfloat16(), prompt="", prompt_embeds=inp["text"].cuda().bfloat16(), height=48,
                             width=64, num_frames=9, num_inference_steps=4, guidance_scale=1.0,
                             generator=torch.Generator().manual_seed(42), controls_or_guidances=cg, output_type="latent",
                             return_dict=False)[0].clone())
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


def test_graph_replay_alternating_two_shapes():
    """the AdaLN job / site tables live in the workspace they describe, so a captured graph of shape A must
    stay correct after shape B ran (per-step modulation path, direct model calls, graph replay on)."""
    cfg = O.default_config(**G.BASE)
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    m = _model(cfg, sd)
    assert m.use_cuda_graph
    cases = {}
    for B in (1, 3):
        inp = O.synthetic_inputs(cfg, B, 3, 6, 8, seed=B, n_actions=8)
        cases[B] = (inp["hidden_states"].cuda().bfloat16(), inp["text"].cuda().bfloat16(),
                    {"actions": inp["actions"].cuda().bfloat16()})
    outs = {B: [] for B in cases}
    with torch.no_grad():
        for rnd in range(4):  # round 0 eager, round 1 captures, rounds 2-3 replay — always alternating A, B, A, B
            for B, (hs, text, cg) in cases.items():
                for t in (801.0, 301.0):
                    outs[B].append(m(hs, text, cg, torch.full((B,), t, device="cuda"), return_dict=False)[0].clone())
    for B in cases:
        for i in range(2, len(outs[B])):
            assert torch.equal(outs[B][i], outs[B][i % 2]), (B, i)
        assert not torch.equal(outs[B][0], outs[B][1])


def test_full_size_config2_forward_vs_fp32_oracle():
    """BASELINE config 2 geometry (S=3226, D=1920, 30 heads) with 3 layers: err(ours) <= 1.25 x err(torch-bf16)."""
    import time
    t0 = time.time()
    cfg = O.default_config(num_attention_heads=30, attention_head_dim=64, in_channels=32, out_channels=16,
                           num_layers=3, sample_width=60, sample_height=40, sample_frames=17,
                           modulate_encoder_hidden_states=True, text_embed_dim=4096, max_text_seq_length=226)
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.02)
    sdr = {k: v.bfloat16().float() for k, v in sd.items()}
    inp = O.synthetic_inputs(cfg, 1, 5, 40, 60, seed=1)
    r = lambda x: x.bfloat16().float()  # noqa: E731
    t = torch.tensor([499])
    with torch.no_grad():
        print(f"[t={time.time() - t0:.1f}s] weights/inputs ready", flush=True)
        ref = O.forward(sdr, cfg, r(inp["hidden_states"]), r(inp["text"]), t, actions=r(inp["actions"]))
        print(f"[t={time.time() - t0:.1f}s] fp32 CPU oracle done", flush=True)
        sdb = {k: v.cuda().bfloat16() for k, v in sd.items()}
        # torch-bf16 on the GPU = the reference's deployed arithmetic (eager bf16 ops); the oracle's host-built
        # tables are moved to the device by running it under a cuda default device
        with torch.device("cuda"):
            tb = O.forward(sdb, cfg, inp["hidden_states"].cuda().bfloat16(), inp["text"].cuda().bfloat16(), t.cuda(),
                           actions=inp["actions"].cuda().bfloat16())
    print(f"[t={time.time() - t0:.1f}s] torch-bf16 GPU done", flush=True)
    m = _model(cfg, sd)
    print(f"[t={time.time() - t0:.1f}s] model built", flush=True)
    with torch.no_grad():
        out = m(inp["hidden_states"].cuda().bfloat16(), inp["text"].cuda().bfloat16(),
                {"actions": inp["actions"].cuda().bfloat16()}, t.cuda(), return_dict=False)[0]
    torch.cuda.synchronize()
    print(f"[t={time.time() - t0:.1f}s] CUDA forward done", flush=True)
    e_ours, e_torch = _rel(out, ref), _rel(tb, ref)
    print(f"full-size rel err: ours {e_ours:.3e}  torch-bf16 {e_torch:.3e}")
    assert e_ours < 1.25 * e_torch + 1e-3
    _, mx_t, row_t = forward_errors(tb, ref)
    # max-abs and worst-row gates, absolute and against the reference's deployed precision on the same inputs
    _, mx, row = assert_forward_close(out, ref, 2e-2, 5e-2, 1e-1, "config 2 full size (3 layers)")
    assert mx < 1.5 * mx_t + 5e-3 and row < 1.5 * row_t + 5e-3, (mx, mx_t, row, row_t)


def _full_geometry_case(cfg_over, B, Fr, H, W, *, controls=False, rope=False, ofs=None, views=1, n_actions=16):
    """Runs the CUDA forward and the fp32 oracle (on the GPU, fp32 matmuls) at a BASELINE.json geometry with a reduced
    layer count; returns (rel err ours, rel err torch-bf16) against the fp32 oracle."""
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = O.default_config(**cfg_over)
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.02)
    inp = O.synthetic_inputs(cfg, B, Fr * views, H, W, seed=1, with_controls=controls, n_actions=n_actions)
    t = torch.full((B,), 499)
    rp = O.pipeline_rope(cfg, H * 8, W * 8, Fr) if rope else None
    ofs_t = torch.tensor([ofs]) if ofs is not None else None
    kw32, kwb = {}, {}
    for k in ("actions", "depths", "labels"):
        if k in inp and (controls or k == "actions"):
            kw32[k] = inp[k].bfloat16().float().cuda()
            kwb[k] = inp[k].cuda().bfloat16()
    with torch.no_grad(), torch.device("cuda"):
        sdr = {k: v.bfloat16().float().cuda() for k, v in sd.items()}
        rope_c = (rp[0].cuda(), rp[1].cuda()) if rp else None
        ref = O.forward(sdr, cfg, inp["hidden_states"].bfloat16().float().cuda(), inp["text"].bfloat16().float().cuda(),
                        t.cuda(), ofs=ofs_t.cuda() if ofs_t is not None else None, rope=rope_c, num_views=views, **kw32)
        del sdr
        sdb = {k: v.cuda().bfloat16() for k, v in sd.items()}
        tb = O.forward(sdb, cfg, inp["hidden_states"].cuda().bfloat16(), inp["text"].cuda().bfloat16(), t.cuda(),
                       ofs=ofs_t.cuda() if ofs_t is not None else None, rope=rope_c, num_views=views, **kwb)
        del sdb
    m = _model(cfg, sd)
    with torch.no_grad():
        out = m(inp["hidden_states"].cuda().bfloat16(), inp["text"].cuda().bfloat16(), kwb, t.cuda(),
                ofs=ofs_t.cuda() if ofs_t is not None else None, image_rotary_emb=rope_c, return_dict=False,
                num_views=views)[0]
    torch.cuda.synchronize()
    assert out.shape == ref.shape and torch.isfinite(out.float()).all()
    _, mx_t, row_t = forward_errors(tb, ref)
    _, mx, row = assert_forward_close(out, ref, 2e-2, 5e-2, 1e-1, "full geometry")
    assert mx < 1.5 * mx_t + 5e-3 and row < 1.5 * row_t + 5e-3, (mx, mx_t, row, row_t)
    return _rel(out, ref), _rel(tb, ref)


def test_full_size_config3_controls():
    """BASELINE config 3: 2B dims + depth/label condition latents, S = 3226, 2 layers."""
    e, eb = _full_geometry_case(dict(num_attention_heads=30, attention_head_dim=64, in_channels=32, out_channels=16,
                                     num_layers=2, sample_width=60, sample_height=40, sample_frames=17,
                                     modulate_encoder_hidden_states=True, text_embed_dim=4096, max_text_seq_length=226,
                                     visual_guidance=True, num_control_blocks=2), 1, 5, 40, 60, controls=True)
    print(f"config 3 geometry: ours {e:.3e} torch-bf16 {eb:.3e}")
    assert e < 1.25 * eb + 1e-3 and e < 2e-2


def test_full_size_config4_5b_cfg_pair():
    """BASELINE config 4: CogVideoX1.5-5B dims (D=3072, 48 heads, FF 12288, p_t=2, RoPE, ofs), batch 2 (CFG pair),
    6 latent frames -> S = 226 + 1800, 2 layers."""
    e, eb = _full_geometry_case(dict(num_attention_heads=48, attention_head_dim=64, in_channels=32, out_channels=16,
                                     num_layers=2, sample_width=60, sample_height=40, sample_frames=21,
                                     modulate_encoder_hidden_states=True, text_embed_dim=4096, max_text_seq_length=226,
                                     patch_size_t=2, use_rotary_positional_embeddings=True, ofs_embed_dim=512,
                                     patch_bias=False), 2, 6, 40, 60, rope=True, ofs=2.0, n_actions=20)
    print(f"config 4 geometry: ours {e:.3e} torch-bf16 {eb:.3e}")
    assert e < 1.25 * eb + 1e-3 and e < 2e-2


def test_full_size_config5_multiview():
    """BASELINE config 5: 2B multiview, 3 views x 5 latent frames of 32x48 latents (S_v = 1920 per view) + conditions,
    2 temporal + 2 multiview blocks."""
    e, eb = _full_geometry_case(dict(num_attention_heads=30, attention_head_dim=64, in_channels=32, out_channels=16,
                                     num_layers=2, sample_width=48, sample_height=32, sample_frames=17,
                                     modulate_encoder_hidden_states=True, text_embed_dim=4096, max_text_seq_length=226,
                                     visual_guidance=True, num_control_blocks=2, multiview=True, max_n_view=3),
                                1, 5, 32, 48, controls=True, views=3)
    print(f"config 5 geometry: ours {e:.3e} torch-bf16 {eb:.3e}")
    assert e < 1.25 * eb + 1e-3 and e < 2e-2


def test_full_size_config1_ddim_two_steps_pipeline():
    """BASELINE config 1 (plumbing): the full CogVideoX-2B geometry, 1 clip 17x320x480, 2 DDIM-trailing steps through
    the public pipeline on the GPU against the CPU oracle's pipeline run in bf16 with the same CPU generator (identical
    RNG stream, so the only difference is bf16 forward noise)."""
    from orv_b200 import CogVideoXDDIMScheduler, CogVideoXImageToVideoPipelineTraj
    from orv_b200.models.pipeline_control import default_vae_config
    cfg = O.default_config(num_attention_heads=30, attention_head_dim=64, in_channels=32, out_channels=16, num_layers=30,
                           sample_width=60, sample_height=40, sample_frames=17, modulate_encoder_hidden_states=True,
                           text_embed_dim=4096, max_text_seq_length=226)
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.02)
    m = _model(cfg, sd)
    pipe = CogVideoXImageToVideoPipelineTraj(None, None, default_vae_config(), m,
                                             CogVideoXDDIMScheduler(timestep_spacing="trailing"))
    inp = O.synthetic_inputs(cfg, 1, 5, 40, 60, seed=1)
    moments = torch.randn(1, 32, 1, 40, 60, generator=torch.Generator().manual_seed(7)).bfloat16()
    out = pipe(image=moments, prompt="", prompt_embeds=inp["text"].cuda().bfloat16(), height=320, width=480,
               num_frames=17, num_inference_steps=2, guidance_scale=1.0, generator=torch.Generator().manual_seed(42),
               controls_or_guidances={"actions": inp["actions"]}, output_type="latent", return_dict=False)[0]
    sdb = {k: v.bfloat16() for k, v in sd.items()}
    del sd
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        ref = O.pipeline_call(sdb, cfg, "ddim", moments, inp["text"].bfloat16(), 17, 320, 480, 2, 1.0,
                              torch.Generator().manual_seed(42), actions=inp["actions"].bfloat16())
    assert out.shape == ref.shape == (1, 5, 16, 40, 60) and out.dtype == torch.bfloat16
    e = _rel(out, ref)
    print(f"config 1 (2 DDIM steps, 30 layers): rel diff vs CPU bf16 oracle pipeline {e:.3e}")
    assert e < 3e-2
    assert_forward_close(out, ref, 3e-2, 1e-1, 2e-1, "config 1 pipeline (both sides bf16, 30 layers)")
