"""Generates tests/golden/*.pt by running the REFERENCE's own code (orv/models/cogvideox_control.py and
components.py, imported unmodified from /root/reference on top of oracle/shim) on seeded synthetic weights/inputs.

    python -m oracle.make_golden            (build container only: needs /root/reference)

Each file stores the case description (config overrides, seeds, shapes), a SHA-256 of the regenerated weights and
inputs (so RNG drift is detected instead of silently mis-comparing) and the reference outputs in fp32.  The
weights/inputs themselves are regenerated from the seeds by `oracle.flat_oracle.synthetic_*` in the tests.
"""
from __future__ import annotations

import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import flat_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

BASE = dict(num_attention_heads=2, attention_head_dim=64, in_channels=32, out_channels=16, num_layers=2,
            sample_width=8, sample_height=6, sample_frames=9, modulate_encoder_hidden_states=True, text_embed_dim=32,
            max_text_seq_length=10, time_embed_dim=64, num_control_blocks=2)

# name -> (config overrides, batch, latent frames, h, w, options)
FORWARD_CASES = {
    "fwd_actions": (dict(), 1, 3, 6, 8, dict(n_actions=8)),
    "fwd_noactions_b2": (dict(), 2, 3, 6, 8, dict(n_actions=0)),
    "fwd_controls_b2": (dict(visual_guidance=True), 2, 3, 6, 8, dict(n_actions=8, controls=True)),
    "fwd_rope_pt2_ofs": (dict(patch_size_t=2, use_rotary_positional_embeddings=True, ofs_embed_dim=64, patch_bias=False),
                         2, 4, 6, 8, dict(n_actions=12, rope=True, ofs=2.0)),
    "fwd_multiview_v3": (dict(multiview=True, max_n_view=3, visual_guidance=True), 1, 2, 6, 8,
                         dict(n_actions=4, controls=True, views=3)),
    "fwd_othergeom": (dict(), 1, 2, 4, 10, dict(n_actions=4)),  # geometry != sample_* -> pos-emb recomputed on the fly
    # modulate_encoder_hidden_states=False (from-scratch 1.4B configs, reference :70-99 / :404-424), with and without actions
    "fwd_nomodtext_b2": (dict(modulate_encoder_hidden_states=False), 2, 3, 6, 8, dict(n_actions=8)),
    "fwd_nomodtext_noact": (dict(modulate_encoder_hidden_states=False), 1, 3, 6, 8, dict(n_actions=0)),
}


def digest(tensors) -> str:
    h = hashlib.sha256()
    for k in sorted(tensors):
        t = tensors[k]
        h.update(k.encode())
        h.update(t.detach().contiguous().float().numpy().tobytes())
    return h.hexdigest()


def build_case(name):
    over, B, Fr, H, W, opt = FORWARD_CASES[name]
    cfg = O.default_config(**dict(BASE, **over))
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    V = opt.get("views", 1)
    inp = O.synthetic_inputs(cfg, B, Fr * V, H, W, seed=1, with_controls=opt.get("controls", False),
                             n_actions=max(opt.get("n_actions", 8), 1))
    if opt.get("n_actions", 8) == 0:
        inp["actions"] = None
    rope = O.pipeline_rope(cfg, H * 8, W * 8, Fr) if opt.get("rope") else None
    ofs = torch.tensor([opt["ofs"]]) if "ofs" in opt else None
    t = torch.full((B,), 499, dtype=torch.int64)
    return cfg, sd, inp, rope, ofs, t, V


def run_reference_forward(mod, cfg, sd, inp, rope, ofs, t, V):
    model = mod.CogVideoXTransformer3DModelTraj(**cfg)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [k for k in missing if "action_recon" not in k], (missing, unexpected)
    model.eval()
    model.action_embed.mask = False
    cg = {}
    if inp["actions"] is not None:
        cg["actions"] = inp["actions"]
    if "depths" in inp:
        cg["depths"], cg["labels"] = inp["depths"], inp["labels"]
    with torch.no_grad():
        out, is_mask, recon = model(inp["hidden_states"], inp["text"], cg, t, ofs=ofs, image_rotary_emb=rope,
                                    return_dict=False, num_views=V)
    return out


# extra pipeline cases (name -> options): control latents through `pipe()` (moments sampled WITHOUT the generator,
# cogvideox_control.py:1331-1364, hence the global seed) and the 3-view path (config 5)
PIPELINE_CASES = {
    "sampler_dpm_3steps_controls": dict(kind="dpm", steps=3, over=dict(visual_guidance=True), controls=True, views=1),
    "sampler_dpm_3steps_multiview": dict(kind="dpm", steps=3, over=dict(multiview=True, max_n_view=3, visual_guidance=True),
                                         controls=True, views=3),
}
CONTROL_SEED = 11


def pipeline_case_inputs(cfg, controls: bool, views: int):
    """Seeded inputs of a pipeline case: text/actions from synthetic_inputs, first-frame moments per view and (optionally)
    depth / label VAE moments [B, 32, V*F, h, w] with a non-trivial log-variance."""
    inp = O.synthetic_inputs(cfg, 1, 3, 6, 8, seed=1, n_actions=8)
    g = torch.Generator().manual_seed(7)
    moments = torch.randn(1, 32, views, 6, 8, generator=g)  # first-frame VAE moments [B, 2*16, V*1, h, w]
    moments[:, 16:] = moments[:, 16:] * 0.5 - 3.0
    cg = {"actions": inp["actions"]}
    if controls:
        for key in ("depths", "labels"):
            m = torch.randn(1, 32, views * 3, 6, 8, generator=g)
            m[:, 16:] = m[:, 16:] * 0.5 - 2.0
            cg[key] = m
    return inp, moments, cg


def sampler_case(mod, kind, steps, guidance, seed=42, over=None, controls=False, views=1):
    """The reference pipeline's __call__ (latent in / latent out) on the small model, fp32 on the CPU."""
    from diffusers.models.autoencoders.autoencoder_kl_cogvideox import AutoencoderKLCogVideoX
    from diffusers.schedulers.scheduling_ddim_cogvideox import CogVideoXDDIMScheduler
    from diffusers.schedulers.scheduling_dpm_cogvideox import CogVideoXDPMScheduler
    cfg = O.default_config(**dict(BASE, **(over or {})))
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    model = mod.CogVideoXTransformer3DModelTraj(**cfg)
    model.load_state_dict(sd, strict=False)
    model.eval()
    model.action_embed.mask = False
    sched_cls = CogVideoXDDIMScheduler if kind == "ddim" else CogVideoXDPMScheduler
    sched = sched_cls(prediction_type="v_prediction", rescale_betas_zero_snr=True, snr_shift_scale=3.0,
                      timestep_spacing="trailing", clip_sample=False)
    pipe = mod.CogVideoXImageToVideoPipelineTraj(None, None, AutoencoderKLCogVideoX(), model, sched)
    if controls or views > 1:
        inp, moments, cg = pipeline_case_inputs(cfg, controls, views)
    else:  # the round-1 cases, byte for byte
        inp = O.synthetic_inputs(cfg, 1, 3, 6, 8, seed=1, n_actions=8)
        g = torch.Generator().manual_seed(7)
        moments = torch.randn(1, 32, 1, 6, 8, generator=g)  # first-frame VAE moments [B, 2*16, F=1, h, w]
        moments[:, 16:] = moments[:, 16:] * 0.5 - 3.0
        cg = {} if guidance > 1.0 else {"actions": inp["actions"]}  # the reference cannot run CFG with controls (P5)
    gen = torch.Generator().manual_seed(seed)
    kw = dict(image=moments, prompt="", prompt_embeds=inp["text"], height=48, width=64, num_frames=9,
              num_inference_steps=steps, guidance_scale=guidance, generator=gen, controls_or_guidances=cg,
              output_type="latent", return_dict=False, num_views=views)
    if guidance > 1.0:
        # positional check_inputs quirk (cogvideox_control.py:1261-1270): with CFG the embeds pair only passes
        # validation when `prompt` is None
        kw["prompt"] = None
        kw["negative_prompt_embeds"] = torch.zeros_like(inp["text"])
    torch.manual_seed(CONTROL_SEED)  # the control-latent draws use the global RNG
    out = pipe(**kw)[0]
    return cfg, sd, inp, moments, out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    mod = ref_loader.load()
    torch.manual_seed(1234)
    for name in FORWARD_CASES:
        cfg, sd, inp, rope, ofs, t, V = build_case(name)
        out = run_reference_forward(mod, cfg, sd, inp, rope, ofs, t, V)
        blob = {"name": name, "weights_sha256": digest(sd),
                "inputs_sha256": digest({k: v for k, v in inp.items() if v is not None}), "output": out.float()}
        torch.save(blob, os.path.join(GOLDEN_DIR, name + ".pt"))
        print(f"{name}: out {tuple(out.shape)} mean|x| {out.abs().mean():.4f}")
    # ActionEmbed with the reference's eval-time random mask: seeded global RNG
    cfg = O.default_config(**BASE)
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.05)
    model = mod.CogVideoXTransformer3DModelTraj(**cfg)
    model.load_state_dict(sd, strict=False)
    model.eval()
    acts = O.synthetic_inputs(cfg, 16, 3, 6, 8, seed=1, n_actions=8)["actions"]
    torch.manual_seed(5)
    with torch.no_grad():
        emb, is_mask = model.action_embed(torch.cat([acts.new_zeros(16, 3, 7), acts], dim=1))
    torch.save({"name": "action_embed_mask", "mask_attr": bool(model.action_embed.mask), "emb": emb.float(),
                "is_mask": is_mask, "weights_sha256": digest(sd)}, os.path.join(GOLDEN_DIR, "action_embed_mask.pt"))
    print("action_embed_mask: mask attr", model.action_embed.mask, "masked", int(is_mask.sum()))
    for kind, steps, guidance in (("ddim", 2, 1.0), ("dpm", 4, 1.0), ("dpm", 3, 6.0)):
        cfg, sd, inp, moments, out = sampler_case(mod, kind, steps, guidance)
        name = f"sampler_{kind}_{steps}steps_g{int(guidance)}"
        torch.save({"name": name, "weights_sha256": digest(sd), "moments": moments, "latents": out.float()},
                   os.path.join(GOLDEN_DIR, name + ".pt"))
        print(f"{name}: latents {tuple(out.shape)} mean|x| {out.abs().mean():.4f}")
    for name, opt in PIPELINE_CASES.items():
        cfg, sd, inp, moments, out = sampler_case(mod, opt["kind"], opt["steps"], 1.0, over=opt["over"],
                                                  controls=opt["controls"], views=opt["views"])
        torch.save({"name": name, "weights_sha256": digest(sd), "moments": moments, "latents": out.float()},
                   os.path.join(GOLDEN_DIR, name + ".pt"))
        print(f"{name}: latents {tuple(out.shape)} mean|x| {out.abs().mean():.4f}")


if __name__ == "__main__":
    main()
