"""CPU oracle for ORV's denoising hot path — TEST INFRASTRUCTURE, never the product path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import this
module; `orv_b200/` must not (tests/test_layout.py enforces it).

What it is: a flat, single-file torch restatement (fp32 by default) of
  * `CogVideoXTransformer3DModelTraj.forward`            reference orv/models/cogvideox_control.py:715-948
  * `CogVideoXLayerNormZero` / `AdaLayerNorm` (ORV)       :41-197
  * `CogVideoXAttnProcessor2_0.__call__` (ORV)            :200-270
  * `MVBlock` / `CogVideoXBlock`                          :273-445
  * `ActionEmbed.forward`                                 reference orv/models/components.py:47-71
  * the sampler loop of `CogVideoXImageToVideoPipelineTraj.__call__`  :1402-1473
plus the un-vendored diffusers (>=0.31.2, configs stamped 0.32.0.dev0 — requirements.txt:21) primitives those
call: CogVideoXPatchEmbed, get_3d_sincos_pos_embed, get_3d_rotary_pos_embed, apply_rotary_emb, Timesteps /
TimestepEmbedding, FeedForward(gelu-approximate), CogVideoXDDIMScheduler / CogVideoXDPMScheduler,
DiagonalGaussianDistribution — restated from the published diffusers 0.32 algorithms.

Pinning status: PARITY PARTLY PINNED.  The reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md §4), and diffusers itself is absent from /root/reference and from this image.  What pins this file:
  (1) `oracle/make_golden.py` imports the reference's OWN modules (cogvideox_control.py, components.py, run
      unmodified on top of `oracle/shim/`, a minimal stand-in for the diffusers symbols they import) and
      stores their outputs under tests/golden/; tests/test_oracle.py checks this file against those vectors.
      That pins every line of ORV-authored arithmetic.
  (2) The diffusers primitives are pinned only against my restatement of their published algorithm
      (the shim and this file are written independently of each other but by the same author), so for those
      the oracle is "parity unpinned" in the task's sense until a diffusers wheel can be imported.

State-dict keys are the reference module's.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------------------
# configuration
# ------------------------------------------------------------------------------------------------------------
def default_config(**over) -> dict:
    """Constructor defaults of CogVideoXTransformer3DModelTraj (cogvideox_control.py:453-494)."""
    cfg = dict(
        num_attention_heads=30, attention_head_dim=64, in_channels=16, out_channels=16, flip_sin_to_cos=True,
        freq_shift=0, time_embed_dim=512, ofs_embed_dim=None, text_embed_dim=4096, num_layers=30, dropout=0.0,
        attention_bias=True, sample_width=90, sample_height=60, sample_frames=49, patch_size=2, patch_size_t=None,
        temporal_compression_ratio=4, max_text_seq_length=226, activation_fn="gelu-approximate",
        timestep_activation_fn="silu", norm_elementwise_affine=True, norm_eps=1e-5,
        spatial_interpolation_scale=1.875, temporal_interpolation_scale=1.0, use_rotary_positional_embeddings=False,
        use_learned_positional_embeddings=False, patch_bias=True, loaded_pretrained_model_name_or_path=None,
        modulate_encoder_hidden_states=False, num_control_blocks=12, recon_action=False, visual_guidance=False,
        num_control_keys=2, multiview=False, max_n_view=3, from_t2v=False,
    )
    cfg.update(over)
    return cfg


# ------------------------------------------------------------------------------------------------------------
# diffusers primitives
# ------------------------------------------------------------------------------------------------------------
def sincos_1d(embed_dim: int, pos: Tensor) -> Tensor:
    """diffusers get_1d_sincos_pos_embed_from_grid (output_type='pt'): [sin | cos], omega in float64."""
    omega = torch.arange(embed_dim // 2, dtype=torch.float64)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = torch.outer(pos.reshape(-1).to(torch.float64), omega)
    return torch.cat([torch.sin(out), torch.cos(out)], dim=1)


def sincos_3d(embed_dim: int, spatial_size: Tuple[int, int], temporal_size: int, spatial_scale: float,
              temporal_scale: float) -> Tensor:
    """diffusers get_3d_sincos_pos_embed -> [T, H*W, D]; spatial_size = (W, H); temporal part first (App. A.2)."""
    w, h = spatial_size
    d_sp, d_t = 3 * embed_dim // 4, embed_dim // 4
    grid_h = torch.arange(h, dtype=torch.float32) / spatial_scale
    grid_w = torch.arange(w, dtype=torch.float32) / spatial_scale
    gw, gh = torch.meshgrid(grid_w, grid_h, indexing="xy")  # each [h, w]; gw varies along axis 1
    emb_a = sincos_1d(d_sp // 2, gw)  # diffusers calls this "emb_h" but feeds grid[0] = the w coordinate
    emb_b = sincos_1d(d_sp // 2, gh)
    pos_sp = torch.cat([emb_a, emb_b], dim=1)  # [h*w, d_sp]
    grid_t = torch.arange(temporal_size, dtype=torch.float32) / temporal_scale
    pos_t = sincos_1d(d_t, grid_t)  # [T, d_t]
    pos_sp = pos_sp[None].expand(temporal_size, -1, -1)
    pos_t = pos_t[:, None].expand(-1, h * w, -1)
    return torch.cat([pos_t, pos_sp], dim=-1).float()


def joint_pos_embedding(cfg: dict, frames_lat: int, height: int, width: int) -> Tensor:
    """CogVideoXPatchEmbed._get_positional_embeddings for latent [frames_lat, height, width] -> [St + N, D]."""
    p = cfg["patch_size"]
    D = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    pos = sincos_3d(D, (width // p, height // p), frames_lat, cfg["spatial_interpolation_scale"],
                    cfg["temporal_interpolation_scale"]).flatten(0, 1)
    joint = torch.zeros(cfg["max_text_seq_length"] + pos.shape[0], D)
    joint[cfg["max_text_seq_length"]:] = pos
    return joint


_TABLES: Dict[tuple, Tensor] = {}


def _cached_table(kind: str, cfg: dict, geom: tuple, like: Tensor) -> Tensor:
    """Position tables are registered buffers in the reference (computed at construction, moved with the module);
    the oracle builds them on the CPU as before and keeps one copy per (geometry, device, dtype)."""
    key = (kind, cfg["num_attention_heads"], cfg["attention_head_dim"], cfg["patch_size"], cfg["max_text_seq_length"],
           cfg["sample_width"], cfg["sample_height"], cfg["max_n_view"], cfg["spatial_interpolation_scale"],
           cfg["temporal_interpolation_scale"], geom, str(like.device), like.dtype)
    t = _TABLES.get(key)
    if t is None:
        t = joint_pos_embedding(cfg, *geom) if kind == "joint" else view_pos_embedding(cfg, *geom)
        t = t.to(device=like.device, dtype=like.dtype)
        if len(_TABLES) > 64:
            _TABLES.clear()
        _TABLES[key] = t
    return t


def view_pos_embedding(cfg: dict, n_view: int) -> Tensor:
    """cogvideox_control.py:659-688: 3-D sin-cos whose 'time' axis is the view index -> [n_view * h*w, D]."""
    p = cfg["patch_size"]
    D = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    full = sincos_3d(D, (cfg["sample_width"] // p, cfg["sample_height"] // p), cfg["max_n_view"],
                     cfg["spatial_interpolation_scale"], 1.0)  # [max_view, hw, D]
    return full[:n_view].flatten(0, 1)


def timestep_sinusoid(t: Tensor, dim: int, flip_sin_to_cos: bool, freq_shift: float) -> Tensor:
    """diffusers get_timestep_embedding (App. A.4)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - freq_shift)
    emb = t[:, None].float() * torch.exp(exponent)[None]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


def rope_1d(dim: int, pos: Tensor, theta: float = 10000.0) -> Tuple[Tensor, Tensor]:
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    fr = torch.outer(pos.float(), freqs)
    return fr.cos().repeat_interleave(2, dim=1).float(), fr.sin().repeat_interleave(2, dim=1).float()


def rope_3d(embed_dim: int, crops_coords, grid_size: Tuple[int, int], temporal_size: int, grid_type: str = "linspace",
            max_size: Optional[Tuple[int, int]] = None) -> Tuple[Tensor, Tensor]:
    """diffusers get_3d_rotary_pos_embed (App. A.5) -> cos, sin [T*H*W, embed_dim]."""
    gh, gw = grid_size
    if grid_type == "linspace":
        start, stop = crops_coords
        grid_h = torch.linspace(start[0], stop[0] * (gh - 1) / gh, gh, dtype=torch.float32)
        grid_w = torch.linspace(start[1], stop[1] * (gw - 1) / gw, gw, dtype=torch.float32)
        grid_t = torch.linspace(0, temporal_size * (temporal_size - 1) / temporal_size, temporal_size,
                                dtype=torch.float32)
    elif grid_type == "slice":
        mh, mw = max_size
        grid_h = torch.arange(mh, dtype=torch.float32)
        grid_w = torch.arange(mw, dtype=torch.float32)
        grid_t = torch.arange(temporal_size, dtype=torch.float32)
    else:
        raise ValueError(grid_type)
    dim_t, dim_h, dim_w = embed_dim // 4, embed_dim // 8 * 3, embed_dim // 8 * 3
    tc, ts = rope_1d(dim_t, grid_t)
    hc, hs = rope_1d(dim_h, grid_h)
    wc, ws = rope_1d(dim_w, grid_w)
    if grid_type == "slice":
        tc, ts, hc, hs, wc, ws = tc[:temporal_size], ts[:temporal_size], hc[:gh], hs[:gh], wc[:gw], ws[:gw]

    def comb(a, b, c):
        a = a[:, None, None, :].expand(-1, gh, gw, -1)
        b = b[None, :, None, :].expand(temporal_size, -1, gw, -1)
        c = c[None, None, :, :].expand(temporal_size, gh, -1, -1)
        return torch.cat([a, b, c], dim=-1).reshape(temporal_size * gh * gw, -1)

    return comb(tc, hc, wc), comb(ts, hs, ws)


def resize_crop_region(src: Tuple[int, int], tgt_width: int, tgt_height: int):
    """get_resize_crop_region_for_grid (reference orv/utils.py:177-193)."""
    tw, th = tgt_width, tgt_height
    h, w = src
    r = h / w
    if r > (th / tw):
        rh, rw = th, int(round(th / h * w))
    else:
        rw, rh = tw, int(round(tw / w * h))
    top, left = int(round((th - rh) / 2.0)), int(round((tw - rw) / 2.0))
    return (top, left), (top + rh, left + rw)


def pipeline_rope(cfg: dict, height_px: int, width_px: int, num_latent_frames: int) -> Tuple[Tensor, Tensor]:
    """diffusers CogVideoXImageToVideoPipeline._prepare_rotary_positional_embeddings (App. A.0 note)."""
    p = cfg["patch_size"]
    gh, gw = height_px // (8 * p), width_px // (8 * p)
    pt = cfg["patch_size_t"]
    base_w, base_h = cfg["sample_width"] // p, cfg["sample_height"] // p
    if pt is None:
        crops = resize_crop_region((gh, gw), base_w, base_h)
        return rope_3d(cfg["attention_head_dim"], crops, (gh, gw), num_latent_frames)
    base_frames = (num_latent_frames + pt - 1) // pt
    return rope_3d(cfg["attention_head_dim"], None, (gh, gw), base_frames, grid_type="slice",
                   max_size=(base_h, base_w))


def apply_rope(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """diffusers apply_rotary_emb(use_real=True, unbind_dim=-1); x [B, H, S, 64]."""
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos[None, None] + rot.float() * sin[None, None]).to(x.dtype)


# ------------------------------------------------------------------------------------------------------------
# model pieces
# ------------------------------------------------------------------------------------------------------------
def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def _ln(sd, prefix, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd.get(prefix + ".weight"), sd.get(prefix + ".bias"), eps)


def patch_embed(sd, cfg, text: Tensor, image: Tensor) -> Tensor:
    """CogVideoXPatchEmbed.forward (App. A.1): text [B,St,4096], image [B,F,C,H,W] -> [B, St+N, D]."""
    text = _lin(sd, "patch_embed.text_proj", text)
    B, Fr, C, H, W = image.shape
    p, pt = cfg["patch_size"], cfg["patch_size_t"]
    if pt is None:
        x = image.reshape(-1, C, H, W)
        x = F.conv2d(x, sd["patch_embed.proj.weight"], sd.get("patch_embed.proj.bias"), stride=p)
        x = x.view(B, Fr, *x.shape[1:]).flatten(3).transpose(2, 3).flatten(1, 2)
    else:
        x = image.permute(0, 1, 3, 4, 2)
        x = x.reshape(B, Fr // pt, pt, H // p, p, W // p, p, C)
        x = x.permute(0, 1, 3, 5, 7, 2, 4, 6).flatten(4, 7).flatten(1, 3)
        x = _lin(sd, "patch_embed.proj", x)
    emb = torch.cat([text, x], dim=1).contiguous()
    if (not cfg["use_rotary_positional_embeddings"]) or cfg["use_learned_positional_embeddings"]:
        pos = _cached_table("joint", cfg, (Fr, H, W), emb)  # a buffer in the reference: built once, not per forward
        emb = emb + pos[None]
    return emb


def action_embed(sd, cfg, actions: Tensor, mask: Optional[Tensor]) -> Tensor:
    """Action padding (cogvideox_control.py:805-811) + ActionEmbed.forward (components.py:47-71).
    `mask` [B] bool: rows replaced by mask_embed (the random draw itself is the caller's)."""
    res = (actions.size(1) + 1) % 4
    if res > 0:
        pad = actions.new_zeros((actions.shape[0], 4 - res, actions.shape[2]))
        actions = torch.cat([pad, actions], dim=1)
    B, Fa, _ = actions.shape
    x = torch.cat([torch.zeros_like(actions[:, :1]), actions], dim=1)
    x = x.reshape(B, (Fa + 1) // 4, -1)
    pt = cfg["patch_size_t"] or 1
    if pt > 1:
        x = x.reshape(B, x.shape[1] // pt, -1)
    x = _lin(sd, "action_embed.mlp.0", x)
    x = F.gelu(x, approximate="tanh")
    x = _lin(sd, "action_embed.mlp.3", x)
    if mask is not None and bool(mask.any()):
        x = x.clone()
        x[mask] = sd["action_embed.mask_embed.weight"][None].repeat(int(mask.sum()), x.shape[1], 1).to(x.dtype)
    return x


def layernorm_zero(sd, prefix, cfg, hidden, enc, temb, action_emb):
    """ORV CogVideoXLayerNormZero.forward: modulate_encoder_hidden_states=False branches (:70-99: the linear is
    [3D, T], the text gets a plain LayerNorm and no gate) and =True branches (:101-145)."""
    D = hidden.shape[-1]
    W, Bv = sd[prefix + ".linear.weight"], sd[prefix + ".linear.bias"]
    eps = cfg["norm_eps"]
    if not cfg["modulate_encoder_hidden_states"]:
        e = _ln(sd, prefix + ".norm", enc, eps)
        if action_emb is None:
            shift, scale, gate = F.linear(F.silu(temb), W, Bv).chunk(3, dim=-1)
            h = _ln(sd, prefix + ".norm", hidden, eps) * (1 + scale)[:, None] + shift[:, None]
            return h, e, gate[:, None], None
        shift, scale, gate = F.linear(F.silu(temb[:, None] + action_emb), W, Bv).chunk(3, dim=-1)
        n = hidden.shape[1] // action_emb.size(1)
        scale, shift, gate = (t.repeat_interleave(n, dim=1) for t in (scale, shift, gate))
        return _ln(sd, prefix + ".norm", hidden, eps) * (1 + scale) + shift, e, gate, None
    if action_emb is None:
        shift, scale, gate, e_shift, e_scale, e_gate = F.linear(F.silu(temb), W, Bv).chunk(6, dim=-1)
        h = _ln(sd, prefix + ".norm", hidden, eps) * (1 + scale)[:, None] + shift[:, None]
        e = _ln(sd, prefix + ".norm", enc, eps) * (1 + e_scale)[:, None] + e_shift[:, None]
        return h, e, gate[:, None], e_gate[:, None]
    shift, scale, gate = F.linear(F.silu(temb[:, None] + action_emb), W[: 3 * D], Bv[: 3 * D]).chunk(3, dim=-1)
    e_shift, e_scale, e_gate = F.linear(F.silu(temb), W[3 * D:], Bv[3 * D:]).chunk(3, dim=-1)
    n = hidden.size(1) // action_emb.size(1)
    scale, shift, gate = (t.repeat_interleave(n, dim=1) for t in (scale, shift, gate))
    h = _ln(sd, prefix + ".norm", hidden, eps) * (1 + scale) + shift
    e = _ln(sd, prefix + ".norm", enc, eps) * (1 + e_scale)[:, None] + e_shift[:, None]
    return h, e, gate, e_gate[:, None]


def attention(sd, prefix, cfg, hidden, enc, rope):
    """ORV CogVideoXAttnProcessor2_0.__call__ (:200-270).  enc=None: video tokens only (text_seq_length = 0, :222-226)."""
    St = enc.size(1) if enc is not None else 0
    x = torch.cat([enc, hidden], dim=1) if enc is not None else hidden
    B, S, D = x.shape
    H = cfg["num_attention_heads"]
    q = _lin(sd, prefix + ".to_q", x).view(B, S, H, -1).transpose(1, 2)
    k = _lin(sd, prefix + ".to_k", x).view(B, S, H, -1).transpose(1, 2)
    v = _lin(sd, prefix + ".to_v", x).view(B, S, H, -1).transpose(1, 2)
    q = _ln(sd, prefix + ".norm_q", q, 1e-6)
    k = _ln(sd, prefix + ".norm_k", k, 1e-6)
    if rope is not None:
        q = torch.cat([q[:, :, :St], apply_rope(q[:, :, St:], *rope)], dim=2)
        k = torch.cat([k[:, :, :St], apply_rope(k[:, :, St:], *rope)], dim=2)
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, S, D)
    o = _lin(sd, prefix + ".to_out.0", o)
    return o[:, St:], o[:, :St]


def block(sd, prefix, cfg, hidden, enc, temb, rope, action_emb):
    """ORV CogVideoXBlock.forward (:394-445).  Without text modulation (:404-424) the attention and the FFN see the
    video tokens only and the text stream passes through unchanged."""
    if not cfg["modulate_encoder_hidden_states"]:
        nh, _, gate, _ = layernorm_zero(sd, prefix + ".norm1", cfg, hidden, enc, temb, action_emb)
        ah, _ = attention(sd, prefix + ".attn1", cfg, nh, None, rope)
        hidden = hidden + gate * ah
        nh, _, gate, _ = layernorm_zero(sd, prefix + ".norm2", cfg, hidden, enc, temb, action_emb)
        x = _lin(sd, prefix + ".ff.net.2", F.gelu(_lin(sd, prefix + ".ff.net.0.proj", nh), approximate="tanh"))
        return hidden + gate * x, enc
    nh, ne, gate, e_gate = layernorm_zero(sd, prefix + ".norm1", cfg, hidden, enc, temb, action_emb)
    ah, ae = attention(sd, prefix + ".attn1", cfg, nh, ne, rope)
    hidden = hidden + gate * ah
    enc = enc + e_gate * ae
    nh, ne, gate, e_gate = layernorm_zero(sd, prefix + ".norm2", cfg, hidden, enc, temb, action_emb)
    St = enc.size(1)
    x = torch.cat([ne, nh], dim=1)
    x = _lin(sd, prefix + ".ff.net.0.proj", x)
    x = F.gelu(x, approximate="tanh")
    x = _lin(sd, prefix + ".ff.net.2", x)
    hidden = hidden + gate * x[:, St:]
    enc = enc + e_gate * x[:, :St]
    return hidden, enc


def mv_block(sd, prefix, cfg, hidden, enc, temb, n_view, n_frame):
    """ORV MVBlock.forward (:313-348), modulate_encoder_hidden_states=True."""
    nh, ne, gate, _ = layernorm_zero(sd, prefix + ".norm1", cfg, hidden, enc, temb, None)
    BV, S, D = nh.shape
    b = BV // n_view
    s = S // n_frame
    nh = nh.view(b, n_view, n_frame, s, D).permute(0, 2, 1, 3, 4).reshape(b * n_frame, n_view * s, D)
    ne = ne.view(b, n_view * ne.shape[1], D)
    ne = ne[:, None].expand(-1, n_frame, -1, -1).reshape(b * n_frame, -1, D)
    ah, _ = attention(sd, prefix + ".attn1", cfg, nh, ne, None)
    ah = _lin(sd, prefix + ".proj_out", ah)
    ah = ah.view(b, n_frame, n_view, s, D).permute(0, 2, 1, 3, 4).reshape(BV, S, D)
    return hidden + gate * ah


def forward(sd: Dict[str, Tensor], cfg: dict, hidden_states: Tensor, encoder_hidden_states: Tensor,
            timestep: Tensor, actions: Optional[Tensor] = None, action_mask: Optional[Tensor] = None,
            depths: Optional[Tensor] = None, labels: Optional[Tensor] = None, ofs: Optional[Tensor] = None,
            rope: Optional[Tuple[Tensor, Tensor]] = None, num_views: int = 1, taps: Optional[dict] = None) -> Tensor:
    """CogVideoXTransformer3DModelTraj.forward (:715-948).  Computes in the dtype of `sd` / inputs."""
    if not cfg["modulate_encoder_hidden_states"] and cfg["multiview"]:
        raise NotImplementedError("multiview without text modulation: no ORV config ships it")
    V = num_views
    if V > 1:
        Bc, VF = hidden_states.shape[:2]
        hidden_states = hidden_states.reshape(Bc, V, VF // V, *hidden_states.shape[2:]).flatten(0, 1)
        encoder_hidden_states = encoder_hidden_states.repeat_interleave(V, dim=0)
    B, Fr, _, H, W = hidden_states.shape
    dt = hidden_states.dtype
    D = cfg["num_attention_heads"] * cfg["attention_head_dim"]

    t_emb = timestep_sinusoid(timestep, D, cfg["flip_sin_to_cos"], cfg["freq_shift"]).to(dt)
    temb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", t_emb)))
    if cfg["ofs_embed_dim"] is not None:
        o = timestep_sinusoid(ofs, cfg["ofs_embed_dim"], cfg["flip_sin_to_cos"], cfg["freq_shift"]).to(dt)
        temb = temb + _lin(sd, "ofs_embedding.linear_2", F.silu(_lin(sd, "ofs_embedding.linear_1", o)))
    if V > 1:
        temb = temb.repeat_interleave(V, dim=0)

    x = patch_embed(sd, cfg, encoder_hidden_states, hidden_states)
    St = encoder_hidden_states.shape[1]
    enc, hid = x[:, :St], x[:, St:]
    if V > 1:
        b = B // V
        s = hid.shape[1] // Fr
        h5 = hid.view(b, V, Fr, s, D).permute(0, 2, 1, 3, 4).reshape(b * Fr, V * s, D)
        h5 = h5 + _cached_table("view", cfg, (V,), h5)[None]
        hid = h5.view(b, Fr, V, s, D).permute(0, 2, 1, 3, 4).reshape(B, Fr * s, D)

    action_emb = None
    if actions is not None:
        action_emb = action_embed(sd, cfg, actions, action_mask)
        if V > 1:
            action_emb = action_emb.repeat_interleave(V, dim=0)

    ctrl = []
    if cfg["visual_guidance"]:
        for c in (depths, labels):
            if c is None:
                continue
            if V > 1:
                c = c.reshape(c.shape[0], V, c.shape[1] // V, *c.shape[2:]).flatten(0, 1)
            ctrl.append(patch_embed(sd, cfg, encoder_hidden_states, c)[:, St:])
    if ctrl:
        assert len(ctrl) == cfg["num_control_keys"]
        c = torch.cat(ctrl, dim=-1)
        hid = hid + _lin(sd, "initial_combine_linear", hid.repeat(1, 1, cfg["num_control_keys"]) + c)

    for i in range(cfg["num_layers"]):
        if cfg["multiview"]:
            hid = mv_block(sd, f"mv_blocks.{i}", cfg, hid, enc, temb, V, Fr)
        hid, enc = block(sd, f"transformer_blocks.{i}", cfg, hid, enc, temb, rope, action_emb)
        if taps is not None and i in taps.get("layers", ()):
            taps[i] = torch.cat([enc, hid], dim=1).clone()

    hid = _ln(sd, "norm_final", hid, cfg["norm_eps"])
    # ORV AdaLayerNorm.forward (:153-197): chunk order shift, scale
    e = temb if action_emb is None else temb[:, None] + action_emb
    e = _lin(sd, "norm_out.linear", F.silu(e))
    if action_emb is None:
        shift, scale = e.chunk(2, dim=1)
        shift, scale = shift[:, None], scale[:, None]
    else:
        shift, scale = e.chunk(2, dim=2)
        n = hid.shape[1] // action_emb.size(1)
        scale, shift = scale.repeat_interleave(n, dim=1), shift.repeat_interleave(n, dim=1)
    hid = _ln(sd, "norm_out.norm", hid, cfg["norm_eps"]) * (1 + scale) + shift
    y = _lin(sd, "proj_out", hid)

    p, pt = cfg["patch_size"], cfg["patch_size_t"]
    if pt is None:
        out = y.reshape(B, Fr, H // p, W // p, -1, p, p).permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)
    else:
        out = y.reshape(B, (Fr + pt - 1) // pt, H // p, W // p, -1, pt, p, p)
        out = out.permute(0, 1, 5, 4, 2, 6, 3, 7).flatten(6, 7).flatten(4, 5).flatten(1, 2)
    if V > 1:
        out = out.reshape(B // V, V * out.shape[1], *out.shape[2:])
    return out


# ------------------------------------------------------------------------------------------------------------
# integer index maps in closed form — used for bit-exact checks
# ------------------------------------------------------------------------------------------------------------
def patchify_index_map(Fr: int, C: int, H: int, W: int, p: int, pt: Optional[int]) -> Tensor:
    """Returns idx [tokens, K] such that patches.flatten()[...] = x[f, c, h, w] flat index (single sample)."""
    t = pt or 1
    Fp, Hp, Wp = Fr // t, H // p, W // p
    fp, i, j, c, tt, kh, kw = torch.meshgrid(torch.arange(Fp), torch.arange(Hp), torch.arange(Wp), torch.arange(C),
                                             torch.arange(t), torch.arange(p), torch.arange(p), indexing="ij")
    src = (((fp * t + tt) * C + c) * H + i * p + kh) * W + j * p + kw
    return src.reshape(Fp * Hp * Wp, C * t * p * p)


# ------------------------------------------------------------------------------------------------------------
# schedulers (App. A.7) and the sampler loop
# ------------------------------------------------------------------------------------------------------------
class Scheduler:
    """CogVideoXDDIMScheduler / CogVideoXDPMScheduler arithmetic (v-prediction, zero-terminal-SNR, trailing)."""

    def __init__(self, kind: str = "dpm", num_train_timesteps: int = 1000, beta_start: float = 0.00085,
                 beta_end: float = 0.012, snr_shift_scale: float = 3.0, rescale_betas_zero_snr: bool = True):
        self.kind = kind
        self.T = num_train_timesteps
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
        ac = torch.cumprod(1.0 - betas, dim=0)
        ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
        if rescale_betas_zero_snr:
            s = ac.sqrt()
            s0, sT = s[0].clone(), s[-1].clone()
            s = (s - sT) * (s0 / (s0 - sT))
            ac = s ** 2
        self.alphas_cumprod = ac
        self.final_alpha_cumprod = torch.tensor(1.0, dtype=torch.float64)
        self.num_inference_steps = None
        self.timesteps = None

    def set_timesteps(self, n: int):
        import numpy as np
        self.num_inference_steps = n
        ts = np.round(np.arange(self.T, 0, -self.T / n)).astype(np.int64) - 1  # "trailing"
        self.timesteps = torch.from_numpy(ts)

    def _alphas(self, t: int):
        prev = t - self.T // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return a_t, a_prev, prev

    @staticmethod
    def _smul(scalar, x: Tensor) -> Tensor:
        """`float64 0-dim CPU tensor * tensor` as torch evaluates it when the tensor lives on CUDA (the reference's
        deployment): the scalar is taken in fp32 ("opmath"), the product is rounded to the tensor's dtype.
        (torch's CPU kernels instead round the scalar to the tensor dtype first; tests/test_gpu_ops.py checks the
        CUDA behaviour against real torch-CUDA ops on the B200.)"""
        return (x.float() * torch.tensor(float(scalar), dtype=torch.float32)).to(x.dtype)

    def step_ddim(self, v: Tensor, t: int, x: Tensor) -> Tensor:
        """CogVideoXDDIMScheduler.step (v-prediction); x may be bf16 (the pipeline keeps latents in bf16), v fp32."""
        a_t, a_prev, _ = self._alphas(t)
        x0 = self._smul(a_t ** 0.5, x) - self._smul((1 - a_t) ** 0.5, v)
        a = ((1 - a_prev) / (1 - a_t)) ** 0.5
        b = a_prev ** 0.5 - a_t ** 0.5 * a
        return self._smul(a, x) + self._smul(b, x0)

    def step_dpm(self, v: Tensor, old_x0: Optional[Tensor], t: int, t_back: Optional[int], x: Tensor,
                 generator: Optional[torch.Generator], noises: Optional[list] = None):
        """CogVideoXDPMScheduler.step.  `noises` (optional) collects the noise tensors actually used."""
        a_t, a_prev, prev = self._alphas(t)
        a_back = self.alphas_cumprod[t_back] if t_back is not None else None
        x0 = self._smul(a_t ** 0.5, x) - self._smul((1 - a_t) ** 0.5, v)
        lamb = ((a_t / (1 - a_t)) ** 0.5).log()
        lamb_next = ((a_prev / (1 - a_prev)) ** 0.5).log()
        h = lamb_next - lamb
        m1 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp()
        m2 = (-2 * h).expm1() * a_prev ** 0.5
        m_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        noise = torch.randn(x.shape, generator=generator, dtype=x.dtype)
        if old_x0 is None or prev < 0:
            if noises is not None:
                noises.append(noise)
            return self._smul(m1, x) - self._smul(m2, x0) + self._smul(m_noise, noise), x0
        lamb_prev = ((a_back / (1 - a_back)) ** 0.5).log()
        r = (lamb - lamb_prev) / h
        m3, m4 = 1 + 1 / (2 * r), 1 / (2 * r)
        d = self._smul(m3, x0) - self._smul(m4, old_x0)
        noise = torch.randn(x.shape, generator=generator, dtype=x.dtype)  # diffusers draws a second time here
        if noises is not None:
            noises.append(noise)
        return self._smul(m1, x) - self._smul(m2, d) + self._smul(m_noise, noise), x0


def sample_loop(sd, cfg, sched: Scheduler, latents: Tensor, image_latents: Tensor, prompt_embeds: Tensor,
                num_steps: int, guidance_scale: float = 1.0, generator: Optional[torch.Generator] = None,
                model_dtype=None, **fwd_kwargs) -> Tensor:
    """Denoise loop of CogVideoXImageToVideoPipelineTraj.__call__ (:1402-1473).  latents keep prompt dtype."""
    sched.set_timesteps(num_steps)
    cfg_on = guidance_scale > 1.0
    old_x0 = None
    ts = sched.timesteps
    for i, t in enumerate(ts.tolist()):
        x_in = torch.cat([latents] * 2) if cfg_on else latents
        img = torch.cat([image_latents] * 2) if cfg_on else image_latents
        x_in = torch.cat([x_in, img], dim=2)
        timestep = torch.full((x_in.shape[0],), t, dtype=torch.int64)
        v = forward(sd, cfg, x_in, prompt_embeds, timestep, **fwd_kwargs).float()
        if cfg_on:
            u, c = v.chunk(2)
            v = u + guidance_scale * (c - u)
        if sched.kind == "ddim":
            latents = sched.step_ddim(v, t, latents)
        else:
            latents, old_x0 = sched.step_dpm(v, old_x0, t, ts[i - 1].item() if i > 0 else None, latents, generator)
        latents = latents.to(prompt_embeds.dtype)
    return latents


def pipeline_call(sd, cfg, kind: str, moments: Tensor, prompt_embeds: Tensor, num_frames: int, height: int,
                  width: int, num_steps: int, guidance_scale: float, generator: torch.Generator,
                  actions: Optional[Tensor] = None, negative_prompt_embeds: Optional[Tensor] = None,
                  scaling_factor: float = 1.15258426, num_views: int = 1,
                  control_moments: Optional[Dict[str, Tensor]] = None, **fwd_kwargs) -> Tensor:
    """CogVideoXImageToVideoPipelineTraj.__call__ for latent-moment inputs and output_type='latent', patch_size_t
    None (:1227-1489): control latents (:1331-1364: depth / label moments sampled with the GLOBAL RNG — no generator
    is passed there —, scaled, and duplicated on the channel axis), prepare_latents (:1115-1225: sample first-frame
    moments, scale, zero-pad every view to F frames, initial noise) followed by the denoise loop.  RNG draws happen in
    the reference's order on `generator`."""
    dt = prompt_embeds.dtype
    B = prompt_embeds.shape[0]
    if guidance_scale > 1.0:
        prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)

    def sample(m: Tensor, gen) -> Tensor:  # DiagonalGaussianDistribution(m).sample(gen), App. A.8
        mean, logvar = torch.chunk(m, 2, dim=1)
        std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
        return mean + std * torch.randn(mean.shape, generator=gen, dtype=m.dtype)

    for key in ("depths", "labels"):
        if control_moments is not None and control_moments.get(key) is not None:
            lat = (scaling_factor * sample(control_moments[key], None)).permute(0, 2, 1, 3, 4)
            fwd_kwargs[key] = torch.cat([lat, lat], dim=2).to(dt)
    lat_frames = (num_frames - 1) // 4 + 1
    V = num_views
    img = (scaling_factor * sample(moments.to(dt), generator)).permute(0, 2, 1, 3, 4)  # [B, V*F_img, C, h, w]
    img = img.reshape(B, V, img.shape[1] // V, *img.shape[2:])
    pad = torch.zeros(B, V, lat_frames - img.shape[2], img.shape[3], height // 8, width // 8, dtype=dt)
    image_latents = torch.cat([img, pad], dim=2).flatten(1, 2)
    latents = torch.randn((B, V * lat_frames, img.shape[3], height // 8, width // 8), generator=generator, dtype=dt)
    sched = Scheduler(kind)
    if V > 1:
        fwd_kwargs["num_views"] = V
    out = sample_loop(sd, cfg, sched, latents, image_latents, prompt_embeds, num_steps, guidance_scale, generator,
                      actions=actions, **fwd_kwargs)
    return out.reshape(B * V, lat_frames, *out.shape[2:])  # 'b (v f) c h w -> (b v) f c h w' (:1476)


# ------------------------------------------------------------------------------------------------------------
# synthetic weights: every path contributes, small std so 30 layers stay well-conditioned
# ------------------------------------------------------------------------------------------------------------
def param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    D = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    T, p, pt, C = cfg["time_embed_dim"], cfg["patch_size"], cfg["patch_size_t"], cfg["in_channels"]
    hd = cfg["attention_head_dim"]
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(name, o, i, bias=True):
        s[name + ".weight"] = (o, i)
        if bias:
            s[name + ".bias"] = (o,)

    def ln(name, d):
        s[name + ".weight"] = (d,)
        s[name + ".bias"] = (d,)

    if pt is None:
        s["patch_embed.proj.weight"] = (D, C, p, p)
        if cfg["patch_bias"]:
            s["patch_embed.proj.bias"] = (D,)
    else:
        lin("patch_embed.proj", D, C * p * p * pt)
    lin("patch_embed.text_proj", D, cfg["text_embed_dim"])
    lin("time_embedding.linear_1", T, D)
    lin("time_embedding.linear_2", T, T)
    if cfg["ofs_embed_dim"] is not None:
        lin("ofs_embedding.linear_1", cfg["ofs_embed_dim"], cfg["ofs_embed_dim"])
        lin("ofs_embedding.linear_2", cfg["ofs_embed_dim"], cfg["ofs_embed_dim"])

    def attn(pre):
        for n in ("to_q", "to_k", "to_v"):
            lin(f"{pre}.{n}", D, D, cfg["attention_bias"])
        ln(f"{pre}.norm_q", hd)
        ln(f"{pre}.norm_k", hd)
        lin(f"{pre}.to_out.0", D, D)

    for i in range(cfg["num_layers"]):
        pre = f"transformer_blocks.{i}"
        for n in ("norm1", "norm2"):
            lin(f"{pre}.{n}.linear", (6 if cfg["modulate_encoder_hidden_states"] else 3) * D, T)
            ln(f"{pre}.{n}.norm", D)
        attn(f"{pre}.attn1")
        lin(f"{pre}.ff.net.0.proj", 4 * D, D)
        lin(f"{pre}.ff.net.2", D, 4 * D)
        if cfg["multiview"]:
            mp = f"mv_blocks.{i}"
            lin(f"{mp}.norm1.linear", 6 * D, T)
            ln(f"{mp}.norm1.norm", D)
            attn(f"{mp}.attn1")
            lin(f"{mp}.cam_encoder", D, 12)
            lin(f"{mp}.proj_out", D, D)
    ln("norm_final", D)
    lin("norm_out.linear", 2 * D, T)
    ln("norm_out.norm", D)
    lin("proj_out", p * p * (pt or 1) * cfg["out_channels"], D)
    lin("action_embed.mlp.0", 4 * T, 7 * 4 * (pt or 1))
    lin("action_embed.mlp.3", T, 4 * T)
    s["action_embed.mask_embed.weight"] = (1, T)
    if cfg["visual_guidance"]:
        lin("initial_combine_linear", D, D * cfg["num_control_keys"])
    return s


def synthetic_state_dict(cfg: dict, seed: int = 0, std: float = 0.02, dtype=torch.float32) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg).items():
        is_ln_weight = name.endswith(".weight") and len(shape) == 1
        t = torch.randn(shape, generator=g) * std
        if is_ln_weight:
            t = 1.0 + t
        sd[name] = t.to(dtype)
    return sd


def synthetic_inputs(cfg: dict, batch: int, frames_lat: int, height: int, width: int, seed: int = 1,
                     with_controls: bool = False, n_actions: int = 16, text_len: Optional[int] = None) -> dict:
    """Seeded inputs per SURVEY §8(d); all fp32 (callers cast)."""
    g = torch.Generator().manual_seed(seed)
    C = cfg["in_channels"]
    St = text_len if text_len is not None else cfg["max_text_seq_length"]
    lat = torch.randn(batch, frames_lat, C // 2, height, width, generator=g)
    img = torch.zeros(batch, frames_lat, C // 2, height, width)
    img[:, 0] = torch.randn(batch, C // 2, height, width, generator=g)
    text = torch.randn(batch, St, cfg["text_embed_dim"], generator=g) * 0.2
    act = (torch.rand(batch, n_actions, 7, generator=g) * 2 - 1) * torch.tensor([20.0] * 6 + [1.0])
    act[..., -1] = act[..., -1].abs()
    out = dict(latents=lat, image_latents=img, hidden_states=torch.cat([lat, img], dim=2), text=text, actions=act)
    if with_controls:
        out["depths"] = torch.randn(batch, frames_lat, C, height, width, generator=g)
        out["labels"] = torch.randn(batch, frames_lat, C, height, width, generator=g)
    return out
