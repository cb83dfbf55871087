"""TEST INFRASTRUCTURE — torch restatement of the DECODE half of diffusers' `AutoencoderKLCogVideoX` (0.32 line),
the call the reference pipeline makes at orv/models/cogvideox_control.py:1476-1479 (`decode_latents`, :1095-1100) with
tiling and slicing switched on by orv/pipeline/inference_control_to_video.py:98-99.  Used only by tests/ and tools/.

**PARITY UNPINNED.**  The algorithm lives in a third-party dependency that is absent from /root/reference and from this
image (diffusers >= 0.31.2, requirements.txt:21; no wheel, no network), and the reference ships no decoded golden
frames.  What follows restates the published module structure from its source as the author remembers it
(`diffusers/models/autoencoders/autoencoder_kl_cogvideox.py`: CogVideoXCausalConv3d, CogVideoXSpatialNorm3D,
CogVideoXResnetBlock3D, CogVideoXMidBlock3D, CogVideoXUpBlock3D, CogVideoXUpsample3D
(`models/upsampling.py`), CogVideoXDecoder3D, AutoencoderKLCogVideoX.{decode,_decode,tiled_decode,blend_v,blend_h}),
with the diffusers parameter names so that a real checkpoint's `state_dict` loads.  The CUDA path is tested against THIS
restatement; agreement with the real package cannot be established here.

Conventions restated (each is a place where a mis-remembered detail would go unnoticed):
  * CausalConv3d, pad_mode "constant": spatial zero padding k//2, temporal FRONT padding by k-1 frames taken from the
    convolution cache (the last k-1 input frames of the previous frame batch) or, for the first batch, the first frame
    repeated; kernel-1 convs have no temporal context.
  * SpatialNorm3D: GroupNorm(32, eps 1e-6)(f) * conv_y(zq') + conv_b(zq'), zq' = nearest-neighbour resize of the
    latent chunk to f's (T, H, W); when T is odd and > 1 the first frame is resized separately from the rest.
  * Upsample3D: nearest x2 in H, W, and — in the first two up blocks (compress_time) — in T, except that the first
    frame of an odd-length batch is not duplicated in time; then a per-frame 3x3 Conv2d.
  * decode: latent frames are decoded in batches of 2 (the first batch takes the remainder, so 5 frames -> [0:3], [3:5])
    that share convolution caches; GroupNorm statistics are per batch.
  * tiled_decode: 30 x 45 latent tiles (sample 480 x 720 // 2 // 8) stepping 25 x 36, blended over 40 x 72 output
    pixels, cropped to 200 x 288.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def default_config(**over) -> dict:
    cfg = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 256, 512), latent_channels=16,
               layers_per_block=3, norm_eps=1e-6, norm_num_groups=32, temporal_compression_ratio=4, sample_height=480,
               sample_width=720, scaling_factor=1.15258426, invert_scale_latents=False)
    cfg.update(over)
    return cfg


# ------------------------------------------------------------------------------------------------------------
# parameter shapes (diffusers key names) and seeded synthetic weights
# ------------------------------------------------------------------------------------------------------------
def decoder_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}
    zc = cfg["latent_channels"]
    rev = tuple(reversed(cfg["block_out_channels"]))

    def conv3(name, cin, cout, k=3):
        s[f"{name}.conv.weight"] = (cout, cin, k, k, k)
        s[f"{name}.conv.bias"] = (cout,)

    def snorm(name, c):
        s[f"{name}.norm_layer.weight"] = (c,)
        s[f"{name}.norm_layer.bias"] = (c,)
        conv3(f"{name}.conv_y", zc, c, 1)
        conv3(f"{name}.conv_b", zc, c, 1)

    def resnet(name, cin, cout):
        snorm(f"{name}.norm1", cin)
        snorm(f"{name}.norm2", cout)
        conv3(f"{name}.conv1", cin, cout)
        conv3(f"{name}.conv2", cout, cout)
        if cin != cout:
            s[f"{name}.conv_shortcut.weight"] = (cout, cin, 1, 1, 1)
            s[f"{name}.conv_shortcut.bias"] = (cout,)

    conv3("decoder.conv_in", zc, rev[0])
    for i in range(2):
        resnet(f"decoder.mid_block.resnets.{i}", rev[0], rev[0])
    cout = rev[0]
    for b, ch in enumerate(rev):
        cin, cout = cout, ch
        for i in range(cfg["layers_per_block"] + 1):
            resnet(f"decoder.up_blocks.{b}.resnets.{i}", cin if i == 0 else cout, cout)
        if b != len(rev) - 1:
            s[f"decoder.up_blocks.{b}.upsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            s[f"decoder.up_blocks.{b}.upsamplers.0.conv.bias"] = (cout,)
    snorm("decoder.norm_out", rev[-1])
    conv3("decoder.conv_out", rev[-1], cfg["out_channels"])
    return s


def synthetic_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Variance-preserving random weights (std = 1 / sqrt(fan_in)), GroupNorm weights around 1."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in decoder_param_shapes(cfg).items():
        if name.endswith("norm_layer.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        elif ".conv_y." in name:  # multiplicative branch of the spatial norm: keep it around 1
            t = torch.randn(shape, generator=g) * (0.3 / (shape[1] ** 0.5))
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) / (fan_in ** 0.5)
        sd[name] = t.to(dtype)
    for k in list(sd):
        if k.endswith("conv_y.conv.bias"):
            sd[k] = sd[k] + 1.0
    return sd


# ------------------------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------------------------
def causal_conv3d(sd, name: str, x: Tensor, cache: Optional[Tensor]) -> Tuple[Tensor, Optional[Tensor]]:
    """CogVideoXCausalConv3d.forward, pad_mode 'constant'.  x [B, C, T, H, W].  Returns (y, new cache)."""
    w, b = sd[f"{name}.conv.weight"], sd[f"{name}.conv.bias"]
    kt, kh, kw = w.shape[2:]
    if kt > 1:
        ctx = [cache] if cache is not None else [x[:, :, :1]] * (kt - 1)
        x = torch.cat(ctx + [x], dim=2)
    new_cache = x[:, :, -(kt - 1):].clone() if kt > 1 else None
    return F.conv3d(x, w, b, stride=1, padding=(0, kh // 2, kw // 2)), new_cache


def _resize_nearest(z: Tensor, size) -> Tensor:
    return F.interpolate(z, size=tuple(size), mode="nearest")


def spatial_norm(sd, name: str, f: Tensor, zq: Tensor, cfg: dict) -> Tensor:
    """CogVideoXSpatialNorm3D.forward."""
    T = f.shape[2]
    if T > 1 and T % 2 == 1:
        z = torch.cat([_resize_nearest(zq[:, :, :1], (1,) + tuple(f.shape[-2:])),
                       _resize_nearest(zq[:, :, 1:], (T - 1,) + tuple(f.shape[-2:]))], dim=2)
    else:
        z = _resize_nearest(zq, f.shape[-3:])
    y, _ = causal_conv3d(sd, f"{name}.conv_y", z, None)
    b, _ = causal_conv3d(sd, f"{name}.conv_b", z, None)
    nf = F.group_norm(f, cfg["norm_num_groups"], sd[f"{name}.norm_layer.weight"], sd[f"{name}.norm_layer.bias"], 1e-6)
    return nf * y + b


def resnet_block(sd, name: str, x: Tensor, zq: Tensor, cfg: dict, cache: dict) -> Tensor:
    """CogVideoXResnetBlock3D.forward (temb is None on the decode path)."""
    h = F.silu(spatial_norm(sd, f"{name}.norm1", x, zq, cfg))
    h, cache[f"{name}.conv1"] = causal_conv3d(sd, f"{name}.conv1", h, cache.get(f"{name}.conv1"))
    h = F.silu(spatial_norm(sd, f"{name}.norm2", h, zq, cfg))
    h, cache[f"{name}.conv2"] = causal_conv3d(sd, f"{name}.conv2", h, cache.get(f"{name}.conv2"))
    if f"{name}.conv_shortcut.weight" in sd:
        x = F.conv3d(x, sd[f"{name}.conv_shortcut.weight"], sd[f"{name}.conv_shortcut.bias"])
    return h + x


def upsample3d(sd, name: str, x: Tensor, compress_time: bool) -> Tensor:
    """CogVideoXUpsample3D.forward."""
    B, C, T, H, W = x.shape
    if compress_time:
        if T > 1 and T % 2 == 1:
            first = F.interpolate(x[:, :, 0], scale_factor=2.0)[:, :, None]
            rest = F.interpolate(x[:, :, 1:], scale_factor=2.0)
            x = torch.cat([first, rest], dim=2)
        elif T > 1:
            x = F.interpolate(x, scale_factor=2.0)
        else:
            x = F.interpolate(x[:, :, 0], scale_factor=2.0)[:, :, None]
    else:
        x = F.interpolate(x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W), scale_factor=2.0)
        x = x.reshape(B, T, C, 2 * H, 2 * W).permute(0, 2, 1, 3, 4)
    B, C, T, H, W = x.shape
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W), sd[f"{name}.conv.weight"], sd[f"{name}.conv.bias"],
                 padding=1)
    return y.reshape(B, T, -1, H, W).permute(0, 2, 1, 3, 4)


def decoder_forward(sd, cfg: dict, z: Tensor, cache: dict) -> Tensor:
    """CogVideoXDecoder3D.forward on one frame batch; `cache` carries the convolution caches to the next batch."""
    rev = tuple(reversed(cfg["block_out_channels"]))
    h, cache["conv_in"] = causal_conv3d(sd, "decoder.conv_in", z, cache.get("conv_in"))
    for i in range(2):
        h = resnet_block(sd, f"decoder.mid_block.resnets.{i}", h, z, cfg, cache)
    compress_level = 0
    r = cfg["temporal_compression_ratio"]
    while (1 << compress_level) < r:
        compress_level += 1
    for b in range(len(rev)):
        for i in range(cfg["layers_per_block"] + 1):
            h = resnet_block(sd, f"decoder.up_blocks.{b}.resnets.{i}", h, z, cfg, cache)
        if b != len(rev) - 1:
            h = upsample3d(sd, f"decoder.up_blocks.{b}.upsamplers.0", h, compress_time=b < compress_level)
    h = F.silu(spatial_norm(sd, "decoder.norm_out", h, z, cfg))
    h, cache["conv_out"] = causal_conv3d(sd, "decoder.conv_out", h, cache.get("conv_out"))
    return h


def frame_batches(num_frames: int, batch: int = 2) -> List[Tuple[int, int]]:
    """AutoencoderKLCogVideoX._decode: the first batch takes the remainder."""
    n = max(num_frames // batch, 1)
    rem = num_frames % batch
    out = []
    for i in range(n):
        start = batch * i + (0 if i == 0 else rem)
        end = batch * (i + 1) + rem
        out.append((start, min(end, num_frames)))
    return out


def decode_untiled(sd, cfg: dict, z: Tensor) -> Tensor:
    cache: dict = {}
    outs = []
    for s, e in frame_batches(z.shape[2]):
        outs.append(decoder_forward(sd, cfg, z[:, :, s:e], cache))
    return torch.cat(outs, dim=2)


def tile_geometry(cfg: dict) -> dict:
    th, tw = cfg["sample_height"] // 2, cfg["sample_width"] // 2
    f = 2 ** (len(cfg["block_out_channels"]) - 1)
    lh, lw = int(th / f), int(tw / f)
    oh, ow = 1 / 6, 1 / 5
    return dict(latent_h=lh, latent_w=lw, step_h=int(lh * (1 - oh)), step_w=int(lw * (1 - ow)), blend_h=int(th * oh),
                blend_w=int(tw * ow), limit_h=th - int(th * oh), limit_w=tw - int(tw * ow))


def blend_v(a: Tensor, b: Tensor, extent: int) -> Tensor:
    extent = min(a.shape[3], b.shape[3], extent)
    for y in range(extent):
        b[:, :, :, y, :] = a[:, :, :, -extent + y, :] * (1 - y / extent) + b[:, :, :, y, :] * (y / extent)
    return b


def blend_h(a: Tensor, b: Tensor, extent: int) -> Tensor:
    extent = min(a.shape[4], b.shape[4], extent)
    for x in range(extent):
        b[:, :, :, :, x] = a[:, :, :, :, -extent + x] * (1 - x / extent) + b[:, :, :, :, x] * (x / extent)
    return b


def decode(sd, cfg: dict, z: Tensor, tiling: bool = True) -> Tensor:
    """AutoencoderKLCogVideoX.decode(z).sample for z [B, 16, T, h, w] (slicing = one sample at a time; same result)."""
    g = tile_geometry(cfg)
    B, _, T, h, w = z.shape
    if not (tiling and (w > g["latent_w"] or h > g["latent_h"])):
        return torch.cat([decode_untiled(sd, cfg, z[i:i + 1]) for i in range(B)], dim=0)
    outs = []
    for bi in range(B):
        zb = z[bi:bi + 1]
        rows = []
        for i in range(0, h, g["step_h"]):
            row = []
            for j in range(0, w, g["step_w"]):
                row.append(decode_untiled(sd, cfg, zb[:, :, :, i:i + g["latent_h"], j:j + g["latent_w"]]))
            rows.append(row)
        result_rows = []
        for i, row in enumerate(rows):
            result_row = []
            for j, tile in enumerate(row):
                if i > 0:
                    tile = blend_v(rows[i - 1][j], tile, g["blend_h"])
                if j > 0:
                    tile = blend_h(row[j - 1], tile, g["blend_w"])
                result_row.append(tile[:, :, :, :g["limit_h"], :g["limit_w"]])
            result_rows.append(torch.cat(result_row, dim=4))
        outs.append(torch.cat(result_rows, dim=3))
    return torch.cat(outs, dim=0)
