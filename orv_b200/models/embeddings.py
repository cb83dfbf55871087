"""Positional tables of the path, built once per geometry on the host and kept resident in HBM.

These are init-time constants (not per-step work): the additive joint 3-D sin-cos table of CogVideoXPatchEmbed
(2B family), the view table of reference cogvideox_control.py:659-688, and the 3-D rotary cos/sin table of the
1.5-5B family (diffusers get_3d_rotary_pos_embed / orv/utils.py:196-239).  Float64 frequency math as diffusers.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch


def _sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    omega = np.arange(dim // 2, dtype=np.float64) / (dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = np.outer(pos.reshape(-1).astype(np.float64), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_pos_embed_3d(embed_dim: int, width: int, height: int, frames: int, spatial_scale: float,
                        temporal_scale: float) -> torch.Tensor:
    """[frames * height * width, embed_dim] fp32; channel layout [temporal D/4 | spatial(w) 3D/8 | spatial(h) 3D/8]."""
    if embed_dim % 4 != 0:
        raise ValueError("`embed_dim` must be divisible by 4")
    d_sp, d_t = 3 * embed_dim // 4, embed_dim // 4
    gh = np.arange(height, dtype=np.float32) / np.float32(spatial_scale)
    gw = np.arange(width, dtype=np.float32) / np.float32(spatial_scale)
    mw, mh = np.meshgrid(gw, gh)  # "xy" indexing: [height, width]
    sp = np.concatenate([_sincos_1d(d_sp // 2, mw), _sincos_1d(d_sp // 2, mh)], axis=1)  # [h*w, d_sp]
    gt = np.arange(frames, dtype=np.float32) / np.float32(temporal_scale)
    tp = _sincos_1d(d_t, gt)  # [frames, d_t]
    sp = np.broadcast_to(sp[None], (frames, height * width, d_sp))
    tp = np.broadcast_to(tp[:, None], (frames, height * width, d_t))
    return torch.from_numpy(np.concatenate([tp, sp], axis=-1).reshape(frames * height * width, embed_dim)).float()


def _rope_1d(dim: int, pos: torch.Tensor, theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    fr = torch.outer(pos.float(), freqs)
    return fr.cos().repeat_interleave(2, dim=1).float(), fr.sin().repeat_interleave(2, dim=1).float()


def get_3d_rotary_pos_embed(embed_dim: int, crops_coords, grid_size: Tuple[int, int], temporal_size: int,
                            theta: float = 10000.0, use_real: bool = True, grid_type: str = "linspace",
                            max_size: Optional[Tuple[int, int]] = None, device=None):
    """Same signature and result as diffusers.models.embeddings.get_3d_rotary_pos_embed (>= 0.32)."""
    if not use_real:
        raise ValueError("`use_real = False` is not currently supported for get_3d_rotary_pos_embed")
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
        raise ValueError("Invalid value passed for `grid_type`.")
    dim_t, dim_h, dim_w = embed_dim // 4, embed_dim // 8 * 3, embed_dim // 8 * 3
    tc, ts = _rope_1d(dim_t, grid_t, theta)
    hc, hs = _rope_1d(dim_h, grid_h, theta)
    wc, ws = _rope_1d(dim_w, grid_w, theta)
    if grid_type == "slice":
        tc, ts, hc, hs, wc, ws = tc[:temporal_size], ts[:temporal_size], hc[:gh], hs[:gh], wc[:gw], ws[:gw]

    def combine(a, b, c):
        a = a[:, None, None, :].expand(-1, gh, gw, -1)
        b = b[None, :, None, :].expand(temporal_size, -1, gw, -1)
        c = c[None, None, :, :].expand(temporal_size, gh, -1, -1)
        return torch.cat([a, b, c], dim=-1).reshape(temporal_size * gh * gw, -1)

    cos, sin = combine(tc, hc, wc), combine(ts, hs, ws)
    if device is not None:
        if torch.device(device).type == "cuda":  # pinned + non-blocking: a pageable upload would synchronise the stream
            cos, sin = cos.pin_memory().to(device, non_blocking=True), sin.pin_memory().to(device, non_blocking=True)
        else:
            cos, sin = cos.to(device), sin.to(device)
    return cos, sin


def get_resize_crop_region_for_grid(src, tgt_width, tgt_height):
    """Reference orv/utils.py:177-193."""
    tw, th = tgt_width, tgt_height
    h, w = src
    r = h / w
    if r > (th / tw):
        resize_height = th
        resize_width = int(round(th / h * w))
    else:
        resize_width = tw
        resize_height = int(round(tw / w * h))
    crop_top = int(round((th - resize_height) / 2.0))
    crop_left = int(round((tw - resize_width) / 2.0))
    return (crop_top, crop_left), (crop_top + resize_height, crop_left + resize_width)


def timestep_frequencies(dim: int, freq_shift: float) -> torch.Tensor:  # documentation helper
    half = dim // 2
    return torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / (half - freq_shift))
