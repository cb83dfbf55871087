"""B200 decode path of the CogVideoX 3-D VAE.

Mirror of the part of diffusers' `AutoencoderKLCogVideoX` that the reference pipeline executes: `decode()` reached from
`decode_latents` (reference orv/models/cogvideox_control.py:1095-1100, :1476-1479) with `enable_slicing()` /
`enable_tiling()` switched on by orv/pipeline/inference_control_to_video.py:98-99.  Same constructor keys, same
`state_dict` key names for the decoder half (so a THUDM/CogVideoX-2b `vae/` checkpoint loads), same `.config` attributes
the sampler reads, same tiling / frame-batch / convolution-cache behaviour.  The encoder is offline tooling in the
reference (`encode_dataset.py`, SURVEY §2 out of scope) and is not built: `encode()` raises.

Arithmetic: liborv_b200.so only.  Activations live channels-last ([T, H, W, C] bf16, one sample — the reference decodes
sample by sample as well: slicing); every causal convolution is an implicit GEMM on tcgen05 (`orvb_conv_cl`: the 3x3x3
taps are TMA boxes of the activation tensor shifted by the tap offset, zero-filled at the borders; its epilogue also
accumulates the GroupNorm statistics of its output), every SpatialNorm3D + SiLU is one fused pass
(`orvb_spatial_norm_cl`; `orvb_gn_stats_cl` only where no convolution produced the tensor) whose conv_y / conv_b branches are
evaluated once per LATENT pixel for all 37 norm sites of a frame batch by a single GEMM and looked up through the
nearest-neighbour map, and the residual add rides in the second convolution's epilogue.  What stays in torch is layout
plumbing on tiny tensors (latent tile -> channels-last, tile blending / concatenation of the decoded frames).
There is no CPU or eager fallback: without the library or on a non-B200 device `decode` raises.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import nn

from .. import ops


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class _Config(dict):
    """dict with attribute access (diffusers FrozenDict as far as the sampler uses it)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v


def _register(root: nn.Module, dotted: str, tensor: torch.Tensor) -> None:
    parts = dotted.split(".")
    m = root
    for p in parts[:-1]:
        if p not in m._modules:
            m.add_module(p, nn.Module())
        m = m._modules[p]
    m.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


def _nearest_map(n_out: int, n_in: int) -> List[int]:
    """Source index of F.interpolate(mode='nearest', size=n_out) along one axis (float32 scale, as ATen computes it)."""
    scale = np.float32(n_in) / np.float32(n_out)
    return [min(int(math.floor(np.float32(i) * scale)), n_in - 1) for i in range(n_out)]


def spatial_norm_frame_map(t_out: int, t_lat: int) -> List[int]:
    """CogVideoXSpatialNorm3D: latent frame each feature frame reads (first frame resized on its own when the frame
    count is odd and > 1)."""
    if t_out > 1 and t_out % 2 == 1:
        return [0] + [1 + i for i in _nearest_map(t_out - 1, t_lat - 1)]
    return _nearest_map(t_out, t_lat)


def upsample_frame_map(t_in: int, compress_time: bool) -> List[int]:
    """CogVideoXUpsample3D: input frame of every output frame."""
    if not compress_time or t_in == 1:
        return list(range(t_in))
    if t_in % 2 == 1:
        return [0] + [1 + i // 2 for i in range(2 * (t_in - 1))]
    return [i // 2 for i in range(2 * t_in)]


def frame_batches(num_frames: int, batch: int = 2) -> List[Tuple[int, int]]:
    """AutoencoderKLCogVideoX._decode: latent frames are decoded `batch` at a time, the first batch takes the remainder."""
    n = max(num_frames // batch, 1)
    rem = num_frames % batch
    out = []
    for i in range(n):
        start = batch * i + (0 if i == 0 else rem)
        end = batch * (i + 1) + rem
        out.append((start, min(end, num_frames)))
    return out


class AutoencoderKLCogVideoX(nn.Module):
    config_name = "config.json"

    def __init__(self, in_channels: int = 3, out_channels: int = 3,
                 down_block_types=("CogVideoXDownBlock3D",) * 4, up_block_types=("CogVideoXUpBlock3D",) * 4,
                 block_out_channels=(128, 256, 256, 512), latent_channels: int = 16, layers_per_block: int = 3,
                 act_fn: str = "silu", norm_eps: float = 1e-6, norm_num_groups: int = 32,
                 temporal_compression_ratio: float = 4, sample_height: int = 480, sample_width: int = 720,
                 scaling_factor: float = 1.15258426, shift_factor=None, latents_mean=None, latents_std=None,
                 force_upcast: bool = True, use_quant_conv: bool = False, use_post_quant_conv: bool = False,
                 invert_scale_latents: bool = False):
        super().__init__()
        if act_fn != "silu":
            raise NotImplementedError("only act_fn='silu' (every released CogVideoX VAE)")
        if use_post_quant_conv:
            raise NotImplementedError("use_post_quant_conv=True is not used by any released CogVideoX VAE")
        block_out_channels = tuple(block_out_channels)
        self.config = _Config(
            in_channels=in_channels, out_channels=out_channels, down_block_types=tuple(down_block_types),
            up_block_types=tuple(up_block_types), block_out_channels=block_out_channels, latent_channels=latent_channels,
            layers_per_block=layers_per_block, act_fn=act_fn, norm_eps=norm_eps, norm_num_groups=norm_num_groups,
            temporal_compression_ratio=temporal_compression_ratio, sample_height=sample_height, sample_width=sample_width,
            scaling_factor=scaling_factor, shift_factor=shift_factor, latents_mean=latents_mean, latents_std=latents_std,
            force_upcast=force_upcast, use_quant_conv=use_quant_conv, use_post_quant_conv=use_post_quant_conv,
            invert_scale_latents=invert_scale_latents)
        self.use_slicing = False
        self.use_tiling = False
        self.num_latent_frames_batch_size = 2
        self.num_sample_frames_batch_size = 8
        # tile geometry exactly as diffusers derives it (autoencoder_kl_cogvideox.py, __init__)
        self.tile_sample_min_height = sample_height // 2
        self.tile_sample_min_width = sample_width // 2
        f = 2 ** (len(block_out_channels) - 1)
        self.tile_latent_min_height = int(self.tile_sample_min_height / f)
        self.tile_latent_min_width = int(self.tile_sample_min_width / f)
        self.tile_overlap_factor_height = 1 / 6
        self.tile_overlap_factor_width = 1 / 5
        for name, shape in self._decoder_shapes().items():
            _register(self, name, torch.zeros(shape))
        self._packed: Optional[dict] = None
        self._maps: Dict[Tuple, torch.Tensor] = {}
        self.last_launches = 0

    # ------------------------------------------------------------------------------------------------------
    # parameters (diffusers names)
    # ------------------------------------------------------------------------------------------------------
    def _resnet_list(self) -> List[Tuple[str, int, int]]:
        """(name, c_in, c_out) of every CogVideoXResnetBlock3D of the decoder, in execution order."""
        rev = tuple(reversed(self.config.block_out_channels))
        out = [(f"decoder.mid_block.resnets.{i}", rev[0], rev[0]) for i in range(2)]
        cout = rev[0]
        for b, ch in enumerate(rev):
            cin, cout = cout, ch
            for i in range(self.config.layers_per_block + 1):
                out.append((f"decoder.up_blocks.{b}.resnets.{i}", cin if i == 0 else cout, cout))
        return out

    def _norm_sites(self) -> List[Tuple[str, int]]:
        sites = []
        for name, cin, cout in self._resnet_list():
            sites += [(f"{name}.norm1", cin), (f"{name}.norm2", cout)]
        sites.append(("decoder.norm_out", self.config.block_out_channels[0]))
        return sites

    def _decoder_shapes(self) -> Dict[str, Tuple[int, ...]]:
        c = self.config
        zc = c.latent_channels
        rev = tuple(reversed(c.block_out_channels))
        s: Dict[str, Tuple[int, ...]] = {}

        def conv3(name, cin, cout, k):
            s[f"{name}.conv.weight"] = (cout, cin, k, k, k)
            s[f"{name}.conv.bias"] = (cout,)

        conv3("decoder.conv_in", zc, rev[0], 3)
        for name, cin, cout in self._resnet_list():
            for nn_, ch in ((f"{name}.norm1", cin), (f"{name}.norm2", cout)):
                s[f"{nn_}.norm_layer.weight"] = (ch,)
                s[f"{nn_}.norm_layer.bias"] = (ch,)
                conv3(f"{nn_}.conv_y", zc, ch, 1)
                conv3(f"{nn_}.conv_b", zc, ch, 1)
            conv3(f"{name}.conv1", cin, cout, 3)
            conv3(f"{name}.conv2", cout, cout, 3)
            if cin != cout:
                s[f"{name}.conv_shortcut.weight"] = (cout, cin, 1, 1, 1)
                s[f"{name}.conv_shortcut.bias"] = (cout,)
        for b, ch in enumerate(rev[:-1]):
            s[f"decoder.up_blocks.{b}.upsamplers.0.conv.weight"] = (ch, ch, 3, 3)
            s[f"decoder.up_blocks.{b}.upsamplers.0.conv.bias"] = (ch,)
        s["decoder.norm_out.norm_layer.weight"] = (rev[-1],)
        s["decoder.norm_out.norm_layer.bias"] = (rev[-1],)
        conv3("decoder.norm_out.conv_y", zc, rev[-1], 1)
        conv3("decoder.norm_out.conv_b", zc, rev[-1], 1)
        conv3("decoder.conv_out", rev[-1], c.out_channels, 3)
        return s

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Encoder / quant-conv keys of a full checkpoint are ignored (decode-only module)."""
        sd = {k: v for k, v in state_dict.items() if k.startswith("decoder.")}
        self._packed = None
        return super().load_state_dict(sd, strict=strict, assign=assign)

    def _apply(self, fn, *a, **k):
        self._packed = None
        self._maps = {}
        return super()._apply(fn, *a, **k)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None, torch_dtype=None, **kwargs):
        from safetensors.torch import load_file
        d = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        with open(os.path.join(d, cls.config_name), "r", encoding="utf-8") as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        cfg.update({k: v for k, v in kwargs.items() if k in cfg})
        model = cls(**cfg)
        sd: Dict[str, torch.Tensor] = {}
        for fn in sorted(x for x in os.listdir(d) if x.endswith(".safetensors")):
            sd.update(load_file(os.path.join(d, fn)))
        if not sd:
            raise FileNotFoundError(f"no .safetensors weights under {d}")
        model.load_state_dict(sd, strict=True)
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        return model.eval()

    # diffusers switches the reference scripts call (inference_control_to_video.py:98-99)
    def enable_tiling(self, tile_sample_min_height=None, tile_sample_min_width=None, tile_overlap_factor_height=None,
                      tile_overlap_factor_width=None):
        self.use_tiling = True
        self.tile_sample_min_height = tile_sample_min_height or self.tile_sample_min_height
        self.tile_sample_min_width = tile_sample_min_width or self.tile_sample_min_width
        f = 2 ** (len(self.config.block_out_channels) - 1)
        self.tile_latent_min_height = int(self.tile_sample_min_height / f)
        self.tile_latent_min_width = int(self.tile_sample_min_width / f)
        self.tile_overlap_factor_height = tile_overlap_factor_height or self.tile_overlap_factor_height
        self.tile_overlap_factor_width = tile_overlap_factor_width or self.tile_overlap_factor_width

    def disable_tiling(self):
        self.use_tiling = False

    def enable_slicing(self):
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    def encode(self, *_, **__):
        raise NotImplementedError("the VAE encoder is offline tooling in the reference (encode_dataset.py); only decode "
                                  "is on the inference path")

    # ------------------------------------------------------------------------------------------------------
    # weight packing: K-major bf16 matrices for the implicit GEMM
    # ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _pack_conv(w: torch.Tensor, b: Optional[torch.Tensor], dev) -> Tuple[torch.Tensor, torch.Tensor, Tuple[int, int, int]]:
        """[c_out, c_in, (kt,) kh, kw] -> [c_out8, kt*kh*kw*c_in64] (tap-major, channels padded with zeros)."""
        if w.dim() == 4:
            w = w[:, :, None]
        cout, cin, kt, kh, kw = w.shape
        cin_p, cout_p = -(-cin // 64) * 64, -(-cout // 8) * 8
        m = torch.zeros((cout_p, kt, kh, kw, cin_p), dtype=torch.bfloat16, device=dev)
        m[:cout, ..., :cin] = w.permute(0, 2, 3, 4, 1).to(device=dev, dtype=torch.bfloat16)
        bias = torch.zeros((cout_p,), dtype=torch.bfloat16, device=dev)
        if b is not None:
            bias[:cout] = b.to(device=dev, dtype=torch.bfloat16)
        return m.reshape(cout_p, -1).contiguous(), bias, (kt, kh, kw)

    def _pack(self) -> dict:
        if self._packed is not None:
            return self._packed
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("AutoencoderKLCogVideoX.decode needs the module on a CUDA device (no CPU path)")
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        P: dict = {"conv": {}, "norm": {}}
        for k in sd:
            if k.endswith(".weight") and (k.endswith("conv.weight") or k.endswith("conv_shortcut.weight")) \
                    and ".conv_y." not in k and ".conv_b." not in k:
                base = k[:-len(".weight")]
                name = base[:-len(".conv")] if base.endswith(".conv") else base  # CausalConv3d wraps its conv
                P["conv"][name] = self._pack_conv(sd[k], sd.get(base + ".bias"), dev)
        # one [N_total, 64] matrix for the conv_y / conv_b branches of every norm site: table = zq64 @ W^T + b
        zc = self.config.latent_channels
        rows, biases, off = [], [], 0
        for name, ch in self._norm_sites():
            ent = {"gamma": sd[f"{name}.norm_layer.weight"].to(dev, torch.bfloat16).contiguous(),
                   "beta": sd[f"{name}.norm_layer.bias"].to(dev, torch.bfloat16).contiguous(), "channels": ch}
            for tag in ("y", "b"):
                w = sd[f"{name}.conv_{tag}.conv.weight"].reshape(ch, zc)
                rows.append(torch.nn.functional.pad(w, (0, 64 - zc)))
                biases.append(sd[f"{name}.conv_{tag}.conv.bias"])
                ent[f"{tag}_off"] = off
                off += ch
            P["norm"][name] = ent
        P["table_w"] = torch.cat(rows, 0).to(dev, torch.bfloat16).contiguous()
        P["table_b"] = torch.cat(biases, 0).to(dev, torch.bfloat16).contiguous()
        self._packed = P
        return P

    def _map(self, idx: List[int]) -> torch.Tensor:
        key = tuple(idx)
        t = self._maps.get(key)
        if t is None:
            t = torch.tensor(idx, dtype=torch.int32, device=self.device)
            self._maps[key] = t
        return t

    # ------------------------------------------------------------------------------------------------------
    # decoder (one tile of one sample, one frame batch)
    # ------------------------------------------------------------------------------------------------------
    def _conv(self, P, name, x, cache: dict, resid=None, stats: bool = False):
        """One convolution; stats=True also returns the GroupNorm statistics of its output (accumulated in the epilogue:
        the SpatialNorm3D that follows needs no pass of its own over the tensor)."""
        w, b, ker = P["conv"][name]
        kt = ker[0]
        prev = cache.get(name) if kt > 1 else None
        fuse = stats and os.environ.get("ORVB_VAE_FUSE_GN", "1") != "0" and w.shape[0] % 64 == 0 and (w.shape[0] // self.config.norm_num_groups) in (2, 4, 8, 16, 32, 64) \
            and self.config.norm_num_groups <= 32
        y = ops.conv_cl(x, w, b, ker, cache=prev, resid=resid, gn=(self.config.norm_num_groups, 1e-6) if fuse else None)
        self.last_launches += 2 if fuse else 1
        if kt > 1:
            # CogVideoXCausalConv3d: the cache is the last kt-1 frames of [context | x]
            if x.shape[0] >= kt - 1:
                cache[name] = x[-(kt - 1):].clone()
            else:
                ctx = prev if prev is not None else x[:1].expand(kt - 1, *x.shape[1:])
                cache[name] = torch.cat([ctx, x], 0)[-(kt - 1):].contiguous()
        if stats:
            return y if fuse else (y, None)
        return y

    def _norm_act(self, P, name, x, table, t_lat, lat_hw, act=1, stats=None):
        ent = P["norm"][name]
        g = self.config.norm_num_groups
        if stats is None:
            stats = ops.gn_stats_cl(x, g, 1e-6)
            self.last_launches += 1
        T, H = x.shape[0], x.shape[1]
        shift = int(round(math.log2(H / lat_hw[0]))) if H >= lat_hw[0] else 0
        t_src = self._map(spatial_norm_frame_map(T, t_lat))
        self.last_launches += 1
        return ops.spatial_norm_cl(x, stats, ent["gamma"], ent["beta"], table, ent["y_off"], ent["b_off"], t_src, lat_hw,
                                   shift, groups=g, act=act)

    def _decoder(self, z: torch.Tensor, cache: dict) -> torch.Tensor:
        """CogVideoXDecoder3D.forward.  z [16, Tb, h, w] (one sample, one frame batch) -> [3, T', 8h, 8w]."""
        P = self._pack()
        c = self.config
        zc, Tb, h, w = z.shape
        z64 = torch.zeros((Tb, h, w, 64), dtype=torch.bfloat16, device=z.device)
        z64[..., :zc] = z.permute(1, 2, 3, 0)
        table = ops.gemm(z64.view(-1, 64), P["table_w"], P["table_b"])
        self.last_launches += 1
        lat = (h, w)
        x, xs = self._conv(P, "decoder.conv_in", z64, cache, stats=True)
        rev = tuple(reversed(c.block_out_channels))
        compress_level = int(np.log2(c.temporal_compression_ratio))

        def resnet(name, cin, cout, x, xs, want_stats=True):
            hdn = self._norm_act(P, f"{name}.norm1", x, table, Tb, lat, stats=xs)
            hdn, hs = self._conv(P, f"{name}.conv1", hdn, cache, stats=True)
            hdn = self._norm_act(P, f"{name}.norm2", hdn, table, Tb, lat, stats=hs)
            if cin != cout:
                x = self._conv(P, f"{name}.conv_shortcut", x, cache)
            if want_stats:
                return self._conv(P, f"{name}.conv2", hdn, cache, resid=x, stats=True)
            return self._conv(P, f"{name}.conv2", hdn, cache, resid=x), None

        for i in range(2):
            x, xs = resnet(f"decoder.mid_block.resnets.{i}", rev[0], rev[0], x, xs)
        cout = rev[0]
        for b, ch in enumerate(rev):
            cin, cout = cout, ch
            last_block = b == len(rev) - 1
            for i in range(c.layers_per_block + 1):
                # the last resnet of a block feeds the upsampler, which has no norm: no statistics needed there
                need = last_block or i < c.layers_per_block
                x, xs = resnet(f"decoder.up_blocks.{b}.resnets.{i}", cin if i == 0 else cout, cout, x, xs, want_stats=need)
            if not last_block:
                x = ops.upsample2x_cl(x, self._map(upsample_frame_map(x.shape[0], b < compress_level)))
                self.last_launches += 1
                x, xs = self._conv(P, f"decoder.up_blocks.{b}.upsamplers.0", x, cache, stats=True)
        x = self._norm_act(P, "decoder.norm_out", x, table, Tb, lat, stats=xs)
        x = self._conv(P, "decoder.conv_out", x, cache)
        self.last_launches += 1
        return ops.cl_to_planar(x, c.out_channels)

    def _decode_untiled(self, z: torch.Tensor) -> torch.Tensor:
        """z [16, T, h, w] -> [3, T', H, W]: frame batches sharing convolution caches (AutoencoderKLCogVideoX._decode)."""
        cache: dict = {}
        outs = [self._decoder(z[:, s:e], cache) for s, e in frame_batches(z.shape[1], self.num_latent_frames_batch_size)]
        return torch.cat(outs, dim=1)

    # ------------------------------------------------------------------------------------------------------
    # tiling (AutoencoderKLCogVideoX.tiled_decode / blend_v / blend_h)
    # ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _blend(a: torch.Tensor, b: torch.Tensor, extent: int, dim: int) -> torch.Tensor:
        """b[..., y, ...] = a[..., -extent + y, ...] * (1 - y / extent) + b[..., y, ...] * (y / extent) for y < extent along
        `dim`, in place on b — the reference's per-line loop as one vector expression with the same rounding (python
        scalar * bf16 tensor = fp32 product rounded to bf16, then a bf16 add)."""
        extent = min(a.shape[dim], b.shape[dim], extent)
        if extent <= 0:
            return b
        wgt = torch.tensor([y / extent for y in range(extent)], dtype=torch.float64, device=b.device)
        w_b = wgt.to(torch.float32)
        w_a = (1.0 - wgt).to(torch.float32)
        shape = [1] * b.dim()
        shape[dim] = extent
        sa = a.narrow(dim, a.shape[dim] - extent, extent)
        sb = b.narrow(dim, 0, extent)
        pa = (sa.float() * w_a.view(shape)).to(b.dtype)
        pb = (sb.float() * w_b.view(shape)).to(b.dtype)
        sb.copy_(pa + pb)
        return b

    def _tiled_decode(self, z: torch.Tensor) -> torch.Tensor:
        """z [16, T, h, w] (one sample) -> [3, T', 8h, 8w]."""
        _, _, h, w = z.shape
        oh = int(self.tile_latent_min_height * (1 - self.tile_overlap_factor_height))
        ow = int(self.tile_latent_min_width * (1 - self.tile_overlap_factor_width))
        bh = int(self.tile_sample_min_height * self.tile_overlap_factor_height)
        bw = int(self.tile_sample_min_width * self.tile_overlap_factor_width)
        lim_h = self.tile_sample_min_height - bh
        lim_w = self.tile_sample_min_width - bw
        rows = []
        for i in range(0, h, oh):
            row = []
            for j in range(0, w, ow):
                tile = z[:, :, i:i + self.tile_latent_min_height, j:j + self.tile_latent_min_width]
                row.append(self._decode_untiled(tile))
            rows.append(row)
        result_rows = []
        for i, row in enumerate(rows):
            result_row = []
            for j, tile in enumerate(row):
                if i > 0:
                    tile = self._blend(rows[i - 1][j], tile, bh, 2)
                if j > 0:
                    tile = self._blend(row[j - 1], tile, bw, 3)
                result_row.append(tile[:, :, :lim_h, :lim_w])
            result_rows.append(torch.cat(result_row, dim=3))
        return torch.cat(result_rows, dim=2)

    def decode_conv_flops(self, frames: int, h: int, w: int) -> float:
        """Algorithmic FLOP (2 * pixels * taps * c_in * c_out) of the convolutions of one sample's decode, tile overlap
        included — it is work the reference's algorithm prescribes.  Used by the benches for the tensor-pipe fraction."""
        c = self.config
        rev = tuple(reversed(c.block_out_channels))
        if self.use_tiling and (w > self.tile_latent_min_width or h > self.tile_latent_min_height):
            oh = int(self.tile_latent_min_height * (1 - self.tile_overlap_factor_height))
            ow = int(self.tile_latent_min_width * (1 - self.tile_overlap_factor_width))
            tiles = [(min(self.tile_latent_min_height, h - i), min(self.tile_latent_min_width, w - j))
                     for i in range(0, h, oh) for j in range(0, w, ow)]
        else:
            tiles = [(h, w)]
        compress_level = int(np.log2(c.temporal_compression_ratio))

        def res(ci, co):
            return 27 * ci * co + 27 * co * co + (ci * co if ci != co else 0)

        total = 0.0
        for th, tw in tiles:
            for s, e in frame_batches(frames, self.num_latent_frames_batch_size):
                t, H, W = e - s, th, tw
                total += 2.0 * t * H * W * (27 * c.latent_channels * rev[0] + 2 * res(rev[0], rev[0]))
                cout = rev[0]
                for b, ch in enumerate(rev):
                    cin, cout = cout, ch
                    total += 2.0 * t * H * W * (res(cin, cout) + c.layers_per_block * res(cout, cout))
                    if b != len(rev) - 1:
                        t = len(upsample_frame_map(t, b < compress_level))
                        H, W = 2 * H, 2 * W
                        total += 2.0 * t * H * W * 9 * cout * cout
                total += 2.0 * t * H * W * 27 * cout * c.out_channels
        return total

    def _decode_sample(self, z: torch.Tensor) -> torch.Tensor:
        _, _, h, w = z.shape
        if self.use_tiling and (w > self.tile_latent_min_width or h > self.tile_latent_min_height):
            return self._tiled_decode(z)
        return self._decode_untiled(z)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """z [B, latent_channels, T, h, w] -> sample [B, 3, 4(T-1)+1, 8h, 8w] in z's dtype."""
        if not z.is_cuda:
            raise RuntimeError("AutoencoderKLCogVideoX.decode: CUDA tensors only (liborv_b200 has no CPU path)")
        if z.dim() != 5 or z.shape[1] != self.config.latent_channels:
            raise ValueError(f"decode expects [B, {self.config.latent_channels}, T, h, w], got {tuple(z.shape)}")
        self.last_launches = 0
        zb = z.to(torch.bfloat16)
        # the reference decodes one sample at a time (enable_slicing); without slicing the per-sample results are the
        # same because no operator of the decoder mixes samples
        dec = torch.stack([self._decode_sample(zb[i]) for i in range(zb.shape[0])], 0).to(z.dtype)
        if not return_dict:
            return (dec,)
        return DecoderOutput(sample=dec)

    def forward(self, *_, **__):
        raise NotImplementedError("training-time forward (encode + decode) is out of scope; call decode()")


def default_vae_namespace(scaling_factor: float = 1.15258426, invert_scale_latents: bool = False) -> SimpleNamespace:
    """Config-only stand-in (no weights) for latent-in / latent-out runs."""
    return SimpleNamespace(config=SimpleNamespace(
        block_out_channels=(128, 256, 256, 512), temporal_compression_ratio=4, scaling_factor=scaling_factor,
        latent_channels=16, invert_scale_latents=invert_scale_latents))
