"""Constructors of diffusers' CogVideoXBlock / CogVideoXTransformer3DModel.  The reference
subclasses both and replaces patch_embed / transformer_blocks / norm_out (cogvideox_control.py:531,554,572), so the
base constructor below does not build the 30 base blocks it would immediately throw away (parameter values are
always loaded from a seeded state dict afterwards, so RNG consumption does not matter)."""
from typing import Optional

import torch
from torch import nn
import torch.nn.functional as F

from ...configuration_utils import ConfigMixin, register_to_config
from ..attention_processor import Attention, CogVideoXAttnProcessor2_0
from ..embeddings import CogVideoXPatchEmbed, TimestepEmbedding, Timesteps
from ..modeling_utils import ModelMixin
from ..normalization import AdaLayerNorm, CogVideoXLayerNormZero


class GELU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int, approximate: str = "none", bias: bool = True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, hidden_states):
        hidden_states = self.proj(hidden_states)
        return F.gelu(hidden_states, approximate=self.approximate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4, dropout: float = 0.0,
                 activation_fn: str = "geglu", final_dropout: bool = False, inner_dim=None, bias: bool = True):
        super().__init__()
        if inner_dim is None:
            inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        if activation_fn == "gelu":
            act_fn = GELU(dim, inner_dim, bias=bias)
        elif activation_fn == "gelu-approximate":
            act_fn = GELU(dim, inner_dim, approximate="tanh", bias=bias)
        else:
            raise NotImplementedError(activation_fn)
        self.net = nn.ModuleList([])
        self.net.append(act_fn)
        self.net.append(nn.Dropout(dropout))
        self.net.append(nn.Linear(inner_dim, dim_out, bias=bias))
        if final_dropout:
            self.net.append(nn.Dropout(dropout))

    def forward(self, hidden_states: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class CogVideoXBlock(nn.Module):
    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, time_embed_dim: int,
                 dropout: float = 0.0, activation_fn: str = "gelu-approximate", attention_bias: bool = False,
                 qk_norm: bool = True, norm_elementwise_affine: bool = True, norm_eps: float = 1e-5,
                 final_dropout: bool = True, ff_inner_dim: Optional[int] = None, ff_bias: bool = True,
                 attention_out_bias: bool = True):
        super().__init__()
        self.norm1 = CogVideoXLayerNormZero(time_embed_dim, dim, norm_elementwise_affine, norm_eps, bias=True)
        self.attn1 = Attention(query_dim=dim, dim_head=attention_head_dim, heads=num_attention_heads,
                               qk_norm="layer_norm" if qk_norm else None, eps=1e-6, bias=attention_bias,
                               out_bias=attention_out_bias, processor=CogVideoXAttnProcessor2_0())
        self.norm2 = CogVideoXLayerNormZero(time_embed_dim, dim, norm_elementwise_affine, norm_eps, bias=True)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn, final_dropout=final_dropout,
                              inner_dim=ff_inner_dim, bias=ff_bias)


class CogVideoXTransformer3DModel(ModelMixin, ConfigMixin):
    _supports_gradient_checkpointing = True

    @register_to_config
    def __init__(self, num_attention_heads: int = 30, attention_head_dim: int = 64, in_channels: int = 16,
                 out_channels: Optional[int] = 16, flip_sin_to_cos: bool = True, freq_shift: int = 0,
                 time_embed_dim: int = 512, ofs_embed_dim: Optional[int] = None, text_embed_dim: int = 4096,
                 num_layers: int = 30, dropout: float = 0.0, attention_bias: bool = True, sample_width: int = 90,
                 sample_height: int = 60, sample_frames: int = 49, patch_size: int = 2,
                 patch_size_t: Optional[int] = None, temporal_compression_ratio: int = 4,
                 max_text_seq_length: int = 226, activation_fn: str = "gelu-approximate",
                 timestep_activation_fn: str = "silu", norm_elementwise_affine: bool = True, norm_eps: float = 1e-5,
                 spatial_interpolation_scale: float = 1.875, temporal_interpolation_scale: float = 1.0,
                 use_rotary_positional_embeddings: bool = False, use_learned_positional_embeddings: bool = False,
                 patch_bias: bool = True, **kwargs):
        super().__init__()
        inner_dim = num_attention_heads * attention_head_dim
        if not use_rotary_positional_embeddings and use_learned_positional_embeddings:
            raise ValueError("There are no CogVideoX checkpoints available with disable rotary embeddings and learned "
                             "positional embeddings.")
        self.patch_embed = None  # replaced by the reference subclass
        self.embedding_dropout = nn.Dropout(dropout)
        self.time_proj = Timesteps(inner_dim, flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(inner_dim, time_embed_dim, timestep_activation_fn)
        self.ofs_proj = None
        self.ofs_embedding = None
        if ofs_embed_dim:
            self.ofs_proj = Timesteps(ofs_embed_dim, flip_sin_to_cos, freq_shift)
            self.ofs_embedding = TimestepEmbedding(ofs_embed_dim, ofs_embed_dim, timestep_activation_fn)
        self.transformer_blocks = nn.ModuleList([])  # replaced by the reference subclass
        self.norm_final = nn.LayerNorm(inner_dim, norm_eps, norm_elementwise_affine)
        self.norm_out = None  # replaced by the reference subclass
        if patch_size_t is None:
            output_dim = patch_size * patch_size * out_channels
        else:
            output_dim = patch_size * patch_size * patch_size_t * out_channels
        self.proj_out = nn.Linear(inner_dim, output_dim)
        self.gradient_checkpointing = False
