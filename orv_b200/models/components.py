"""Host-side mirrors of reference orv/models/components.py for the pieces that sit on the denoising path.

`ActionEmbed` keeps the reference's parameter names (`mlp.0`, `mlp.3`, `mask_embed`) and its host-visible
behaviour (shape checks, the per-call `torch.rand(B) < 0.1` draw, the `mask` flag); the MLP itself runs inside
liborv_b200 (`orvb_forward`), so this module only owns parameters and integer bookkeeping.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import nn


@dataclass
class Transformer2DModelOutput:
    sample: torch.Tensor


@dataclass
class Transformer3DModelTrajOutput(Transformer2DModelOutput):
    """Reference components.py:13-17."""
    is_action_mask: Optional[torch.Tensor]
    actions_recon: Optional[torch.Tensor]


class ActionEmbed(nn.Module):
    """Parameter container + input bookkeeping of reference components.py:20-71."""

    def __init__(self, state_dim: int, hidden_size: int, dropout: float = 0.0, compress_ratio: int = 1,
                 patch_size_t: Optional[int] = None, mask: Optional[bool] = False) -> None:
        super().__init__()
        self.state_dim = state_dim
        self.compress_ratio = compress_ratio
        self.patch_size_t = patch_size_t or 1
        self.mask = mask
        self.mlp = nn.Sequential(
            nn.Linear(state_dim * compress_ratio * self.patch_size_t, hidden_size * 4, bias=True),
            nn.GELU(approximate="tanh"),
            nn.Dropout(dropout),
            nn.Linear(hidden_size * 4, hidden_size, bias=True),
            nn.Dropout(dropout),
        )
        self.mask_embed = nn.Embedding(num_embeddings=1, embedding_dim=hidden_size)

    def mlp_input(self, x: torch.Tensor) -> torch.Tensor:
        """The integer bookkeeping of `forward` in front of the MLP (components.py:47-62): first-frame padding and the
        compress / patch_size_t reshapes.  [B, F, state] -> [B, F', state*compress*pt]."""
        B, Fr, state_dim = x.shape
        if state_dim != self.state_dim:
            raise ValueError(f"Got mismatched {x.shape=} and {self.state_dim=}.")
        x = torch.cat([torch.zeros_like(x[:, :1, ...]), x], dim=1)  # pad the first frame
        if self.compress_ratio > 1:
            x = x.reshape(B, (Fr + 1) // self.compress_ratio, -1)
        if self.patch_size_t > 1:
            _, Fr2, _ = x.shape
            x = x.reshape(B, Fr2 // self.patch_size_t, -1)
        return x.contiguous()

    def prepare(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Everything of `forward` (components.py:47-71) except the MLP arithmetic.

        Returns (mlp_input [B, F', state*compress*pt], is_mask [B] bool, apply_mask [B] uint8)."""
        x = self.mlp_input(x)
        is_mask = torch.rand(x.shape[0], device=x.device) < 0.1  # drawn on every call, as the reference does (:66)
        apply = is_mask if self.mask else torch.zeros_like(is_mask)
        return x, is_mask, apply.to(torch.uint8)

    def forward(self, x):  # pragma: no cover - the arithmetic lives in liborv_b200
        raise RuntimeError("ActionEmbed runs inside orvb_forward; call the transformer, not this module")


class ActionRecon(nn.Module):
    """Train-only head (reference components.py:74-104); kept so checkpoints with recon_action=True load."""

    def __init__(self, state_dim: int, hidden_size: int, compress_ratio: int = 1) -> None:
        super().__init__()
        self.state_dim = state_dim
        self.compress_ratio = compress_ratio
        self.mlp = nn.Sequential(
            nn.Linear(hidden_size, hidden_size * 4, bias=True),
            nn.GELU(approximate="tanh"),
            nn.Linear(hidden_size * 4, state_dim * compress_ratio, bias=True),
        )
