"""Constructors of CogVideoXLayerNormZero / AdaLayerNorm (their forwards are overridden by the reference)."""
from typing import Optional

import torch
from torch import nn


class CogVideoXLayerNormZero(nn.Module):
    def __init__(self, conditioning_dim: int, embedding_dim: int, elementwise_affine: bool = True, eps: float = 1e-5,
                 bias: bool = True) -> None:
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(conditioning_dim, 6 * embedding_dim, bias=bias)
        self.norm = nn.LayerNorm(embedding_dim, eps=eps, elementwise_affine=elementwise_affine)


class AdaLayerNorm(nn.Module):
    def __init__(self, embedding_dim: int, num_embeddings: Optional[int] = None, output_dim: Optional[int] = None,
                 norm_elementwise_affine: bool = False, norm_eps: float = 1e-5, chunk_dim: int = 0):
        super().__init__()
        self.chunk_dim = chunk_dim
        output_dim = output_dim or embedding_dim * 2
        self.emb = nn.Embedding(num_embeddings, embedding_dim) if num_embeddings is not None else None
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, output_dim)
        self.norm = nn.LayerNorm(output_dim // 2, norm_eps, norm_elementwise_affine)
