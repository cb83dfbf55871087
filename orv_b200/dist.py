"""Multi-GPU plumbing of the path: one process per GPU, clips sharded by rank range, ONE weight broadcast at init,
no per-step collective (reference evaluation_control_to_video.py:118-131, :212-222)."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE (torchrun) and initialises the default process group when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        os.environ.pop("NCCL_P2P_DISABLE", None)  # the reference scripts export it; NVSwitch wants P2P on
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(num_samples: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Rank r takes samples [r*n/N, (r+1)*n/N); the remainder goes to the last rank
    (reference evaluation_control_to_video.py:212-222)."""
    per = num_samples // world_size
    start = rank * per
    end = start + per if rank != world_size - 1 else num_samples
    return start, end


def broadcast_arena(arena: torch.Tensor, src: int = 0) -> torch.Tensor:
    """The single collective of the path: rank `src`'s packed weight arena -> every rank (NCCL over NVLink)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(arena, src=src)
    return arena


def broadcast_weights(model, src: int = 0) -> None:
    """Broadcasts a CogVideoXTransformer3DModelTraj's weight arena.  bf16 parameters are views into the arena, so they
    hold rank `src`'s weights right after the call; the few parameters that cannot alias it (zero-padded matrices,
    non-bf16 tensors) are refreshed from the received arena, so `state_dict()` and a later re-pack agree with it."""
    broadcast_arena(model.weight_arena(), src)
    if hasattr(model, "sync_parameters_from_arena"):
        model.sync_parameters_from_arena()


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
