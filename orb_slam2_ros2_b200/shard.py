"""Frame sharding for batch sequences (offline map / vocabulary building): frames are independent on this path, so a
sequence of F frames is split into contiguous blocks of ceil(F / R) frames, one block per rank / GPU, with no exchange
during compute.  The only collective is the gather of fixed-stride result records (north_star: "only the descriptor
gather is collected, with NCCL over NVLink"); with the gloo backend the same code runs on CPU tensors (tests).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def frames_per_rank(n_frames: int, world: int) -> int:
    return math.ceil(n_frames / world) if n_frames > 0 else 0


def frame_range(n_frames: int, rank: int, world: int) -> range:
    """contiguous block [rank * ceil(F/R), min(F, (rank + 1) * ceil(F/R)))  """
    per = frames_per_rank(n_frames, world)
    return range(min(n_frames, rank * per), min(n_frames, (rank + 1) * per))


def gather_sequence(local: torch.Tensor, n_frames: int, rank: int, world: int) -> torch.Tensor:
    """All-gather per-frame records.  `local` holds this rank's frames ([len(frame_range), ...], any trailing shape);
    the result holds all F frames in sequence order on every rank.  Ragged last shards are padded to the block size."""
    per = frames_per_rank(n_frames, world)
    mine = len(frame_range(n_frames, rank, world))
    assert local.shape[0] == mine, (local.shape, mine)
    if world == 1:
        return local
    padded = local
    if mine < per:
        pad = torch.zeros((per - mine, *local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], 0)
    out = torch.empty((world * per, *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous())
    return out[:n_frames]
