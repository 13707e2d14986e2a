"""Frame sharding across the GPUs of one box.

Video frames are independent units: GPU r of G owns the contiguous frame range
[r*ceil(N/G), min(N,(r+1)*ceil(N/G))), holds a full weight replica, computes the whole mel locally
(cheaper than slicing audio with halos, and bit-identical) and gathers only its own windows.  There is
NO collective on the compute path; the only exchange is one final gather of the generated frames
(NCCL all_gather of equally padded shards over NVLink 5 / NVSwitch).  One process per GPU,
``torch.distributed`` (nccl on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    per = -(-n_frames // world) if n_frames > 0 else 0
    lo = min(n_frames, rank * per)
    return lo, min(n_frames, lo + per)


def gather_frames(local: torch.Tensor, n_frames: int, group=None) -> torch.Tensor:
    """local: this rank's frames [n_r, ...] (n_r from shard_range) -> all frames [n_frames, ...] on every
    rank, in frame order.  Shards are padded to ceil(N/G) so a single all_gather moves everything."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == n_frames
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-n_frames // world)
    lo, hi = shard_range(n_frames, rank, world)
    assert local.shape[0] == hi - lo, (local.shape, lo, hi)
    padded = local.new_zeros((per,) + tuple(local.shape[1:]))
    padded[: hi - lo] = local
    out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    return out[:n_frames]
