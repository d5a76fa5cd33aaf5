"""Data-parallel plumbing (the reference is single-device, const.py:7).

Global batch -> contiguous per-rank shards (rank r gets rows [r*N/W, (r+1)*N/W)); parameters, momentum slots and BN
moving statistics are replicated (identical initial values by shared seed or broadcast); BatchNorm uses LOCAL
(per-replica) batch statistics; one all-reduce(sum) of the flat fp32 gradient per step, scaled by 1/world inside
the momentum kernel.  torch.distributed is only the rendezvous / unique-id carrier; the all-reduce itself is issued by
libtcr_b200 on its own NCCL communicator (csrc/tcr_comm.cu)."""
from __future__ import annotations

from typing import Tuple


def shard_bounds(n_global: int, rank: int, world: int) -> Tuple[int, int]:
    if n_global % world:
        raise ValueError(f"global batch {n_global} is not divisible by world size {world}")
    per = n_global // world
    return rank * per, (rank + 1) * per


def broadcast_bytes(payload: bytes | None, src: int = 0) -> bytes:
    """Ship a small byte string (the 128-byte ncclUniqueId) from `src` to every rank."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def broadcast_variables(*tensors, src: int = 0):
    """Make replicas identical (e.g. after loading a checkpoint on rank 0)."""
    import torch.distributed as dist
    for t in tensors:
        dist.broadcast(t, src=src)
