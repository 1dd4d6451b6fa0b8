"""Multi-GPU plumbing: one process per GPU, rows of X_all in contiguous blocks, torch.distributed for the
rendezvous (NCCL unique id, barriers, max-over-ranks timing); the data-path collectives are issued by the
CUDA library itself on its own stream (int64 limb all-reduces, fp64 K x K all-reduce)."""
from __future__ import annotations

import os

from .datasets import shard_bounds  # noqa: F401


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def exchange_unique_id(make_id, rank: int, group=None) -> bytes:
    """Rank 0 creates the 128-byte communicator id; everyone receives it through torch.distributed."""
    import torch.distributed as dist

    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def init_context_comm(ctx, rank: int, world: int, group=None):
    """Create the library's NCCL communicator for this rank (no-op for world == 1)."""
    if world == 1:
        ctx.comm_init(None, 0, 1)
        return
    from .api import Context

    uid = exchange_unique_id(Context.comm_unique_id, rank, group)
    ctx.comm_init(uid, rank, world)
