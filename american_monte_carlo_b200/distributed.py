"""Multi-GPU plumbing: one process per GPU (torchrun), paths sharded along the path axis.

torch.distributed is used ONLY for the rendezvous (broadcasting the 128-byte NCCL unique id and, in tests, for
gloo collectives on the CPU); the per-step all-reduce of the moment sums runs inside libamc on its own NCCL
communicator, on the same CUDA stream as the kernels, with no host round trip.
"""
from __future__ import annotations

import os

from .api import Context, set_default_context, shard_range  # noqa: F401


def exchange_unique_id(make_id, rank: int, src: int = 0) -> bytes:
    """Rank `src` calls make_id() (-> 128 bytes); every rank returns the same bytes.  Needs an initialised
    torch.distributed process group (any backend)."""
    import torch.distributed as dist
    box = [make_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("NCCL unique id exchange failed")
    return bytes(uid)


def init_distributed(stream=None, make_default=True) -> Context:
    """Create this rank's Context on cuda:LOCAL_RANK and join the libamc communicator.

    Call after torch.distributed.init_process_group(...).  With world size 1 no communicator is created.
    """
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    ctx = Context(int(os.environ.get("LOCAL_RANK", "0")), stream=stream)
    if world > 1:
        uid = exchange_unique_id(Context.new_unique_id, rank)
        ctx.init_comm(world, rank, uid)
    if make_default:
        set_default_context(ctx)
    return ctx
