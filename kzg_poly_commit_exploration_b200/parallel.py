"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The path shards by POINT RANGE for a single MSM (rank g owns bases and scalars
[g n/N, (g+1) n/N)) and by BLOB for batches; the only exchange is an all-gather of one
fixed-size partial-sum blob per rank (192 B XYZZ), folded on the device by kzg_g1_combine_dev."""
from __future__ import annotations


def shard_range(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous range [first, first+count) of rank `rank`; remainders go to the low ranks."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def gather_blobs(blob, world: int):
    """all-gather one equally sized uint8 tensor per rank, rank-major.  `blob` lives on the device the
    process group's backend uses (CUDA for nccl, CPU for gloo)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return blob.clone()
    out = torch.empty(blob.numel() * world, dtype=blob.dtype, device=blob.device)
    dist.all_gather_into_tensor(out, blob.contiguous())
    return out


def exchange_handles(handle: bytes, world: int) -> bytes:
    """all-gather one fixed-size opaque handle per rank (host objects), rank-major"""
    import torch.distributed as dist

    if world == 1:
        return handle
    got = [None] * world
    dist.all_gather_object(got, handle)
    assert all(isinstance(h, bytes) and len(h) == len(handle) for h in got)
    return b"".join(got)


def connect_peers(ctx, world: int, rank: int) -> None:
    """Map every rank's mailbox into this process (CUDA IPC over NVLink) so that kzg_msm_g1_sharded_dev can exchange
    the per-GPU partial sums with plain peer stores: torch.distributed only carries the 64-byte handles, once."""
    import torch.distributed as dist

    handles = exchange_handles(ctx.p2p_init(world, rank), world)
    ctx.p2p_connect(handles)
    if world > 1:
        dist.barrier()
