"""The only collective on this path: an all-gather of the per-rank evaluation counts.

Images are independent, so encode / decode / NMS / TP flagging shard over ranks with no communication; the
per-class (TP, FP, n_gt) int64 table (C x 3) is gathered so every rank can form the global counts.  Uses
torch.distributed (NCCL on GPUs over NVLink/NVSwitch, gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous batch split: items [lo, hi) of rank `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_counts(counts, process_group=None):
    """counts: (C, 3) int64 tensor on this rank -> (world_size, C, 3) tensor with every rank's table."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return counts.unsqueeze(0)
    world = dist.get_world_size(process_group)
    flat = counts.contiguous().view(-1)
    out = torch.empty(world * flat.numel(), dtype=counts.dtype, device=counts.device)  # concatenated form
    dist.all_gather_into_tensor(out, flat, group=process_group)
    return out.view((world,) + tuple(counts.shape))
