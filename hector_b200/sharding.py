"""Multi-GPU plumbing: ensembles shard by member (no exchange during the run); the single
collective of the job is an all-gather of the per-member summary outputs at the end
(SURVEY.md section 8(e)).  Works with NCCL (GPU) and gloo (CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_members, rank, world):
    """contiguous, balanced member range [lo, hi) of `rank`"""
    base, extra = divmod(n_members, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_summary(block, n_members, world):
    """block: [..., members_of_this_rank] tensor (last dim = members, possibly padded beyond the
    rank's share).  Returns [..., n_members] on every rank with members in global order."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(n_members, rank, world)
    mine = block[..., : hi - lo].contiguous()
    if world == 1:
        return mine
    width = shard_range(n_members, 0, world)[1]  # the largest share
    pad = torch.zeros(mine.shape[:-1] + (width,), dtype=mine.dtype, device=mine.device)
    pad[..., : hi - lo] = mine
    flat = torch.empty(world * pad.numel(), dtype=pad.dtype, device=pad.device)
    dist.all_gather_into_tensor(flat, pad.reshape(-1))
    out = flat.view((world,) + tuple(pad.shape))
    parts = []
    for k in range(world):
        a, b = shard_range(n_members, k, world)
        parts.append(out[k][..., : b - a])
    return torch.cat(parts, dim=-1)
