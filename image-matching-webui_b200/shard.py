"""Pair-stream sharding across ranks: pairs are independent, so rank r takes a contiguous
shard, weights are replicated and the only exchange is one all_gather of per-pair match counts."""
import torch


def shard_range(n_pairs: int, rank: int, world: int):
    """Contiguous, balanced shard [lo, hi) of a stream of n_pairs for `rank` of `world`."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_match_counts(local_counts: torch.Tensor, n_pairs: int, group=None):
    """all_gather of int32 per-pair match counts with ragged shard sizes; returns the [n_pairs] vector on every
    rank.  NCCL on GPUs (NVLink/NVSwitch, latency-bound at this size), gloo in the CPU tests."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_counts
    world = dist.get_world_size(group)
    cap = (n_pairs + world - 1) // world
    buf = torch.zeros(cap, dtype=torch.int32, device=local_counts.device)
    buf[: local_counts.numel()] = local_counts
    out = torch.empty(world * cap, dtype=torch.int32, device=local_counts.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = [out[r * cap: r * cap + (shard_range(n_pairs, r, world)[1] - shard_range(n_pairs, r, world)[0])] for r in range(world)]
    return torch.cat(parts)


def gather_stream_counts(local_counts: torch.Tensor, group=None):
    """The stream's single collective: every rank contributes the int32 match counts of all pairs it processed
    (equal-sized shards: steps x pairs_per_step), every rank receives the [world * n] vector.  No-op without a process
    group (single GPU)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_counts
    world = dist.get_world_size(group)
    out = torch.empty(world * local_counts.numel(), dtype=local_counts.dtype, device=local_counts.device)
    dist.all_gather_into_tensor(out, local_counts.contiguous(), group=group)
    return out
