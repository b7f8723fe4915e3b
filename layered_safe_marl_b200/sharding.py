"""Multi-GPU: environments are independent, so the path shards with NO data-path collective.
Rank g owns the contiguous env range returned by `shard_range`; the only collective is the
all-reduce of the 8 episode-summary sums (+ a count) when the runner logs."""
from __future__ import annotations

import torch

from . import layout as LY


def shard_range(num_envs_total: int, world_size: int, rank: int):
    """Contiguous, balanced split: (first_env, num_envs) of `rank`. first_env is the env_id_base that
    keys the reset RNG stream, so a sharded run draws the same scenarios as a single-GPU run."""
    base, rem = divmod(int(num_envs_total), int(world_size))
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def allreduce_episode_stats(ep_info: torch.Tensor, group=None, sums: torch.Tensor = None) -> dict:
    """ep_info: (n_local, 8) per-env episode summaries. Returns the mean over ALL shards.
    `sums`: the 9 doubles of lsm_episode_stats (column sums + env count) when the caller already has them on the device."""
    import torch.distributed as dist
    s = sums if sums is not None else torch.cat([ep_info.double().sum(dim=0),
                                                 torch.tensor([float(ep_info.shape[0])], dtype=torch.float64, device=ep_info.device)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    vals = (s[:-1] / s[-1]).cpu().numpy()
    return {k: float(vals[j]) for j, k in enumerate(LY.EP_INFO_KEYS)}
