"""Batch sharding of the sampling job across the GPUs of one box: one process per GPU, no
collective inside the reverse loop (the denoiser's only cross-point op is the max over the points
of ONE cloud, reference networks.py:807).  The global sample index keys the Philox noise so a job
produces the same clouds for 1/2/4/8 ranks."""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first global sample index, number of samples) owned by `rank`; contiguous, balanced to 1."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def dist_info() -> Tuple[int, int]:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


@torch.no_grad()
def sample_sharded(model, total_samples: int, num_points: int, num_steps: int, kind: str = "ddpm", *, seed: int = 0,
                   x_T_seed: int = 5, max_batch: int = 512):
    """This rank's shard of a `total_samples` generation job (BASELINE config 3): returns
    (clouds [count, N, 3] on the model's device, first global index).  x_T for global sample g is
    drawn from a CPU generator seeded with (x_T_seed, g) so it does not depend on the sharding."""
    rank, world = dist_info()
    start, count = shard_range(total_samples, rank, world)
    outs = []
    for off in range(0, count, max_batch):
        b = min(max_batch, count - off)
        xT = torch.stack([torch.randn(num_points, 3, generator=torch.Generator().manual_seed(x_T_seed * 1_000_003 + start + off + i))
                          for i in range(b)])
        if kind == "ddpm":
            outs.append(model.sample2(b, num_points, num_steps, x_T=xT, seed=seed, sample_offset=start + off))
        else:
            outs.append(model.sample(b, num_points, num_steps, x_T=xT, sample_offset=start + off))
    out = torch.cat(outs) if outs else torch.empty(0, num_points, 3, device=model.device)
    return out, start
