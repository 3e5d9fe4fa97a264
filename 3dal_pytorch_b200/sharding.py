"""Multi-GPU host logic: tracks are independent, so the forward shards by contiguous blocks of tracks
with no data-path collective.  The only exchanges are

  * ``gather_boxes``   one all_gather of the refined (tracks, 7) boxes (NCCL on GPUs, gloo in the CPU tests);
  * ``global_choice_tables``  parity mode only: the reference consumes ONE numpy RNG stream over the whole
    batch (tools/static_model.py:36-45), so the per-rank foreground counts are all-gathered, every rank
    replays the identical stream and keeps the rows of its own tracks.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import engine


def shard_range(total, rank, world):
    """Contiguous block [lo, hi) of `total` tracks for `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_boxes(local_boxes, total, group=None):
    """local_boxes (n_local, 7) -> (total, 7) on every rank, in track order (uneven shards are padded)."""
    world = dist.get_world_size(group)
    if world == 1:
        return local_boxes
    width = local_boxes.shape[1]
    n_max = -(-int(total) // world)
    padded = torch.zeros((n_max, width), dtype=local_boxes.dtype, device=local_boxes.device)
    padded[: local_boxes.shape[0]] = local_boxes
    out = torch.empty((world * n_max, width), dtype=local_boxes.dtype, device=local_boxes.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        parts.append(out[r * n_max: r * n_max + (hi - lo)])
    return torch.cat(parts, 0)


def global_choice_tables(local_counts, total, n_pts, group=None):
    """local_counts (n_local,) int tensor of foreground counts -> this rank's (n_local, n_pts) int32 rows of
    the table the single-process reference run would draw (the numpy RNG state must be identical on all
    ranks on entry, e.g. np.random.seed(s) everywhere)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_max = -(-int(total) // world)
    padded = torch.zeros((n_max,), dtype=torch.int64, device=local_counts.device)
    padded[: local_counts.shape[0]] = local_counts.to(torch.int64)
    allc = torch.empty((world * n_max,), dtype=torch.int64, device=local_counts.device)
    dist.all_gather_into_tensor(allc, padded, group=group)
    allc = allc.cpu().numpy()
    counts = np.concatenate([allc[r * n_max: r * n_max + (shard_range(total, r, world)[1] - shard_range(total, r, world)[0])]
                             for r in range(world)])
    table = engine.choice_table_numpy_legacy(counts, n_pts)
    lo, hi = shard_range(total, rank, world)
    return table[lo:hi]
