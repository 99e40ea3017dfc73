"""Multi-GPU: rays are independent, so the SOURCE axis is sharded in contiguous blocks over ranks
(one process per GPU, `torch.distributed`); no collective touches the ray path.  The only exchange
is one all-reduce (sum) of the integer gridded accumulators (NCCL over NVLink on GPUs, gloo on CPU).

A ray's result does not depend on which other rays share its batch (tests/test_gpu_parity.py::
test_batch_composition_independence; SURVEY.md probe9), so the concatenation of the ranks' slices
equals the single-GPU result bit for bit.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(nsource: int, rank: int, world: int):
    """Contiguous block [s0, s1) of the source axis for `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    base, rem = divmod(int(nsource), int(world))
    s0 = rank * base + min(rank, rem)
    s1 = s0 + base + (1 if rank < rem else 0)
    return s0, s1


def shard_indices(nsource: int, rank: int, world: int, mode: str = "block"):
    """Source indices of `rank`.  "block": the contiguous block of `shard_bounds` (a rank's trajectories are a slice
    of the source axis).  "stride": rank, rank + world, ... -- for grid-only sweeps, whose result does not depend on
    which rank integrates which source; sources are ordered by latitude and the work per source varies with it, so
    strided shards are balanced where latitude bands are not."""
    if mode == "block":
        s0, s1 = shard_bounds(nsource, rank, world)
        return np.arange(s0, s1)
    if mode == "stride":
        if world < 1 or not (0 <= rank < world):
            raise ValueError("need 0 <= rank < world")
        return np.arange(rank, int(nsource), world)
    raise ValueError("mode must be 'block' or 'stride'")


def allreduce_grids(*grids):
    """In-place sum of integer accumulator tensors over all ranks (no-op without a process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for g in grids:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
    return grids


def sweep_sharded(make_wr, src_lon_deg, src_lat_deg, grid, rank=None, world=None, shard="stride", strict=False,
                  from_records=False):
    """Grid-only sweep of this rank's share of the sources (`gridding.sweep`: no trajectory arrays) followed by the
    path's one collective: a single all-reduce (sum) of the (3, ncx, ncy) int64 accumulators.  Returns
    dict(wr=..., indices=..., grids=all-reduced (3, ncx, ncy) tensor)."""
    import torch.distributed as dist
    from . import gridding
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    idx = shard_indices(len(src_lon_deg), rank, world, shard)
    w = make_wr(len(idx))
    w.set_source_array(np.asarray(src_lon_deg)[idx], np.asarray(src_lat_deg)[idx])
    grids = gridding.sweep(w, grid[0], grid[1], strict=strict, from_records=from_records)
    allreduce_grids(grids)
    return {"wr": w, "indices": idx, "grids": grids}


def run_sharded(make_wr, src_lon_deg, src_lat_deg, grid=None, rank=None, world=None, to_host=False):
    """Run this rank's block of sources.

    make_wr(nsource_local) -> a `WR` with its background state set and `bs.ready()` done.
    Returns dict(wr=..., bounds=(s0, s1), grids=(count, fx, fy) or None) with the grids all-reduced.
    """
    import torch.distributed as dist
    from . import gridding
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    s0, s1 = shard_bounds(len(src_lon_deg), rank, world)
    w = make_wr(s1 - s0)
    w.set_source_array(np.asarray(src_lon_deg)[s0:s1], np.asarray(src_lat_deg)[s0:s1])
    w.ray_run(mode="numpy", root_method="numpy", inte_method="rk45", to_host=to_host)
    grids = None
    if grid is not None:
        grids = gridding.accumulate(w, grid[0], grid[1])
        allreduce_grids(*grids)
    return {"wr": w, "bounds": (s0, s1), "grids": grids}
