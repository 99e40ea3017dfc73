"""Ray-density / wave-ray-flux gridding of trajectories on the GPU (`rwrt_grid_accumulate`).

The reference ships no gridding code (its manual describes tools that are not in the repository),
so this is an addition with this project's own definition (see `csrc/rwrt_grid.cu`): counts of live
ray samples and fixed-point sums of the group velocity per cell of a (dlon, dlat) grid spanning three
longitude circles [-360, 720) deg.  Integer accumulators make the result independent of the order
of accumulation, hence bit-identical for any sharding of the rays over GPUs.
"""
from __future__ import annotations

import math

from . import lib as _lib

FLUX_SCALE = 1048576.0   # fx, fy are sums of rint(ug * 2^20)


def grid_shape(dlon_deg, dlat_deg):
    return int(math.ceil(1080.0 / dlon_deg)), int(math.ceil(180.0 / dlat_deg))


def accumulate(wr, dlon_deg, dlat_deg, out=None):
    """Accumulate the trajectories held on the GPU by `wr` (after ray_run).  Returns (count, fx, fy),
    torch int64 CUDA tensors of shape (ncx, ncy); pass `out=(count, fx, fy)` to add into existing grids."""
    import torch
    if wr.d_out is None:
        raise RuntimeError("ray_run first")
    dev = wr.d_out.device
    ncx, ncy = grid_shape(dlon_deg, dlat_deg)
    if out is None:
        out = tuple(torch.zeros((ncx, ncy), dtype=torch.int64, device=dev) for _ in range(3))
    count, fx, fy = out
    nslot = 3 * wr.nsource * wr.nzwn
    with torch.cuda.device(dev):
        _lib.check(_lib.load().rwrt_grid_accumulate(wr.bs._h, nslot, wr.nt, _lib.ptr(wr.d_out), float(dlon_deg),
                                                    float(dlat_deg), ncx, ncy, _lib.ptr(count), _lib.ptr(fx),
                                                    _lib.ptr(fy), _lib.stream_ptr()), "rwrt_grid_accumulate")
    return count, fx, fy


def flux(count, fx, fy):
    """Mean group velocity per cell (m/s) from the integer accumulators (NaN where no ray passed)."""
    import torch
    c = count.to(torch.float64)
    c = torch.where(c > 0, c, torch.full_like(c, float("nan")))
    return fx.to(torch.float64) / FLUX_SCALE / c, fy.to(torch.float64) / FLUX_SCALE / c
