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


def new_grids(dlon_deg, dlat_deg, device):
    """One zeroed (3, ncx, ncy) int64 tensor = count, fx, fy: a single contiguous buffer, so that the ranks of a
    sharded sweep merge with ONE all-reduce (`dist.allreduce_grids(g)`)."""
    import torch
    ncx, ncy = grid_shape(dlon_deg, dlat_deg)
    return torch.zeros((3, ncx, ncy), dtype=torch.int64, device=device)


def sweep(wr, dlon_deg, dlat_deg, out=None, strict=False, from_records=False, want_accepted=True):
    """Grid-only sweep (`rwrt_ray_sweep_grid`): t = 0 state + RK45 + gridding of `wr`'s sources in one pass with NO
    trajectory arrays -- neither the (7, nt, N) planes the reference allocates up front (wr.py:157-167) nor ray
    records: the integrator adds every output sample to the grids as it produces it (from_records=True: time
    windows of ray records gridded by a second kernel -- same bits).  Returns the (3, ncx, ncy) int64 grids
    (count, fx, fy), added into `out` when given.  Sets wr.total_accepted and, if want_accepted, wr.accepted / wr.status."""
    import numpy as np
    import torch
    wr._check_ready()
    dev = torch.device("cuda", wr.bs.device)
    ncx, ncy = grid_shape(dlon_deg, dlat_deg)
    if out is None:
        out = new_grids(dlon_deg, dlat_deg, dev)
    N = 3 * wr.nsource * wr.nzwn
    d_lon = torch.from_numpy(np.ascontiguousarray(wr.source_lon, dtype=np.float64)).to(dev)
    d_lat = torch.from_numpy(np.ascontiguousarray(wr.source_lat, dtype=np.float64)).to(dev)
    d_zwn = torch.from_numpy(np.ascontiguousarray(wr.zwn, dtype=np.float64)).to(dev)
    d_acc = torch.empty(N, dtype=torch.int64, device=dev) if want_accepted else None
    d_status = torch.empty(N, dtype=torch.int8, device=dev) if want_accepted else None
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    flags = (_lib.MODE_STRICT if strict else _lib.MODE_FAST) | (_lib.GRID_FROM_RECORDS if from_records else 0)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().rwrt_ray_sweep_grid(
            wr.bs._h, wr.nsource, _lib.ptr(d_lon), _lib.ptr(d_lat), None, wr.nzwn, _lib.ptr(d_zwn), float(wr.freq[0]),
            wr.nt, float(wr.tstep[0]), float(wr.ttotal), float(wr.rtol), float(wr.atol), float(wr.MinStepFactor),
            float(wr.cut_off_factor), flags, float(dlon_deg), float(dlat_deg), ncx, ncy, _lib.ptr(out[0]), _lib.ptr(out[1]),
            _lib.ptr(out[2]), _lib.ptr(d_acc), _lib.ptr(d_status), _lib.ptr(d_total), _lib.stream_ptr()),
            "rwrt_ray_sweep_grid")
        wr.total_accepted = int(d_total.item())
    if want_accepted:
        shp = (3, wr.nsource, wr.nzwn)
        wr.accepted = d_acc.cpu().numpy().reshape(shp)
        wr.status = d_status.cpu().numpy().reshape(shp)
    return out


def checksum(grids):
    """Order-independent 64-bit fingerprint of integer grids (sum of value * (2 * index + 1), wrapping): equal across
    any sharding of the rays iff the reduced grids are equal (bench.py prints it per GPU count)."""
    import torch
    g = grids.reshape(-1)
    w = torch.arange(g.numel(), dtype=torch.int64, device=g.device) * 2 + 1
    return int((g * w).sum().item()) & 0xFFFFFFFFFFFFFFFF


def flux(count, fx, fy):
    """Mean group velocity per cell (m/s) from the integer accumulators (NaN where no ray passed)."""
    import torch
    c = count.to(torch.float64)
    c = torch.where(c > 0, c, torch.full_like(c, float("nan")))
    return fx.to(torch.float64) / FLUX_SCALE / c, fy.to(torch.float64) / FLUX_SCALE / c
