"""GPU tests: gridded ray counts / flux sums bit-exact against the NumPy definition on the same
trajectories; virtual-rank sharding (R ranks emulated sequentially on one GPU) reproduces the
single-batch result bit for bit, and the summed integer grids equal the single-batch grids."""
import numpy as np
import pytest

import parity
import rwrt_b200  # noqa: F401
from rwrt_b200 import dist as rdist
from rwrt_b200 import gridding, synthetic
from rwrt_b200.wr import WR
from oracle import gridding as ogrid

pytestmark = pytest.mark.gpu


def make(u, v, zwn, nsource, days):
    nlon, nlat = u.shape
    w = WR(len(zwn), nsource, 7200.0, days * 86400.0, 0.0, nx=nlon, ny=nlat)
    w.bs.lon, w.bs.lat = synthetic.grid_coords(nlon, nlat)
    w.bs.u, w.bs.v = u, v
    w.bs.ready(xcyclic=True)
    w.set_zwn(zwn)
    return w


@pytest.mark.parametrize("cell", [(2.5, 2.5), (1.0, 1.0), (7.0, 3.0)])
def test_grid_counts_bit_exact(cell):
    u, v = synthetic.clim(360, 181)
    sl, sa = synthetic.random_sources(400, 61)
    zwn = np.array([1.0, 2.0, 3.0, 4.0])
    w = make(u, v, zwn, 400, 12.0)
    w.set_source_array(sl, sa)
    w.ray_run(mode="numpy", inte_method="rk45")
    count, fx, fy = gridding.accumulate(w, *cell)
    out7 = np.stack([getattr(w, n) for n in parity.NAMES])
    rc, rx, ry = ogrid.accumulate(out7, *cell)
    assert count.shape == rc.shape
    assert np.array_equal(count.cpu().numpy(), rc)
    assert np.array_equal(fx.cpu().numpy(), rx) and np.array_equal(fy.cpu().numpy(), ry)
    assert rc.sum() > 1e5
    # rays do leave [0, 360): the three-circle domain is really used
    assert rc[: rc.shape[0] // 3].sum() + rc[2 * rc.shape[0] // 3:].sum() > 0
    mx, my = gridding.flux(count, fx, fy)
    assert np.isfinite(mx.cpu().numpy()[rc > 0]).all()


def test_virtual_ranks_equal_single_batch():
    """Source sharding is exact: 1 / 2 / 3 / 8 virtual ranks, run one after another on this GPU."""
    u, v = synthetic.clim(360, 181)
    sl, sa = synthetic.random_sources(203, 77)
    zwn = np.array([1.0, 2.0, 5.0])
    days = 8.0
    full = rdist.run_sharded(lambda n: make(u, v, zwn, n, days), sl, sa, grid=(2.0, 2.0), rank=0, world=1, to_host=True)
    ref = {n: getattr(full["wr"], n).copy() for n in parity.NAMES}
    ref_acc = full["wr"].accepted.copy()
    ref_grids = [g.cpu().numpy() for g in full["grids"]]
    for world in (2, 3, 8):
        grids = [np.zeros_like(g) for g in ref_grids]
        total = 0
        for rank in range(world):
            r = rdist.run_sharded(lambda n: make(u, v, zwn, n, days), sl, sa, grid=(2.0, 2.0), rank=rank, world=world,
                                  to_host=True)
            s0, s1 = r["bounds"]
            for n in parity.NAMES:
                assert np.array_equal(getattr(r["wr"], n), ref[n][:, :, s0:s1, :], equal_nan=True), (world, rank, n)
            assert np.array_equal(r["wr"].accepted, ref_acc[:, s0:s1, :])
            for g, part in zip(grids, r["grids"]):
                g += part.cpu().numpy()
            total += r["wr"].total_accepted
        for g, rg in zip(grids, ref_grids):
            assert np.array_equal(g, rg)
        assert total == full["wr"].total_accepted
