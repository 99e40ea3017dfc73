"""GPU tests of the grid-only sweep (`rwrt_ray_sweep_grid`) and of the device-resident ray records
(`rwrt_ray_run_rk45_records`, `rwrt_records_expand`, `rwrt_grid_accumulate_records`).

The gridding definition is this project's own (oracle/gridding.py; the reference ships no gridding code: parity
unpinned against the authors' tool -- what the tool would read is wr.py:947-959), so "parity" here is:
  * the grids of a sweep, which never materialises a trajectory, equal -- bit for bit -- the NumPy definition applied
    to the trajectories `ray_run` produces for the same sources (those are pinned against the reference elsewhere);
  * the fused form (the integrator accumulates) equals the records form (a second kernel grids each window's records);
  * records expanded on demand equal the planes of `ray_run`, bit for bit, for any row range;
  * any sharding of the sources over (virtual) ranks sums to the same integer grids.
"""
import ctypes as C

import numpy as np
import pytest

import parity
import rwrt_b200  # noqa: F401
from rwrt_b200 import dist as rdist
from rwrt_b200 import gridding, lib, synthetic
from rwrt_b200.wr import WR
from oracle import gridding as ogrid

pytestmark = pytest.mark.gpu


def make(u, v, zwn, nsource, days, freq=0.0):
    nlon, nlat = u.shape
    w = WR(len(zwn), nsource, 7200.0, days * 86400.0, freq, nx=nlon, ny=nlat)
    w.bs.lon, w.bs.lat = synthetic.grid_coords(nlon, nlat)
    w.bs.u, w.bs.v = u, v
    w.bs.ready(xcyclic=True)
    w.set_zwn(zwn)
    return w


def _case(ns=300, seed=7, days=12.0, zwn=(1.0, 2.0, 3.0, 5.0), grid=(360, 181)):
    u, v = synthetic.clim(*grid)
    sl, sa = synthetic.random_sources(ns, seed)
    return u, v, sl, sa, np.array(zwn), days


@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
@pytest.mark.parametrize("cell", [(2.5, 2.5), (1.0, 1.0), (7.0, 3.0)])
def test_sweep_grids_equal_oracle_gridding_of_trajectories(cell, strict):
    u, v, sl, sa, zwn, days = _case()
    w = make(u, v, zwn, len(sl), days)
    w.set_source_array(sl, sa)
    w.ray_run(mode="numpy", inte_method="rk45", strict=strict)
    out7 = np.stack([getattr(w, n) for n in parity.NAMES])
    ref = np.stack(ogrid.accumulate(out7, *cell))
    acc_ref, status_ref, total_ref = w.accepted.copy(), w.status.copy(), w.total_accepted
    assert ref[0].sum() > 1e5
    # fused: the integrator adds each output sample to the grids itself -- no rows are stored anywhere
    w2 = make(u, v, zwn, len(sl), days)
    w2.set_source_array(sl, sa)
    g = gridding.sweep(w2, *cell, strict=strict)
    assert g.shape == ref.shape and g.dtype.is_floating_point is False
    assert np.array_equal(g.cpu().numpy(), ref)
    assert w2.d_out is None                                       # nothing of shape (7, nt, N) was ever allocated
    assert np.array_equal(w2.accepted, acc_ref) and np.array_equal(w2.status, status_ref) and w2.total_accepted == total_ref
    # records form: time windows of 64-byte ray records, gridded by grid_records_kernel
    g2 = gridding.sweep(w2, *cell, strict=strict, from_records=True)
    assert np.array_equal(g2.cpu().numpy(), ref)
    # accumulate-into semantics: a second sweep into the same buffer doubles every cell
    g3 = gridding.sweep(w2, *cell, out=g, strict=strict)
    assert g3 is g and np.array_equal(g.cpu().numpy(), 2 * ref)


def test_sweep_from_records_in_time_windows(monkeypatch):
    u, v, sl, sa, zwn, days = _case(ns=160, seed=21, days=9.0)
    w = make(u, v, zwn, len(sl), days)
    w.set_source_array(sl, sa)
    ref = gridding.sweep(w, 2.0, 2.0).cpu().numpy()
    for nwin in ("3", "7", "1000"):
        monkeypatch.setenv("RWRT_WINDOWS", nwin)
        got = gridding.sweep(w, 2.0, 2.0, from_records=True).cpu().numpy()
        assert np.array_equal(got, ref), nwin
    monkeypatch.delenv("RWRT_WINDOWS")


def test_sweep_quarter_degree_and_frequency():
    """BASELINE configs[4] / configs[2] in miniature: 0.25-degree background (100 MB point table), non-zero frequency."""
    u, v = synthetic.clim(1440, 721)
    sl, sa = synthetic.random_sources(90, 5, max_abs_lat_deg=48.0)
    zwn = np.arange(1.0, 7.0)
    freq = 2 * np.pi / (30.0 * 86400.0)
    w = make(u, v, zwn, len(sl), 8.0, freq)
    w.set_source_array(sl, sa)
    w.ray_run(mode="numpy", inte_method="rk45")
    ref = np.stack(ogrid.accumulate(np.stack([getattr(w, n) for n in parity.NAMES]), 0.25, 0.25))
    g = gridding.sweep(w, 0.25, 0.25)
    assert g.shape == (3, 4320, 720)
    assert np.array_equal(g.cpu().numpy(), ref) and ref[0].sum() > 1e4
    assert gridding.checksum(g) == gridding.checksum(__import__("torch").from_numpy(ref).cuda())


def test_sweep_virtual_ranks_sum_to_single_batch():
    """The north-star multi-GPU path on one GPU: R virtual ranks sweep their share (contiguous blocks and strided
    shards) one after another; the summed integer grids equal the single-batch grids bit for bit."""
    u, v, sl, sa, zwn, days = _case(ns=203, seed=77, days=8.0, zwn=(1.0, 2.0, 5.0))
    full = rdist.sweep_sharded(lambda n: make(u, v, zwn, n, days), sl, sa, (2.0, 2.0), rank=0, world=1)
    ref = full["grids"].cpu().numpy()
    cs = gridding.checksum(full["grids"])
    for world, shard in ((2, "block"), (3, "stride"), (8, "stride"), (8, "block")):
        tot = np.zeros_like(ref)
        steps = 0
        seen = []
        for rank in range(world):
            r = rdist.sweep_sharded(lambda n: make(u, v, zwn, n, days), sl, sa, (2.0, 2.0), rank=rank, world=world, shard=shard)
            tot += r["grids"].cpu().numpy()
            steps += r["wr"].total_accepted
            seen.append(r["indices"])
        assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(len(sl)))
        assert np.array_equal(tot, ref), (world, shard)
        assert steps == full["wr"].total_accepted
        assert gridding.checksum(__import__("torch").from_numpy(tot).cuda()) == cs


def _records_run(u, v, sl, sa, zwn, days, flags=lib.MODE_FAST):
    import torch
    nlon, nlat = u.shape
    lon, lat = synthetic.grid_coords(nlon, nlat)
    L = lib.load()
    h = lib.vp()
    lib.check(L.rwrt_create(0, nlon, nlat, 1, 1, C.byref(h)))
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(np.ascontiguousarray(x)).to(dev) for k, x in
         dict(u=u, v=v, lon=lon, lat=lat, cl=np.cos(lat), sl=np.sin(lat), slon=np.radians(sl), slat=np.radians(sa), zwn=zwn).items()}
    sp = lib.stream_ptr()
    lib.check(L.rwrt_bs_ready(h, lib.ptr(d["u"]), lib.ptr(d["v"]), lib.ptr(d["lon"]), lib.ptr(d["lat"]), lib.ptr(d["cl"]),
                              lib.ptr(d["sl"]), None, None, sp))
    ns, nz = len(sl), len(zwn)
    N = 3 * ns * nz
    nt = int(days * 86400.0 / 7200.0) + 1
    row0 = torch.empty((7, N), dtype=torch.float64, device=dev)
    acc = torch.full((N,), -1, dtype=torch.int64, device=dev)
    status = torch.full((N,), -1, dtype=torch.int8, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    lib.check(L.rwrt_ray_init(h, ns, lib.ptr(d["slon"]), lib.ptr(d["slat"]), None, nz, lib.ptr(d["zwn"]), 0.0, 1, lib.ptr(row0), sp))
    lib.check(L.rwrt_ray_run_rk45_records(h, ns, None, nz, nt, 7200.0, days * 86400.0, 1e-6, 1e-6, 1e-3, 0.1, flags,
                                          lib.ptr(row0), lib.ptr(acc), None, lib.ptr(status), lib.ptr(total), sp),
              "rwrt_ray_run_rk45_records")
    torch.cuda.synchronize()
    return L, h, dev, row0, acc, status, int(total.item()), nt, N


@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
def test_records_native_output_expands_to_the_reference_planes(strict):
    import torch
    u, v, sl, sa, zwn, days = _case(ns=240, seed=31, days=10.0)
    flags = lib.MODE_STRICT if strict else lib.MODE_FAST
    w = make(u, v, zwn, len(sl), days)
    w.set_source_array(sl, sa)
    w.ray_run(mode="numpy", inte_method="rk45", strict=strict)
    ref = np.stack([getattr(w, n) for n in parity.NAMES]).reshape(7, w.nt, -1)
    L, h, dev, row0, acc, status, total, nt, N = _records_run(u, v, sl, sa, zwn, days, flags)
    try:
        assert total == w.total_accepted
        assert np.array_equal(acc.cpu().numpy().reshape(w.accepted.shape), w.accepted)
        assert np.array_equal(status.cpu().numpy().reshape(w.status.shape), w.status)
        assert np.array_equal(row0.cpu().numpy(), ref[:, 0], equal_nan=True)
        # the records themselves
        nray, rows = C.c_int64(0), C.c_int(0)
        p = [lib.vp() for _ in range(4)]
        lib.check(L.rwrt_records_info(h, C.byref(nray), C.byref(rows), *[C.byref(x) for x in p]))
        live = ~np.isnan(ref[3, 0])
        assert nray.value == int(live.sum()) and rows.value >= nt - 1
        # on-demand expansion: all rows at once, and an inner range
        for r0, r1 in ((1, nt), (17, 60), (nt - 1, nt)):
            dst = torch.full((7, r1 - r0, N), -3.0, dtype=torch.float64, device=dev)
            lib.check(L.rwrt_records_expand(h, r0, r1, lib.ptr(dst), lib.stream_ptr()))
            torch.cuda.synchronize()
            assert np.array_equal(dst.cpu().numpy(), ref[:, r0:r1], equal_nan=True), (r0, r1)
        # gridding straight from the records == NumPy definition on the planes
        for cell in ((2.0, 2.0), (1.0, 0.5)):
            g = gridding.new_grids(*cell, dev)
            ncx, ncy = gridding.grid_shape(*cell)
            lib.check(L.rwrt_grid_accumulate_records(h, cell[0], cell[1], ncx, ncy, lib.ptr(g[0]), lib.ptr(g[1]), lib.ptr(g[2]),
                                                     lib.stream_ptr()))
            torch.cuda.synchronize()
            assert np.array_equal(g.cpu().numpy(), np.stack(ogrid.accumulate(ref.reshape(7, nt, 3, len(sl), len(zwn)), *cell)))
        # error paths: rows outside the run, records gone after trim
        assert L.rwrt_records_expand(h, 0, 5, lib.ptr(row0), None) != 0
        lib.check(L.rwrt_trim(h))
        assert L.rwrt_records_expand(h, 1, 2, lib.ptr(row0), None) != 0 and b"no ray records" in L.rwrt_last_error()
    finally:
        L.rwrt_destroy(h)


def test_sweep_argument_errors_and_empty():
    import torch
    u, v = synthetic.jet_z()
    w = make(u, v, np.array([1.0, 2.0]), 0, 2.0)
    w.set_source_array([], [])
    g = gridding.sweep(w, 5.0, 5.0)
    assert int(g.sum().item()) == 0 and w.total_accepted == 0
    L = lib.load()
    z = torch.zeros(4, dtype=torch.int64, device="cuda")
    rc = L.rwrt_ray_sweep_grid(w.bs._h, 0, None, None, None, 2, None, 0.0, 5, 7200.0, 28800.0, 1e-6, 1e-6, 1e-3, 0.1, 0,
                               -1.0, 5.0, 2, 2, lib.ptr(z), lib.ptr(z), lib.ptr(z), None, None, None, None)
    assert rc == -1 and b"grid" in L.rwrt_last_error()


def test_results_on_device_are_fetched_not_nan(recwarn):
    """after ray_run(to_host=False) the result arrays must come from the GPU, never the NaN pre-fill."""
    u, v, sl, sa, zwn, days = _case(ns=40, seed=3, days=3.0)
    w = make(u, v, zwn, len(sl), days)
    w.set_source_array(sl, sa)
    w.ray_run(mode="numpy", inte_method="rk45", to_host=False)
    assert np.isfinite(w.rlon[-1]).any()                          # fetched lazily
    w2 = make(u, v, zwn, len(sl), days)
    w2.set_source_array(sl, sa)
    w2.ray_run(mode="numpy", inte_method="rk45")
    for n in parity.NAMES:
        assert np.array_equal(getattr(w, n), getattr(w2, n), equal_nan=True)
    w3 = make(u, v, zwn, len(sl), days)
    w3.set_source_array(sl, sa)
    w3.ray_run(mode="numpy", inte_method="rk45", to_host=False)
    with pytest.warns(RuntimeWarning):
        w3.clean()


def test_wn_grid_maps_equal_ray_init_of_the_oracle():
    """`WN.cal_wave` (wn.py:58-87 as intended; parity unpinned: the reference's method never runs): maps of l, root
    count, ug, vg at every grid point == the oracle's t = 0 state for sources at the grid points."""
    from rwrt_b200.wn import WN
    from oracle import oracle
    u, v = synthetic.clim(144, 73)
    lon, lat = synthetic.grid_coords(144, 73)
    zwn = np.array([1.0, 3.0, 6.0])
    wn = WN(144, 73, 3, freq=2 * np.pi / (40 * 86400.0))
    wn.bs.lon, wn.bs.lat, wn.bs.u, wn.bs.v = lon, lat, u, v
    wn.bs.ready(xcyclic=True)
    wn.set_zwn(zwn)
    wn.cal_wave()
    F = oracle.bs_ready(u, v, lat, True, np.cos(lat), np.sin(lat))
    ref = oracle.ray_init(F, lon, lat, np.repeat(lon, 73), np.tile(lat, 144), zwn, wn.freq, 1)[:, 0]     # (7, 3, ns, nz)
    l_ref = np.moveaxis(ref[3], 0, -1).reshape(144, 73, 3, 3)
    assert wn.mwn.shape == (144, 73, 3, 3) and wn.rootnum.shape == (144, 73, 3)
    assert np.array_equal(np.isnan(wn.mwn), np.isnan(l_ref))
    np.testing.assert_allclose(wn.mwn, l_ref, rtol=1e-12, equal_nan=True)
    assert np.array_equal(wn.rootnum, (~np.isnan(l_ref)).sum(-1))
    ug_ref = np.where(np.isnan(l_ref), 0.0, np.moveaxis(ref[5], 0, -1).reshape(144, 73, 3, 3))
    np.testing.assert_allclose(wn.ug, ug_ref, rtol=1e-10, atol=1e-12)
    assert wn.rootnum.max() == 3 and wn.rootnum.min() == 0 and (wn.rootnum[:, 0] == 0).all()     # nothing propagates at the pole rows
    with pytest.raises(ValueError):
        wn.set_zwn([1.0])
