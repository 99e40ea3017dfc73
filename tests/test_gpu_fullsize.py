"""Full-size GPU checks (BASELINE.json configs[1]: 63 720 sources x k=1..6 x 3 roots = 1 146 960 ray slots,
30 days, 23.2 GB of trajectories on the device) through the C ABI with device buffers.  The oracle cannot
run this size in seconds, so the run is checked through size-independent properties of the reference's
algorithm, plus bit-exact agreement of a random sample of sources with small runs of the same engine
(which the other tests pin against the oracle and the reference fixtures):

  * a ray's bits do not depend on its batch (SURVEY.md probe9)              -> sample == small runs, bit for bit
  * NaN is absorbing (wr.py:838-855: a terminated ray stays NaN)            -> NaN mask monotone in time
  * NaN-root slots never move                      -> lon/lat/k rows constant, rest NaN
  * |lat| < pi/2 wherever a ray is alive (wr.py:838-843)
  * accepted-step bookkeeping: sum(accepted) == total, accepted == 0 on NaN-root slots, >= alive rows - 1 on live
  * idempotence and time-window invariance (checksum of the raw bits of all seven planes)
"""
import ctypes as C

import numpy as np
import pytest

import rwrt_b200  # noqa: F401
from rwrt_b200 import lib, synthetic

pytestmark = pytest.mark.gpu

TSTEP, DAYS = 7200.0, 30.0


def _setup(nlon, nlat, slon_deg, slat_deg, zwn):
    import torch
    L = lib.load()
    lib.require_gpu()
    dev = torch.device("cuda", 0)
    u, v = synthetic.clim(nlon, nlat)
    lon, lat = synthetic.grid_coords(nlon, nlat)
    h = lib.vp()
    lib.check(L.rwrt_create(0, nlon, nlat, 1, 1, C.byref(h)), "rwrt_create")
    d = {k: torch.from_numpy(np.ascontiguousarray(x)).to(dev) for k, x in
         dict(u=u, v=v, lon=lon, lat=lat, cl=np.cos(lat), sl=np.sin(lat), slon=np.radians(slon_deg),
              slat=np.radians(slat_deg), zwn=zwn).items()}
    sp = lib.stream_ptr()
    lib.check(L.rwrt_bs_ready(h, lib.ptr(d["u"]), lib.ptr(d["v"]), lib.ptr(d["lon"]), lib.ptr(d["lat"]),
                              lib.ptr(d["cl"]), lib.ptr(d["sl"]), None, None, sp), "rwrt_bs_ready")
    return L, h, d, dev


def _run(L, h, d, dev, nsource, nzwn, flags=lib.MODE_FAST):
    import torch
    nt = int(DAYS * 86400.0 / TSTEP) + 1
    N = 3 * nsource * nzwn
    out = torch.empty((7, nt, N), dtype=torch.float64, device=dev)
    acc = torch.zeros(N, dtype=torch.int64, device=dev)
    status = torch.zeros(N, dtype=torch.int8, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    sp = lib.stream_ptr()
    lib.check(L.rwrt_ray_init(h, nsource, lib.ptr(d["slon"]), lib.ptr(d["slat"]), None, nzwn, lib.ptr(d["zwn"]), 0.0, nt,
                              lib.ptr(out), sp), "rwrt_ray_init")
    lib.check(L.rwrt_ray_run_rk45(h, nsource, None, nzwn, nt, TSTEP, DAYS * 86400.0, 1e-6, 1e-6, 1e-3, 0.1, flags,
                                  lib.ptr(out), lib.ptr(acc), None, lib.ptr(status), lib.ptr(total), sp), "rwrt_ray_run_rk45")
    torch.cuda.synchronize()
    return out, acc, status, int(total.item()), nt, N


def _checksum(out):
    # sum of the raw 64-bit patterns (wraps modulo 2^64): equal bits <=> equal checksum, NaNs included
    return int(out.view(-1).view(dtype=__import__("torch").int64).sum().item())


def test_c2_full_size_properties(monkeypatch):
    import torch
    nlon, nlat = 360, 181
    slon_deg, slat_deg = synthetic.all_gridpoint_sources(nlon, nlat, 88.0)
    zwn = np.arange(1.0, 7.0)
    ns, nz = len(slon_deg), len(zwn)
    assert 3 * ns * nz == 1146960
    L, h, d, dev = _setup(nlon, nlat, slon_deg, slat_deg, zwn)
    try:
        out, acc, status, total, nt, N = _run(L, h, d, dev, ns, nz)
        assert int(acc.sum().item()) == total and total > 2.0e8
        nan_l = torch.isnan(out[3])                                        # (nt, N): meridional wavenumber plane
        static = nan_l[0]                                                  # NaN root at t = 0
        # NaN is absorbing
        assert bool((nan_l[1:] | ~nan_l[:-1]).all()), "a ray came back from NaN"
        # NaN-root slots: lon, lat, k constant; l, A, ug, vg NaN; zero steps; status NO_ROOT
        for pl in range(3):
            col = out[pl][:, static]
            assert bool(((col == col[0:1]) | (torch.isnan(col) & torch.isnan(col[0:1]))).all()), pl
        for pl in (3, 5, 6):
            assert bool(torch.isnan(out[pl][1:, static]).all()), pl
        assert int(acc[static].abs().sum().item()) == 0
        assert bool((status[static] == lib.RAY_NO_ROOT).all()) and bool((status[~static] != lib.RAY_NO_ROOT).all())
        # alive rows: every plane finite, |lat| < pi/2
        alive = ~nan_l
        alive[:, static] = False
        for pl in range(5):
            assert bool(torch.isfinite(out[pl][alive]).all()), pl
        assert float(out[1][alive].abs().max().item()) < np.pi / 2
        # dead rows of live slots: all seven planes NaN
        dead = nan_l.clone()
        dead[:, static] = False
        for pl in range(7):
            assert bool(torch.isnan(out[pl][dead]).all()), pl
        # bookkeeping: a live ray that produced r rows took at least r - 1 accepted steps; survivors are ALIVE
        rows_alive = alive.sum(dim=0)
        live = ~static
        assert bool((acc[live] >= rows_alive[live] - 1).all())
        assert bool((status[live & (rows_alive == nt)] == lib.RAY_ALIVE).all())
        assert bool((status[live & (rows_alive < nt)] >= lib.RAY_POLE).all())
        del nan_l, alive, dead
        # sample of sources re-run alone in small batches: identical bits (fast build)
        rng = np.random.default_rng(7)
        idx = np.sort(rng.choice(ns, size=384, replace=False))
        L2, h2, d2, _ = _setup(nlon, nlat, slon_deg[idx], slat_deg[idx], zwn)
        try:
            o2, a2, s2, t2, _, N2 = _run(L2, h2, d2, dev, len(idx), nz)
        finally:
            L2.rwrt_destroy(h2)
        big = out.view(7, nt, 3, ns, nz)[:, :, :, torch.from_numpy(idx).to(dev), :]
        small = o2.view(7, nt, 3, len(idx), nz)
        same = (big == small) | (torch.isnan(big) & torch.isnan(small))
        assert bool(same.all()), "a ray's bits depend on its batch"
        assert torch.equal(acc.view(3, ns, nz)[:, torch.from_numpy(idx).to(dev), :], a2.view(3, len(idx), nz))
        del big, small, same, o2
        # idempotence and time-window invariance on the whole 23 GB
        ck, acc0 = _checksum(out), acc.clone()
        del out
        torch.cuda.empty_cache()
        out, acc, status, total2, _, _ = _run(L, h, d, dev, ns, nz)
        assert _checksum(out) == ck and total2 == total and torch.equal(acc, acc0)
        del out
        torch.cuda.empty_cache()
        monkeypatch.setenv("RWRT_WINDOWS", "5")
        out, acc, status, total3, _, _ = _run(L, h, d, dev, ns, nz)
        assert _checksum(out) == ck and total3 == total and torch.equal(acc, acc0)
    finally:
        L.rwrt_destroy(h)


def test_schedule_and_gather_independence(monkeypatch):
    """A quarter of C2 (287 k slots): the result may not depend on how rays are scheduled onto lanes (resident
    blocks per SM, time windows) nor on the output path (per-ray records + transpose, the default, or direct
    stores into the reference layout after a pre-fill) -- any stale or racing use of a warp's shared staging rows
    would show up here --
    and, in the strict build, the warp-cooperative gather must agree bit for bit with the per-lane gather kernel
    (two independent implementations of interpolation.py:103-135)."""
    nlon, nlat = 360, 181
    slon_deg, slat_deg = synthetic.all_gridpoint_sources(nlon, nlat, 88.0)
    slon_deg, slat_deg = slon_deg[1::4], slat_deg[1::4]
    zwn = np.arange(1.0, 7.0)
    ns, nz = len(slon_deg), len(zwn)
    L, h, d, dev = _setup(nlon, nlat, slon_deg, slat_deg, zwn)
    try:
        for flags in (lib.MODE_FAST, lib.MODE_STRICT):
            out, acc, status, total, nt, N = _run(L, h, d, dev, ns, nz, flags)
            ck, acc0 = _checksum(out), acc.clone()
            del out
            variants = [{"RWRT_BLOCKS_PER_SM": "1"}, {"RWRT_BLOCKS_PER_SM": "2"},
                        {"RWRT_BLOCKS_PER_SM": "2", "RWRT_WINDOWS": "4"}, {"RWRT_WINDOWS": "9"},
                        {"RWRT_RECORDS": "0"}, {"RWRT_RECORDS": "0", "RWRT_WINDOWS": "3"}]   # direct stores + pre-fill
            if flags == lib.MODE_STRICT:
                variants.append({"RWRT_PER_LANE_GATHER": "1"})
            for env in variants:
                for k, v in env.items():
                    monkeypatch.setenv(k, v)
                try:
                    out, acc, status, total2, _, _ = _run(L, h, d, dev, ns, nz, flags)
                finally:
                    for k in env:
                        monkeypatch.delenv(k)
                assert _checksum(out) == ck and total2 == total and bool((acc == acc0).all()), (flags, env)
                del out
    finally:
        L.rwrt_destroy(h)


def test_host_entry_quarter_size(monkeypatch):
    """rwrt_real2d_host at a quarter of C2 (287 k slots, 5.8 GB of result planes in pinned host memory): the default
    form (packed rows of the live rays cross the link window by window into pinned staging buffers, host threads write
    the planes, windows go straight into the caller's planes when the threads are behind -- every staging buffer and
    all device window buffers are reused several times), every forced routing, and the device-expanded form
    (RWRT_HOST_EXPAND=0) must return the bits of the device-resident run, repeatedly on one handle."""
    nlon, nlat = 360, 181
    slon_deg, slat_deg = synthetic.all_gridpoint_sources(nlon, nlat, 88.0)
    slon_deg, slat_deg = slon_deg[1::4], slat_deg[1::4]
    zwn = np.arange(1.0, 7.0)
    ns, nz = len(slon_deg), len(zwn)
    L, h, d, dev = _setup(nlon, nlat, slon_deg, slat_deg, zwn)
    try:
        out, acc, status, total, nt, N = _run(L, h, d, dev, ns, nz)
        ck = _checksum(out)
        acc0, status0 = acc.cpu().numpy(), status.cpu().numpy()
        del out
        __import__("torch").cuda.empty_cache()
        u, v = synthetic.clim(nlon, nlat)
        lon, lat = synthetic.grid_coords(nlon, nlat)
        cl, sn = np.cos(lat), np.sin(lat)
        slon, slat = np.radians(slon_deg), np.radians(slat_deg)
        import torch
        host_t = torch.empty((7, nt, N), dtype=torch.float64, pin_memory=True)   # pinned: windows may also go direct
        host = host_t.numpy()
        for env in ({}, {"RWRT_HOST_THREADS": "3", "RWRT_HOST_WINDOWS": "11"}, {"RWRT_HOST_EXPAND": "0"}, {"RWRT_HOST_ROUTE": "hx"},
                    {"RWRT_HOST_ROUTE": "alt"}, {"RWRT_HOST_ROUTE": "direct", "RWRT_HOST_WINDOWS": "5"},
                    {"RWRT_HOST_THREADS": "1"}, {}):     # one writer thread: the backlog rule sends windows direct
            for k, val in env.items():
                monkeypatch.setenv(k, val)
            try:
                host.fill(-7.0)
                acc_h = np.full(N, -1, dtype=np.int64)
                st_h = np.full(N, -1, dtype=np.int8)
                tot_h = np.zeros(1, dtype=np.int64)
                lib.check(L.rwrt_real2d_host(h, lib.ptr(u), lib.ptr(v), lib.ptr(lon), lib.ptr(lat), lib.ptr(cl), lib.ptr(sn), ns,
                                             lib.ptr(slon), lib.ptr(slat), None, nz, lib.ptr(zwn), 0.0, nt, TSTEP,
                                             DAYS * 86400.0, 1e-6, 1e-6, 1e-3, 0.1, lib.MODE_FAST, lib.ptr(host), lib.ptr(acc_h),
                                             lib.ptr(st_h), lib.ptr(tot_h), None), "rwrt_real2d_host")
            finally:
                for k in env:
                    monkeypatch.delenv(k)
            with np.errstate(over="ignore"):
                ck_h = int(host.reshape(-1).view(np.int64).sum(dtype=np.int64))
            assert ck_h == ck, env
            assert int(tot_h[0]) == total and np.array_equal(acc_h, acc0) and np.array_equal(st_h, status0), env
    finally:
        L.rwrt_destroy(h)


def _sweep(L, h, d, dev, slon, slat, zwn, nt, ttotal, cell, freq=0.0, flags=lib.MODE_FAST, member=None):
    import torch
    from rwrt_b200 import gridding
    ns, nz = len(slon), len(zwn)
    sl = torch.from_numpy(np.ascontiguousarray(slon)).to(dev)
    sa = torch.from_numpy(np.ascontiguousarray(slat)).to(dev)
    zw = torch.from_numpy(np.ascontiguousarray(zwn)).to(dev)
    mem = None if member is None else torch.from_numpy(np.ascontiguousarray(member, dtype=np.int32)).to(dev)
    g = gridding.new_grids(*cell, dev)
    ncx, ncy = gridding.grid_shape(*cell)
    acc = torch.zeros(3 * ns * nz, dtype=torch.int64, device=dev)
    tot = torch.zeros(1, dtype=torch.int64, device=dev)
    lib.check(L.rwrt_ray_sweep_grid(h, ns, lib.ptr(sl), lib.ptr(sa), lib.ptr(mem), nz, lib.ptr(zw), freq, nt, TSTEP, ttotal, 1e-6, 1e-6,
                                    1e-3, 0.1, flags, cell[0], cell[1], ncx, ncy, lib.ptr(g[0]), lib.ptr(g[1]), lib.ptr(g[2]), lib.ptr(acc),
                                    None, lib.ptr(tot), lib.stream_ptr()), "rwrt_ray_sweep_grid")
    torch.cuda.synchronize()
    return g, acc, int(tot.item())


def test_c5_full_size_sweep_is_deterministic_and_independent_of_sharding():
    """BASELINE configs[4] at FULL size through the north-star path: 1440 x 385 sources x k = 1..6 x 3 roots = 9 979 200 ray
    slots, 0.25-degree background, 60 days, gridded without any trajectory array (the reference layout would need 403 GB,
    wr.py:157-167).  The integer grids and the per-slot step counts must be identical from run to run and for the sources split
    into two strided shards (what `dist.sweep_sharded` gives two ranks), and consistent with each other."""
    import torch
    from rwrt_b200 import gridding
    nlon, nlat = 1440, 721
    slon_deg, slat_deg = synthetic.all_gridpoint_sources(nlon, nlat, 48.125)
    zwn = np.arange(1.0, 7.0)
    ns = len(slon_deg)
    assert 3 * ns * len(zwn) == 9979200
    L, h, d, dev = _setup(nlon, nlat, slon_deg[:1], slat_deg[:1], zwn)
    try:
        slon, slat = np.radians(slon_deg), np.radians(slat_deg)
        nt, ttotal, cell = 721, 60 * 86400.0, (0.25, 0.25)
        g1, a1, t1 = _sweep(L, h, d, dev, slon, slat, zwn, nt, ttotal, cell)
        assert t1 > 6.0e9 and int(a1.sum().item()) == t1
        samples = int(g1[0].sum().item())
        # every output row of a live ray is one sample unless it left the three longitude circles: samples <= steps, and a
        # ray that is alive for r rows took at least r accepted steps
        assert 3.0e9 < samples <= t1
        cs = gridding.checksum(g1)
        g2, a2, t2 = _sweep(L, h, d, dev, slon, slat, zwn, nt, ttotal, cell)
        assert t2 == t1 and torch.equal(a1, a2) and torch.equal(g1, g2)
        del g2, a2
        tot = torch.zeros_like(g1)
        steps = 0
        for r in range(2):
            gr, ar, tr = _sweep(L, h, d, dev, slon[r::2], slat[r::2], zwn, nt, ttotal, cell)
            tot += gr
            steps += tr
            assert torch.equal(ar.view(3, -1, 6), a1.view(3, ns, 6)[:, r::2, :])
        assert steps == t1 and torch.equal(tot, g1) and gridding.checksum(tot) == cs
    finally:
        L.rwrt_destroy(h)


def test_c3_full_size_frequency_sweep_three_gridding_paths_agree():
    """BASELINE configs[2] at full size: the 100 x 100 source matrix, k = 1..8, periods inf / 50 / 30 / 20 d, 30 days
    (240 000 ray slots per frequency).  Per frequency the grids from (a) the fused sweep, (b) the per-window ray records and
    (c) the trajectories of rwrt_ray_run_rk45 (rwrt_grid_accumulate) are the same integers; omega changes them."""
    import torch
    from rwrt_b200 import gridding
    nlon, nlat = 360, 181
    ix, iy = np.meshgrid(np.arange(100), np.arange(100), indexing="xy")
    slon_deg, slat_deg = (3.6 * ix.ravel()) % 360.0, -49.5 + 1.0 * iy.ravel()
    zwn = np.arange(1.0, 9.0)
    ns, nz = len(slon_deg), len(zwn)
    assert 3 * ns * nz == 240000
    L, h, d, dev = _setup(nlon, nlat, slon_deg, slat_deg, zwn)
    try:
        slon, slat = np.radians(slon_deg), np.radians(slat_deg)
        nt, ttotal, cell = 361, 30 * 86400.0, (1.0, 1.0)
        ncx, ncy = gridding.grid_shape(*cell)
        seen = []
        for period in (np.inf, 50.0, 30.0, 20.0):
            freq = 0.0 if np.isinf(period) else 2 * np.pi / (period * 86400.0)
            g_fused, acc, tot = _sweep(L, h, d, dev, slon, slat, zwn, nt, ttotal, cell, freq)
            g_rec, acc2, tot2 = _sweep(L, h, d, dev, slon, slat, zwn, nt, ttotal, cell, freq, flags=lib.MODE_FAST | lib.GRID_FROM_RECORDS)
            assert tot2 == tot and torch.equal(acc, acc2) and torch.equal(g_fused, g_rec)
            N = 3 * ns * nz
            out = torch.empty((7, nt, N), dtype=torch.float64, device=dev)
            sp = lib.stream_ptr()
            lib.check(L.rwrt_ray_init(h, ns, lib.ptr(d["slon"]), lib.ptr(d["slat"]), None, nz, lib.ptr(d["zwn"]), freq, nt, lib.ptr(out), sp))
            lib.check(L.rwrt_ray_run_rk45(h, ns, None, nz, nt, TSTEP, ttotal, 1e-6, 1e-6, 1e-3, 0.1, lib.MODE_FAST, lib.ptr(out), None, None,
                                          None, None, sp))
            g_pl = gridding.new_grids(*cell, dev)
            lib.check(L.rwrt_grid_accumulate(h, N, nt, lib.ptr(out), cell[0], cell[1], ncx, ncy, lib.ptr(g_pl[0]), lib.ptr(g_pl[1]),
                                             lib.ptr(g_pl[2]), sp))
            torch.cuda.synchronize()
            assert torch.equal(g_fused, g_pl), period
            # the grid counts are the finite (lon, lat, l) samples of rows 1.. inside the three longitude circles
            fin = torch.isfinite(out[0, 1:]) & torch.isfinite(out[1, 1:]) & ~torch.isnan(out[3, 1:])
            lon_deg = out[0, 1:][fin] * (180.0 / np.pi)
            assert int(g_fused[0].sum().item()) == int(((lon_deg >= -360.0) & (lon_deg < 720.0)).sum().item())
            del out, g_pl
            seen.append(g_fused[0].clone())
            assert tot > 2.0e7
        assert not torch.equal(seen[0], seen[1]) and not torch.equal(seen[1], seen[3])     # omega enters the t = 0 cubic (bs.py:1005)
    finally:
        L.rwrt_destroy(h)
