"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI
(`librwrt.so` via the ctypes facade), against the golden fixtures made from the unmodified
reference and against the C oracle on the same seeded inputs."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import parity
import rwrt_b200  # noqa: F401
from rwrt_b200 import lib, synthetic
from rwrt_b200.wr import WR
from oracle import oracle

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_wr(u, v, zwn, slon_rad, slat_rad, days, freq=0.0, tstep_h=2.0, xcyclic=True, **kw):
    nlon, nlat = u.shape
    w = WR(len(zwn), len(slon_rad), tstep_h * 3600.0, days * 86400.0, freq, nx=nlon, ny=nlat, **kw)
    w.bs.lon, w.bs.lat = synthetic.grid_coords(nlon, nlat)
    w.bs.u, w.bs.v = u, v
    w.bs.ready(xcyclic=xcyclic)
    w.set_zwn(zwn)
    w.source_lon = np.asarray(slon_rad, dtype=np.float64)
    w.source_lat = np.asarray(slat_rad, dtype=np.float64)
    return w


FIELD_CASES = {"jet_g1": ("jet_z", {}), "gc_g1": ("great_circle", {}),
               "clim_g1": ("clim", dict(nlon=144, nlat=73)), "clim_g2": ("clim", dict(nlon=360, nlat=181))}


@pytest.mark.parametrize("tag", list(FIELD_CASES))
def test_bs_ready_bit_exact(tag):
    """K1 vs bs.py:318-407: all 18 stacked fields and 9 diagnostic planes bit-identical."""
    g = parity.load("fields")
    fn, kw = FIELD_CASES[tag]
    u, v = getattr(synthetic, fn)(**kw)
    w = make_wr(u, v, [1.0], [0.1], [0.1], 1.0)
    assert w.bs.fields.shape == tuple(g[tag + "_fields_shape"])
    np.testing.assert_array_equal(w.bs.fields.reshape(-1)[g[tag + "_sample_idx"]], g[tag + "_sample_val"])
    assert sha(w.bs.fields) == str(g[tag + "_fields_sha"])
    for nm in ("q", "uxx", "uyy", "vxx", "vyy", "uxy", "vxy", "betam", "KS"):
        assert sha(getattr(w.bs, nm)) == str(g[f"{tag}_{nm}_sha"]), nm


def test_bs_ready_device_trig_tables():
    """coslat/sinlat = NULL: tables computed on the device; fields agree with the oracle to 1e-12."""
    import torch
    u, v = synthetic.clim(144, 73)
    lon, lat = synthetic.grid_coords(144, 73)
    L = lib.load()
    h = lib.vp()
    lib.check(L.rwrt_create(0, 144, 73, 1, 1, C.byref(h)))
    du, dv = torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()
    dlon, dlat = torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda()
    f18 = torch.empty((145, 73, 18), dtype=torch.float64, device="cuda")
    lib.check(L.rwrt_bs_ready(h, lib.ptr(du), lib.ptr(dv), lib.ptr(dlon), lib.ptr(dlat), None, None, lib.ptr(f18),
                              None, lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = oracle.bs_ready(u, v, lat, True)
    got = f18.cpu().numpy()
    # device cos/sin differ from libm in the last ulp; q divides by cos(lat) next to the poles, so compare
    # each field plane-wise (max abs difference relative to the plane's magnitude)
    for k in range(18):
        scale = np.max(np.abs(ref[..., k])) + 1e-300
        assert np.max(np.abs(got[..., k] - ref[..., k])) / scale < 1e-10, k
    L.rwrt_destroy(h)


def test_sampler_rhs_vs_reference():
    """rwrt_sample_mercator / rwrt_rhs vs bs.py:781-887 and wr.py:492-556 outputs of the reference."""
    import torch
    g = parity.load("pointwise")
    u, v = synthetic.clim(144, 73)
    w = make_wr(u, v, [1.0], [0.1], [0.1], 1.0)
    got18 = w.bs.cal_bs_mercator_point(g["plon"], g["plat"], mode="numpy").T      # all 18 rows (bs.py:884-887)
    ref18 = g["merc18"]
    assert got18.shape == ref18.shape == (len(g["plon"]), 18)
    assert np.array_equal(np.isnan(got18), np.isnan(ref18))
    ok = np.isfinite(ref18)
    assert np.max(np.abs(got18[ok] - ref18[ok]) / np.maximum(np.abs(ref18[ok]), 1e-6)) < 1e-8
    got, ref = got18[:, :12], g["merc"]
    y = torch.from_numpy(np.ascontiguousarray(g["y"].T)).cuda()       # (5, n)
    for flags in (lib.MODE_FAST, lib.MODE_STRICT):
        dy = torch.empty_like(y)
        lib.check(lib.load().rwrt_rhs(w.bs._h, y.shape[1], lib.ptr(y), None, lib.ptr(dy), flags, lib.stream_ptr()))
        got = dy.cpu().numpy().T
        ref = g["dy"]
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        ok = np.isfinite(ref)
        assert np.max(np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-12)) < 1e-9
    # and against the oracle, tighter (same libm class)
    orc = oracle.rhs(oracle.bs_ready(u, v, w.bs.lat, True), w.bs.lon, w.bs.lat, g["y"])
    ok = np.isfinite(orc)
    assert np.max(np.abs(got[ok] - orc[ok]) / np.maximum(np.abs(orc[ok]), 1e-12)) < 1e-10


def adversarial_longitudes():
    """Longitudes that stress np.remainder(lon, 2 pi) (bs.py:519, interpolation.py:80): both signs, many turns, exact
    multiples of the double 2 pi and their neighbours, +-0, values that round to 2 pi, huge / inf / NaN."""
    tp = 2.0 * np.pi
    rng = np.random.default_rng(31415)
    mult = np.arange(-40, 41, dtype=np.float64) * tp
    near = np.concatenate([np.nextafter(mult, np.inf), np.nextafter(mult, -np.inf), mult])
    tiny = np.array([0.0, -0.0, 5e-324, -5e-324, 1e-300, -1e-300, 1e-20, -1e-20, np.nextafter(tp, 0.0), -np.nextafter(tp, 0.0)])
    big = np.array([1e9, -1e9, 1e9 + 1.0, 3.7e12, -8.1e15, 1.2e300, np.inf, -np.inf, np.nan])
    return np.concatenate([rng.uniform(-60.0, 60.0, 3000), rng.uniform(-1e6, 1e6, 500), near, tiny, big,
                           mult[:, None].repeat(3, 1).ravel() + rng.uniform(0.0, tp, 243)])


def test_rhs_on_adversarial_longitudes():
    """The longitude wrap of the sampler (round 2: evaluated on the signed value with one fused multiply-add, general
    routine only outside [0, 2 pi)) against the oracle's fmod-based np.remainder on longitudes chosen to break it: a
    wrong quotient or sign rule puts the sample in another cell (O(1) differences), so the comparison is sharp."""
    import torch
    u, v = synthetic.clim(144, 73)
    w = make_wr(u, v, [1.0], [0.1], [0.1], 1.0)
    lon = adversarial_longitudes()
    rng = np.random.default_rng(27)
    n = lon.size
    ys = np.stack([lon, np.radians(rng.uniform(-80.0, 80.0, n)), rng.uniform(1.0, 6.0, n), rng.uniform(-8.0, 8.0, n),
                   np.ones(n)], axis=1)                                  # (lon, lat, k, l, A) as the ray state holds them
    orc = oracle.rhs(oracle.bs_ready(u, v, w.bs.lat, True), w.bs.lon, w.bs.lat, ys)
    assert np.isfinite(orc).all(axis=1).sum() > 0.9 * n                  # the case samples real cells, not NaN
    y = torch.from_numpy(np.ascontiguousarray(ys.T)).cuda()
    for flags in (lib.MODE_FAST, lib.MODE_STRICT):
        dy = torch.empty_like(y)
        lib.check(lib.load().rwrt_rhs(w.bs._h, n, lib.ptr(y), None, lib.ptr(dy), flags, lib.stream_ptr()))
        got = dy.cpu().numpy().T
        assert np.array_equal(np.isnan(got), np.isnan(orc))
        ok = np.isfinite(orc)
        assert np.max(np.abs(got[ok] - orc[ok]) / np.maximum(np.abs(orc[ok]), 1e-12)) < 1e-9


TRAJ = ["c1_jet", "great_circle", "clim_g1", "clim_g1_freq", "clim_g2"]


def run_gpu(g, strict):
    u, v = parity.background(g)
    w = make_wr(u, v, g["zwn"], g["src_lon_rad"], g["src_lat_rad"], float(g["days"]), float(g["freq"]),
                float(g["tstep_h"]), rtol=float(g["rtol"]), atol=float(g["atol"]),
                MinStepFactor=float(g["min_step_factor"]))
    assert sha(w.bs.fields) == str(g["fields_sha"])
    w.ray_run(mode="numpy", root_method="numpy", inte_method="rk45", strict=strict)
    return w


@pytest.mark.parametrize("name", TRAJ)
def test_initial_state_vs_reference(name):
    """K2 vs wr.py:344-395: slot layout identical everywhere; values bit-identical."""
    g = parity.load(name)
    w = run_gpu(g, strict=True)
    for n in parity.NAMES:
        got, ref = getattr(w, n)[0], g[n][0]
        assert np.array_equal(np.isnan(got), np.isnan(ref)), n
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=0, equal_nan=True, err_msg=n)
    exact = np.mean([(getattr(w, n)[0] == g[n][0])[np.isfinite(g[n][0])].mean() for n in ("rmwn", "rug", "rvg")])
    assert exact > 0.8     # device sin/cos differ from NumPy's in the last ulp -> roots differ in the last digits


@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
@pytest.mark.parametrize("name", TRAJ)
def test_trajectories_vs_reference(name, strict):
    """K3 vs wr.py:767-887 on the reference's own outputs: < 1e-6 deg, equal step counts (non-grazing)."""
    g = parity.load(name)
    w = run_gpu(g, strict)
    got = {n: getattr(w, n) for n in parity.NAMES}
    ref = {n: g[n] for n in parity.NAMES}
    st = parity.compare(got, w.accepted, ref, g["accepted"], g["nongrazing"], f"{name}/{'strict' if strict else 'fast'}")
    assert w.total_accepted == int(w.accepted.sum())
    live = ~np.isnan(g["rmwn"][0])
    assert (w.status[~live] == lib.RAY_NO_ROOT).all() and (w.status[live] != lib.RAY_NO_ROOT).all()
    parity.record(f"fixture {name} (reference-generated) vs GPU, {'strict' if strict else 'fast'}", st)
    print(name, st)


oracle_run = parity.oracle_run


_MEDIUM = {}


def _medium_oracle():
    if not _MEDIUM:
        u, v = synthetic.clim(360, 181)
        sl, sa = synthetic.random_sources(600, 99)
        _MEDIUM["r"] = parity.oracle_nongrazing(u, v, np.radians(sl), np.radians(sa), np.arange(1.0, 7.0), 10.0)
    return _MEDIUM["r"]


@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_vs_oracle_medium(strict):
    """1-degree Clim, 600 seeded sources x k=1..6, 10 days: CUDA vs the C oracle on the same inputs;
    the non-grazing set comes from five oracle runs (sources moved by +-1 ulp in lon / lat)."""
    u, v = synthetic.clim(360, 181)
    sl, sa = synthetic.random_sources(600, 99)
    slon, slat = np.radians(sl), np.radians(sa)
    zwn = np.arange(1.0, 7.0)
    ref, acc, ng = _medium_oracle()
    w = make_wr(u, v, zwn, slon, slat, 10.0)
    w.ray_run(mode="numpy", root_method="numpy", inte_method="rk45", strict=strict)
    got = {n: getattr(w, n) for n in parity.NAMES}
    st = parity.compare(got, w.accepted, ref, acc, ng, "medium", max_outliers=int(ng.sum()) // 2000)
    parity.record(f"1 deg Clim, 600 sources x k=1..6, 10 d vs oracle, {'strict' if strict else 'fast'}", st)
    print(st)
    assert st["nongrazing"] > 0.7 * st["live"]
    # whole-batch sanity: total accepted steps within 0.5 % of the oracle's even counting grazing rays
    assert abs(st["steps_total"] - st["steps_total_ref"]) < 1e-2 * st["steps_total_ref"]


@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_noncyclic_grid_clamped_cells(strict):
    """xcyclic=False (the reference's own default, bs.py:318): no wrap column, so every position in the last longitude
    cell has x1 clamped onto x0 (interpolation.py:116-119) -- the cells that the cooperative gather does not put on its
    fetch list (their owner lane copies them itself).  2.5-degree Clim, 240 sources of which half start within 12
    degrees west of the seam, k = 1..4, 8 days, against the oracle on its non-grazing set."""
    u, v = synthetic.clim(144, 73)
    rng = np.random.default_rng(2718)
    sl = np.concatenate([rng.uniform(346.0, 359.9, 120), rng.uniform(0.0, 346.0, 120)])
    sa = rng.uniform(-60.0, 60.0, 240)
    slon, slat = np.radians(sl), np.radians(sa)
    zwn = np.arange(1.0, 5.0)
    ref, acc, ng = parity.oracle_nongrazing(u, v, slon, slat, zwn, 8.0, xcyclic=False)
    # the case does what it is for: live rays spend output rows inside the clamped cell column
    seam = np.radians(357.5)
    with np.errstate(invalid="ignore"):
        in_seam = (np.mod(ref["rlon"], 2 * np.pi) >= seam) & np.isfinite(ref["rmwn"])
    assert in_seam[1:].sum() > 2000
    w = make_wr(u, v, zwn, slon, slat, 8.0, xcyclic=False)
    w.ray_run(mode="numpy", root_method="numpy", inte_method="rk45", strict=strict)
    got = {n: getattr(w, n) for n in parity.NAMES}
    st = parity.compare(got, w.accepted, ref, acc, ng, "noncyclic", max_outliers=0)
    parity.record(f"2.5 deg Clim, xcyclic=False, 240 sources x k=1..4, 8 d vs oracle, {'strict' if strict else 'fast'}", st)
    print(st)
    assert st["nongrazing"] > 0.5 * st["live"]
    assert (ng & in_seam[1:].any(axis=0)).sum() > 50              # ... and non-grazing rays among them


def test_batch_composition_independence():
    """A ray's result does not depend on which other rays share its batch (SURVEY.md probe9):
    running a slice of the sources alone reproduces the same bits -- what makes source sharding exact."""
    u, v = synthetic.clim(360, 181)
    sl, sa = synthetic.random_sources(300, 5)
    slon, slat = np.radians(sl), np.radians(sa)
    zwn = np.array([1.0, 3.0, 5.0])
    w = make_wr(u, v, zwn, slon, slat, 8.0)
    w.ray_run(mode="numpy", inte_method="rk45")
    w2 = make_wr(u, v, zwn, slon[100:180], slat[100:180], 8.0)
    w2.ray_run(mode="numpy", inte_method="rk45")
    for n in parity.NAMES:
        assert np.array_equal(getattr(w, n)[:, :, 100:180, :], getattr(w2, n), equal_nan=True), n
    assert np.array_equal(w.accepted[:, 100:180, :], w2.accepted)
    # idempotence: running again reproduces the same bits (no order-dependent state in the engine)
    first = {n: getattr(w, n).copy() for n in parity.NAMES}
    first_acc = w.accepted.copy()
    w.ray_run(mode="numpy", inte_method="rk45")
    for n in parity.NAMES:
        assert np.array_equal(first[n], getattr(w, n), equal_nan=True), n
    # time windows (rays parked at a window end, state saved and restored) are exact as well
    import os
    os.environ["RWRT_WINDOWS"] = "7"
    try:
        w.ray_run(mode="numpy", inte_method="rk45")
    finally:
        del os.environ["RWRT_WINDOWS"]
    for n in parity.NAMES:
        assert np.array_equal(first[n], getattr(w, n), equal_nan=True), ("windows", n)
    assert np.array_equal(first_acc, w.accepted)


def test_c_abi_host_entry_equals_facade(monkeypatch):
    """rwrt_real2d_host (host pointers, H2D/D2H inside; integration in time windows whose ray records -- or, with
    RWRT_HOST_EXPAND=0, finished rows -- are copied back while the next window runs) returns what the facade returns,
    with 1 ... 30 windows (more windows than staging buffers: every buffer is reused)."""
    u, v = synthetic.clim(144, 73)
    ns = 120
    sl, sa = synthetic.random_sources(ns, 3)
    slon, slat = np.radians(sl), np.radians(sa)
    zwn = np.array([1.0, 2.0, 4.0])
    days = 5.0
    w = make_wr(u, v, zwn, slon, slat, days)
    w.ray_run(mode="numpy", inte_method="rk45")
    # the facade delivers through rwrt_ray_run_rk45_d2h: device planes and host planes of one run are the same bits
    assert np.array_equal(w.d_out.cpu().numpy(), w._host_planes(), equal_nan=True)
    lon, lat = synthetic.grid_coords(144, 73)
    L = lib.load()
    nt = w.nt
    cl, sn = np.cos(lat), np.sin(lat)
    import torch
    pinned_t = torch.empty((7, nt, 3, ns, 3), dtype=torch.float64, pin_memory=True)
    # windows, RWRT_HOST_EXPAND (1: packed rows travel, host threads expand; 0: the device expands), RWRT_HOST_ROUTE
    # (which windows are copied straight into PINNED caller planes instead: none / all / every other / by backlog)
    for chunk_mb, expand, route in (("1", "1", "hx"), ("4", "1", "hx"), ("9", "1", "hx"), ("30", "1", "hx"), ("1", "0", "hx"),
                                    ("4", "0", "hx"), ("9", "1", "direct"), ("30", "1", "alt"), ("9", "1", "alt"), ("30", "1", "auto"),
                                    ("2", "1", "alt")):
        monkeypatch.setenv("RWRT_HOST_WINDOWS", chunk_mb)
        monkeypatch.setenv("RWRT_HOST_EXPAND", expand)
        monkeypatch.setenv("RWRT_HOST_ROUTE", route)
        h = lib.vp()
        lib.check(L.rwrt_create(0, 144, 73, 1, 1, C.byref(h)))
        out = np.full((7, nt, 3, ns, 3), -7.0) if route == "hx" else pinned_t.numpy()   # pageable memory never goes direct
        out.fill(-7.0)
        acc = np.full(3 * ns * 3, -1, dtype=np.int64)
        status = np.full(3 * ns * 3, -1, dtype=np.int8)
        total = np.zeros(1, dtype=np.int64)
        lib.check(L.rwrt_real2d_host(h, lib.ptr(u), lib.ptr(v), lib.ptr(lon), lib.ptr(lat), lib.ptr(cl), lib.ptr(sn), ns,
                                     lib.ptr(slon), lib.ptr(slat), None, 3, lib.ptr(zwn), 0.0, nt, 7200.0,
                                     days * 86400.0, 1e-6, 1e-6, 1e-3, 0.1, lib.MODE_FAST, lib.ptr(out), lib.ptr(acc),
                                     lib.ptr(status), lib.ptr(total), None), "rwrt_real2d_host")
        for k, n in enumerate(parity.NAMES):
            assert np.array_equal(out[k], getattr(w, n), equal_nan=True), (chunk_mb, expand, route, n)
        assert np.array_equal(acc.reshape(3, ns, 3), w.accepted) and int(total[0]) == w.total_accepted
        assert np.array_equal(status.reshape(3, ns, 3), w.status)
        hs = (C.c_double * 5)()
        lib.check(L.rwrt_last_host_stats(h, hs))
        if expand == "1" and route == "direct":
            assert hs[4] == hs[2] > 0           # every window went straight into the pinned planes
        if expand == "1" and route == "alt" and hs[2] >= 2:
            assert 0 < hs[4] < hs[2]
        if route == "hx":
            assert hs[4] == 0
        L.rwrt_destroy(h)
    # the facade's delivery (rwrt_ray_run_rk45_d2h: full planes on the device) with pinned result planes and mixed routes
    for route in ("alt", "direct"):
        monkeypatch.setenv("RWRT_HOST_ROUTE", route)
        monkeypatch.setenv("RWRT_HOST_WINDOWS", "7")
        w2 = make_wr(u, v, zwn, slon, slat, days)
        w2.pin_results = True
        w2.ray_run(mode="numpy", inte_method="rk45")
        assert np.array_equal(w2._host_planes(), w._host_planes(), equal_nan=True), route
        assert np.array_equal(w2.d_out.cpu().numpy(), w._host_planes(), equal_nan=True), route


def test_empty_and_error_paths():
    """Edge cases: zero sources; call-order and argument errors surface as exceptions, not crashes."""
    u, v = synthetic.jet_z()
    w = make_wr(u, v, [1.0, 2.0], np.zeros(0), np.zeros(0), 2.0)
    w.ray_run(mode="numpy", inte_method="rk45")
    assert w.rlon.shape == (25, 3, 0, 2) and w.total_accepted == 0
    with pytest.raises(NotImplementedError):
        w.ray_run(mode="original", inte_method="rk45")
    L = lib.load()
    h = lib.vp()
    lib.check(L.rwrt_create(0, 144, 73, 1, 1, C.byref(h)))
    rc = L.rwrt_ray_init(h, 1, None, None, None, 1, None, 0.0, 2, None, None)
    assert rc != 0 and b"NULL" in L.rwrt_last_error() or b"before" in L.rwrt_last_error()
    L.rwrt_destroy(h)
    with pytest.raises(ValueError):
        lib.check(L.rwrt_create(0, 1, 2, 1, 1, C.byref(h)))


# termination at the pole / speed cut-off: tests/test_gpu_parity_full.py::test_termination_exact_on_nongrazing_set (exact, both builds)


def test_ensemble_members_match_single_runs():
    """BASELINE configs[3] in miniature: an ensemble of background states batched in one handle
    (nmember > 1, per-source member index) gives, bit for bit, what separate single-member runs give."""
    import torch
    members = [synthetic.ensemble_member(e, 144, 73) for e in range(3)]
    lon, lat = synthetic.grid_coords(144, 73)
    sl, sa = synthetic.random_sources(30, 9)
    slon, slat = np.radians(sl), np.radians(sa)
    zwn = np.array([1.0, 3.0])
    days, nt = 6.0, 73
    member_of = np.arange(30, dtype=np.int32) % 3
    L = lib.load()
    h = lib.vp()
    lib.check(L.rwrt_create(0, 144, 73, 1, 3, C.byref(h)))
    dev = torch.device("cuda", 0)
    U = torch.from_numpy(np.stack([m[0] for m in members])).to(dev)
    V = torch.from_numpy(np.stack([m[1] for m in members])).to(dev)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         dict(lon=lon, lat=lat, cl=np.cos(lat), sl=np.sin(lat), slon=slon, slat=slat, zwn=zwn, mem=member_of).items()}
    lib.check(L.rwrt_bs_ready(h, lib.ptr(U), lib.ptr(V), lib.ptr(d["lon"]), lib.ptr(d["lat"]), lib.ptr(d["cl"]),
                              lib.ptr(d["sl"]), None, None, lib.stream_ptr()))
    N = 3 * 30 * 2
    out = torch.empty((7, nt, N), dtype=torch.float64, device=dev)
    acc = torch.zeros(N, dtype=torch.int64, device=dev)
    lib.check(L.rwrt_ray_init(h, 30, lib.ptr(d["slon"]), lib.ptr(d["slat"]), lib.ptr(d["mem"]), 2, lib.ptr(d["zwn"]),
                              0.0, nt, lib.ptr(out), lib.stream_ptr()))
    lib.check(L.rwrt_ray_run_rk45(h, 30, lib.ptr(d["mem"]), 2, nt, 7200.0, days * 86400.0, 1e-6, 1e-6, 1e-3, 0.1,
                                  lib.MODE_FAST, lib.ptr(out), lib.ptr(acc), None, None, None, lib.stream_ptr()))
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(7, nt, 3, 30, 2)
    got_acc = acc.cpu().numpy().reshape(3, 30, 2)
    L.rwrt_destroy(h)
    for e in range(3):
        idx = np.where(member_of == e)[0]
        w = make_wr(members[e][0], members[e][1], zwn, slon[idx], slat[idx], days)
        w.ray_run(mode="numpy", inte_method="rk45")
        for k, n in enumerate(parity.NAMES):
            assert np.array_equal(got[k][:, :, idx, :], getattr(w, n), equal_nan=True), (e, n)
        assert np.array_equal(got_acc[:, idx, :], w.accepted)


def test_quarter_degree_grid_and_gather_cross_check():
    """BASELINE configs[4] in miniature: 0.25-degree grid (100 MB point table), a few hundred sources, against
    the oracle on its non-grazing set; also checks that the two integrators (warp-cooperative gather / per-lane
    gather: independent implementations of the sampler) agree bit for bit on a 1-degree and on the 0.25-degree case."""
    import os
    u, v = synthetic.clim(1440, 721)
    sl, sa = synthetic.random_sources(120, 123)
    slon, slat = np.radians(sl), np.radians(sa)
    zwn = np.array([2.0, 4.0])
    ref, acc, ng = parity.oracle_nongrazing(u, v, slon, slat, zwn, 6.0)
    w = make_wr(u, v, zwn, slon, slat, 6.0)
    w.ray_run(mode="numpy", inte_method="rk45", strict=True)
    got = {n: getattr(w, n) for n in parity.NAMES}
    st = parity.compare(got, w.accepted, ref, acc, ng, "quarter-degree")
    parity.record("0.25 deg Clim, 120 sources x k=2,4, 6 d vs oracle, strict", st)
    assert st["nongrazing"] > 0.5 * st["live"]
    # same rays through both kernels on a grid where both exist
    u2, v2 = synthetic.clim(360, 181)
    w1 = make_wr(u2, v2, zwn, slon, slat, 6.0)
    w1.ray_run(mode="numpy", inte_method="rk45", strict=True)
    os.environ["RWRT_PER_LANE_GATHER"] = "1"
    try:
        w2 = make_wr(u2, v2, zwn, slon, slat, 6.0)
        w2.ray_run(mode="numpy", inte_method="rk45", strict=True)
    finally:
        del os.environ["RWRT_PER_LANE_GATHER"]
    for n in parity.NAMES:
        assert np.array_equal(getattr(w1, n), getattr(w2, n), equal_nan=True), n
    assert np.array_equal(w1.accepted, w2.accepted)
    os.environ["RWRT_PER_LANE_GATHER"] = "1"
    try:
        w3 = make_wr(u, v, zwn, slon, slat, 6.0)
        w3.ray_run(mode="numpy", inte_method="rk45", strict=True)
    finally:
        del os.environ["RWRT_PER_LANE_GATHER"]
    for n in parity.NAMES:
        assert np.array_equal(getattr(w, n), getattr(w3, n), equal_nan=True), ("quarter-degree", n)


def test_frequency_sweep_with_gridding():
    """BASELINE configs[2] in miniature: periods inf / 50 / 30 / 20 days x k = 1..8 from a source matrix,
    gridded ray counts bit-exact against the NumPy definition; omega enters only the t = 0 cubic (bs.py:1005)."""
    from rwrt_b200 import gridding
    from oracle import gridding as ogrid
    u, v = synthetic.clim(360, 181)
    zwn = np.arange(1.0, 9.0)
    total = None
    first_l0 = None
    for period in (np.inf, 50.0, 30.0, 20.0):
        freq = 0.0 if np.isinf(period) else 2 * np.pi / (period * 86400.0)
        w = WR(8, 100, 7200.0, 6 * 86400.0, freq, nx=360, ny=181)
        w.bs.lon, w.bs.lat = synthetic.grid_coords(360, 181)
        w.bs.u, w.bs.v = u, v
        w.bs.ready(xcyclic=True)
        w.set_zwn(zwn)
        w.set_source_matrix(0.0, -49.5, 36.0, 10.0, 10, 10)
        w.ray_run(mode="numpy", inte_method="rk45")
        grids = gridding.accumulate(w, 2.0, 2.0)
        out7 = np.stack([getattr(w, n) for n in parity.NAMES])
        ref = ogrid.accumulate(out7, 2.0, 2.0)
        for g, r in zip(grids, ref):
            assert np.array_equal(g.cpu().numpy(), r)
        if first_l0 is None:
            first_l0 = w.rmwn[0].copy()
        else:
            assert not np.array_equal(first_l0, w.rmwn[0], equal_nan=True)   # omega changes the roots
        total = ref[0] if total is None else total + ref[0]
    assert total.sum() > 1e5


@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
@pytest.mark.parametrize("name", ["rk4_jet", "rk4_clim_g1"])
def test_rk4_path_vs_reference(name, strict):
    """rwrt_ray_run_rk4 vs wr.py:702-765 on the reference's outputs (main_wr.py's default inte_method='')."""
    g = parity.load(name)
    u, v = parity.background(g)
    w = make_wr(u, v, g["zwn"], g["src_lon_rad"], g["src_lat_rad"], float(g["days"]), float(g["freq"]),
                float(g["tstep_h"]))
    w.ray_run(mode="numpy", root_method="numpy", inte_method="", strict=strict)
    got = {n: getattr(w, n) for n in parity.NAMES}
    ref = {n: g[n] for n in parity.NAMES}
    st = parity.compare(got, None, ref, g["accepted"], g["nongrazing"], name)
    parity.record(f"fixture {name} (reference-generated, fixed-step RK4) vs GPU, {'strict' if strict else 'fast'}", st)
    assert w.accepted is None and w.total_accepted == 0
    dead = np.isnan(g["rmwn"][0])
    assert np.isnan(w.rlon[1:, dead]).all()
    print(name, st)


def test_rk4_cooperative_kernel_equals_per_lane_kernel(monkeypatch):
    """The RK4 product kernel (warp-cooperative gather, FSAL: the evaluation at y_new serves the output and the next
    k1) against the straightforward per-lane kernel (5 right-hand sides per step, wr.py:583-622 literally): bit for
    bit in the strict build, NaN patterns identical and < 1e-9 deg apart in the fast build -- 20 000 slots, 1 deg, 20 days."""
    u, v = synthetic.clim(360, 181)
    sl, sa = synthetic.random_sources(1111, 17, max_abs_lat_deg=85.0)
    slon, slat = np.radians(sl), np.radians(sa)
    zwn = np.arange(1.0, 7.0)
    for strict in (True, False):
        monkeypatch.setenv("RWRT_PER_LANE_GATHER", "0")      # force the cooperative kernel (the strict default is per-lane)
        w = make_wr(u, v, zwn, slon, slat, 20.0)
        w.ray_run(mode="numpy", root_method="numpy", inte_method="", strict=strict)
        monkeypatch.setenv("RWRT_PER_LANE_GATHER", "1")
        try:
            w2 = make_wr(u, v, zwn, slon, slat, 20.0)
            w2.ray_run(mode="numpy", root_method="numpy", inte_method="", strict=strict)
        finally:
            monkeypatch.delenv("RWRT_PER_LANE_GATHER")
        assert np.isfinite(w.rlon).mean() > 0.3
        for n in parity.NAMES:
            a, b = getattr(w, n), getattr(w2, n)
            assert np.array_equal(np.isnan(a), np.isnan(b)), (strict, n)
            if strict:
                assert np.array_equal(a, b, equal_nan=True), n
        if not strict:
            with np.errstate(invalid="ignore"):
                dev = np.degrees(np.nanmax(np.abs(w.rlon[:61] - w2.rlon[:61])))     # first 5 days: before chaos amplifies rounding
            assert dev < 1e-9, dev


def test_driver_entry_real2d_hnf_with_netcdf(tmp_path):
    """The reference's driver call `real2d_hnf(**parameters)` (main_wr.py:31-100) end to end: NetCDF input with
    descending latitudes (flipped on read, bs.py:251-256), bsfile and ray file written and read back, and the
    result identical to driving the WR object directly."""
    from scipy.io import netcdf_file
    from rwrt_b200 import main_wr, ncio
    u, v = synthetic.clim(144, 73)
    lat_deg = np.linspace(90, -90, 73).astype("float32")
    lon_deg = (np.arange(144) * 2.5).astype("float32")
    src = str(tmp_path / "basic_flow.nc")
    ds = netcdf_file(src, "w", version=2)
    ds.createDimension("lat", 73)
    ds.createDimension("lon", 144)
    ds.createVariable("lat", "f4", ("lat",))[:] = lat_deg
    ds.createVariable("lon", "f4", ("lon",))[:] = lon_deg
    ds.createVariable("u", "f4", ("lat", "lon"))[:] = u.T[::-1]
    ds.createVariable("v", "f4", ("lat", "lon"))[:] = v.T[::-1]
    ds.close()
    par = dict(main_wr.parameters)
    par.update(inputuv=src, bsfile=str(tmp_path / "bs_out.nc"), ncfile=str(tmp_path / "rays_out.nc"), ttotal=5.0,
               zwn=np.array([1., 2., 3.]), nzwn=3, nnx=5, nny=4, SW_lon=70., SW_lat=-4., dlon=4, dlat=2)
    w = main_wr.real2d_hnf(**par, quiet=True)
    assert w.rlon.shape == (61, 3, 20, 3)
    # the same through the object API with the arrays injected (what the oracle harness does)
    d = ncio.read_uv(src, 144, 73)
    assert np.array_equal(d["u"], u) and np.array_equal(d["v"], v)
    w2 = WR(3, 20, 7200.0, 5 * 86400.0, 0.0, nx=144, ny=73)
    w2.bs.lon, w2.bs.lat, w2.bs.u, w2.bs.v = d["lon"], d["lat"], d["u"], d["v"]
    w2.bs.ready(xcyclic=True)
    w2.set_zwn(par["zwn"])
    w2.set_source_matrix(70., -4., 4, 2, 5, 4)
    w2.ray_run(mode="numpy", root_method="numpy", inte_method="rk45")
    for n in parity.NAMES:
        assert np.array_equal(getattr(w, n), getattr(w2, n), equal_nan=True), n
    # the ray file: degrees for rlon / rlat (wr.py:947-948), radians in memory; dims (time, root, source, zwn)
    rf, close = ncio._open_read(par["ncfile"] if os.path.exists(par["ncfile"]) else par["ncfile"] + ".npz")
    try:
        assert rf["rlon"].shape == (61, 3, 20, 3)
        assert np.allclose(rf["rlon"], np.degrees(w.rlon), rtol=0, atol=1e-9, equal_nan=True)
        assert np.array_equal(np.isnan(rf["rmwn"]), np.isnan(w.rmwn))
        assert np.array_equal(rf["rug"], w.rug, equal_nan=True)
    finally:
        close()
    bf, close = ncio._open_read(par["bsfile"] if os.path.exists(par["bsfile"]) else par["bsfile"] + ".npz")
    try:
        assert "u" in bf and "qy" in bf or len(bf) >= 20     # 23 variables (bs.py:461-511)
    finally:
        close()
