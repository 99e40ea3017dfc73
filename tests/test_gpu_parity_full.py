"""GPU parity at BASELINE's own sizes and durations, and on the reference's own root fixtures.

  * `cal_ky` / `change_roots_order` / LAPACK eigenvalue order (bs.py:38-40, 942-1040) ON THE GPU against the
    reference-generated slot layouts of tests/golden/pointwise.npz (5 x 3 000 cells) and ky_frac.npz (6 x 2 000 cells,
    fractional zonal wavenumbers): identical slot layout on every cell, values bit for bit;
  * seeded samples of each BASELINE config at its full duration against the C oracle (wr.py:767-887), non-grazing
    masks from the oracle's four +-1 ulp runs, strict and fast builds: C2 (1 deg, k = 1..6, 30 d), C5 (0.25 deg, 60 d),
    C3 (4 frequencies x k = 1..8), C4 (ensemble members, batched in one handle);
  * termination (wr.py:838-850, 97-112) in a violent flow, exact on the non-grazing set in BOTH builds (the fast build
    uses a threshold form of the haversine test).
Every comparison appends its statistics (incl. how many GRAZING rays are within tolerance anyway) to
gpurun_out/parity_table.jsonl; profiles/r2_parity_table.md is that file, tabulated.
"""
import ctypes as C

import numpy as np
import pytest

import parity
import rwrt_b200  # noqa: F401
from rwrt_b200 import lib, synthetic
from rwrt_b200.wr import WR

pytestmark = pytest.mark.gpu

BUILDS = [pytest.param(True, id="strict"), pytest.param(False, id="fast")]


# ---------------------------------------------------------------------------------------------
# root solve on the reference's own fixtures
# ---------------------------------------------------------------------------------------------
def _gpu_cal_ky(fu, fv, fqx, fqy, freq, k):
    """rwrt_ray_init on a synthetic table in which source i samples exactly (fu, fv, fqx, fqy)[i]: grid points 2i and
    2i + 1 of the three latitude rows hold cell i's values, the source sits midway between them on the equator row, and
    the grid spacings are powers of two -- so the bilinear weights are exactly (0.5, 0.5, 0, 0), cos = 1, tan = sin = 0
    and the Mercator quantities equal the raw fields bit for bit (bs.py:856-883)."""
    import torch
    n = len(fu)
    nlon, nlat = 2 * n, 3
    dlon = 2.0 ** -11
    assert nlon * dlon < 2 * np.pi
    lon = np.arange(nlon) * dlon
    lat = np.array([-0.5, 0.0, 0.5])
    f18 = np.zeros((nlon, nlat, 18))
    for col, a in ((0, fu), (1, fv), (6, fqx), (7, fqy)):        # bs.py:349-368: u, v, ..., qx (6), qy (7)
        f18[:, :, col] = np.repeat(a, 2)[:, None]
    L = lib.load()
    h = lib.vp()
    lib.check(L.rwrt_create(0, nlon, nlat, 0, 1, C.byref(h)))
    try:
        dev = torch.device("cuda", 0)
        d = {k_: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k_, v in
             dict(f18=f18, lon=lon, lat=lat, slon=(2 * np.arange(n) + 0.5) * dlon, slat=np.zeros(n), zwn=np.array([float(k)])).items()}
        lib.check(L.rwrt_bs_set_fields(h, lib.ptr(d["f18"]), lib.ptr(d["lon"]), lib.ptr(d["lat"]), lib.stream_ptr()))
        out = torch.empty((7, 1, 3, n, 1), dtype=torch.float64, device=dev)
        lib.check(L.rwrt_ray_init(h, n, lib.ptr(d["slon"]), lib.ptr(d["slat"]), None, 1, lib.ptr(d["zwn"]), float(freq), 1,
                                  lib.ptr(out), lib.stream_ptr()))
        torch.cuda.synchronize()
        return out.cpu().numpy()[3, 0, :, :, 0].T.copy()          # (n, 3) slot-ordered l, like cal_ky's return
    finally:
        L.rwrt_destroy(h)


@pytest.mark.parametrize("fixture", ["pointwise", "ky_frac"])
def test_gpu_cal_ky_on_reference_root_fixtures(fixture):
    g = parity.load(fixture)
    keys = [k for k in g.files if k.startswith("ky_")]
    assert len(keys) >= 5
    cells = three = 0
    for key in keys:
        kz, fr = g["kyarg_" + key[3:]]
        got = _gpu_cal_ky(g["fu"], g["fv"], g["fqx"], g["fqy"], fr, kz)
        ref = g[key]
        assert np.array_equal(np.isnan(got), np.isnan(ref)), f"{fixture}/{key}: slot layout differs on {int((np.isnan(got) != np.isnan(ref)).any(axis=1).sum())} cells"
        assert np.array_equal(got, ref, equal_nan=True), f"{fixture}/{key}: {int((got != ref)[np.isfinite(ref)].sum())} root values differ in the last bits"
        cells += len(ref)
        three += int(np.isfinite(ref).all(axis=1).sum())
    assert cells >= 12000 and three > 300        # the order-dependent three-real-root case is well covered
    parity.record(f"cal_ky GPU vs reference {fixture}", dict(cells=cells, three_root_cells=three, layout_mismatch=0, value_mismatch=0))


# ---------------------------------------------------------------------------------------------
# BASELINE configs at their own durations, seeded samples, against the oracle
# ---------------------------------------------------------------------------------------------
def make_wr(u, v, zwn, slon_rad, slat_rad, days, freq=0.0):
    nlon, nlat = u.shape
    w = WR(len(zwn), len(slon_rad), 7200.0, days * 86400.0, freq, nx=nlon, ny=nlat)
    w.bs.lon, w.bs.lat = synthetic.grid_coords(nlon, nlat)
    w.bs.u, w.bs.v = u, v
    w.bs.ready(xcyclic=True, keep_fields18=False, keep_diag=False)
    w.set_zwn(zwn)
    w.source_lon = np.asarray(slon_rad, dtype=np.float64)
    w.source_lat = np.asarray(slat_rad, dtype=np.float64)
    return w


_ORACLE = {}


def _config(name):
    """(u, v, slon_rad, slat_rad, zwn, days, freq) of a seeded sample of a BASELINE config (SURVEY.md 8-d sizes)."""
    rng = np.random.default_rng(2025)
    if name == "C2":        # 1 deg Clim, sources at every grid point |lat| <= 88, k = 1..6, stationary, 30 d
        u, v = synthetic.clim(360, 181)
        sl, sa = synthetic.all_gridpoint_sources(360, 181, 88.0)
        idx = np.sort(rng.choice(len(sl), 500, replace=False))
        return u, v, np.radians(sl[idx]), np.radians(sa[idx]), np.arange(1.0, 7.0), 30.0, 0.0
    if name == "C5":        # 0.25 deg Clim, sources at every grid point |lat| <= 48.125, k = 1..6, 60 d
        u, v = synthetic.clim(1440, 721)
        sl, sa = synthetic.all_gridpoint_sources(1440, 721, 48.125)
        idx = np.sort(rng.choice(len(sl), 500, replace=False))
        return u, v, np.radians(sl[idx]), np.radians(sa[idx]), np.arange(1.0, 7.0), 60.0, 0.0
    if name.startswith("C3"):   # frequency sweep: periods inf / 50 / 30 / 20 d x k = 1..8 from the 100 x 100 source matrix
        period = {"C3-inf": np.inf, "C3-50d": 50.0, "C3-30d": 30.0, "C3-20d": 20.0}[name]
        u, v = synthetic.clim(360, 181)
        ix, iy = np.meshgrid(np.arange(100), np.arange(100), indexing="xy")
        sl, sa = (0.0 + 3.6 * ix.ravel()) % 360.0, -49.5 + 1.0 * iy.ravel()
        idx = np.sort(rng.choice(len(sl), 125, replace=False))
        freq = 0.0 if np.isinf(period) else 2 * np.pi / (period * 86400.0)
        return u, v, np.radians(sl[idx]), np.radians(sa[idx]), np.arange(1.0, 9.0), 30.0, freq
    raise KeyError(name)


def _oracle(name):
    if name not in _ORACLE:
        u, v, slon, slat, zwn, days, freq = _config(name)
        _ORACLE[name] = parity.oracle_nongrazing(u, v, slon, slat, zwn, days, freq)
    return _ORACLE[name]


@pytest.mark.parametrize("strict", BUILDS)
@pytest.mark.parametrize("name", ["C2", "C5", "C3-inf", "C3-50d", "C3-30d", "C3-20d"])
def test_baseline_config_sample_vs_oracle(name, strict):
    u, v, slon, slat, zwn, days, freq = _config(name)
    ref, acc, ng = _oracle(name)
    w = make_wr(u, v, zwn, slon, slat, days, freq)
    w.ray_run(mode="numpy", root_method="numpy", inte_method="rk45", strict=strict)
    got = {n: getattr(w, n) for n in parity.NAMES}
    label = f"{name} sample ({len(slon)} sources x k=1..{len(zwn)}, {days:.0f} d) vs oracle, {'strict' if strict else 'fast'}"
    st = parity.compare(got, w.accepted, ref, acc, ng, label, max_outliers=int(ng.sum()) // 2000)
    live = ~np.isnan(ref["rmwn"][0])
    st["status_counts_nongrazing"] = {int(c): int(((w.status == c) & ng & live).sum()) for c in (1, 2, 3, 4)}   # critical-line deaths (4) occur here
    st["died_nongrazing"] = int((np.isnan(ref["rlon"][-1]) & ng & live).sum())
    st["died_all_live"] = int((np.isnan(ref["rlon"][-1]) & live).sum())
    st["died_mask_agreement_all_live"] = float(((np.isnan(ref["rlon"][-1]) == np.isnan(w.rlon[-1])))[live].mean())
    parity.record(label, st)
    # enough rays carry the assertion (the non-grazing share shrinks with the duration: 57 % of the live rays at 30 d on
    # the 1 deg background, 23 % at 60 d on the 0.25 deg one -- the reference's own sensitivity)
    assert st["nongrazing"] >= min(0.4 * st["live"], 900), st
    assert abs(st["steps_total"] - st["steps_total_ref"]) < 1e-2 * st["steps_total_ref"]
    # t = 0: slot layout identical on every (source, k) cell of the sample, values to 1e-13
    for n in ("rmwn", "rug", "rvg"):
        assert np.array_equal(np.isnan(got[n][0]), np.isnan(ref[n][0])), n
        np.testing.assert_allclose(got[n][0], ref[n][0], rtol=1e-13, atol=0, equal_nan=True)


@pytest.mark.parametrize("strict", BUILDS)
def test_c4_ensemble_members_vs_oracle(strict):
    """BASELINE configs[3]: members of the 1000-member ensemble (seeds 20250327 + e, 10 % jitter), 56 sources x k = 1..6
    each, batched in ONE handle (nmember = 3, per-source member index) -- each member against its own oracle run."""
    import torch
    members = (0, 500, 999)
    nlon, nlat = 360, 181
    bg = [synthetic.ensemble_member(e, nlon, nlat) for e in members]
    lon, lat = synthetic.grid_coords(nlon, nlat)
    rng = np.random.default_rng(4)
    sl, sa = rng.uniform(0.0, 360.0, 56), rng.uniform(-70.0, 70.0, 56)
    slon, slat = np.radians(sl), np.radians(sa)
    zwn = np.arange(1.0, 7.0)
    days = 30.0
    nt = int(days * 12) + 1
    ns, nz = 56 * len(members), len(zwn)
    member_of = np.repeat(np.arange(len(members), dtype=np.int32), 56)
    L = lib.load()
    h = lib.vp()
    lib.check(L.rwrt_create(0, nlon, nlat, 1, len(members), C.byref(h)))
    try:
        dev = torch.device("cuda", 0)
        d = {k: torch.from_numpy(np.ascontiguousarray(x)).to(dev) for k, x in
             dict(u=np.stack([b[0] for b in bg]), v=np.stack([b[1] for b in bg]), lon=lon, lat=lat, cl=np.cos(lat), sl=np.sin(lat),
                  slon=np.tile(slon, len(members)), slat=np.tile(slat, len(members)), zwn=zwn, mem=member_of).items()}
        sp = lib.stream_ptr()
        lib.check(L.rwrt_bs_ready(h, lib.ptr(d["u"]), lib.ptr(d["v"]), lib.ptr(d["lon"]), lib.ptr(d["lat"]), lib.ptr(d["cl"]),
                                  lib.ptr(d["sl"]), None, None, sp))
        N = 3 * ns * nz
        out = torch.empty((7, nt, N), dtype=torch.float64, device=dev)
        accd = torch.zeros(N, dtype=torch.int64, device=dev)
        lib.check(L.rwrt_ray_init(h, ns, lib.ptr(d["slon"]), lib.ptr(d["slat"]), lib.ptr(d["mem"]), nz, lib.ptr(d["zwn"]), 0.0, nt,
                                  lib.ptr(out), sp))
        lib.check(L.rwrt_ray_run_rk45(h, ns, lib.ptr(d["mem"]), nz, nt, 7200.0, days * 86400.0, 1e-6, 1e-6, 1e-3, 0.1,
                                      lib.MODE_STRICT if strict else lib.MODE_FAST, lib.ptr(out), lib.ptr(accd), None, None, None, sp))
        torch.cuda.synchronize()
        got_all = out.cpu().numpy().reshape(7, nt, 3, ns, nz)
        acc_all = accd.cpu().numpy().reshape(3, ns, nz)
    finally:
        L.rwrt_destroy(h)
    for j, e in enumerate(members):
        key = f"C4-{e}"
        if key not in _ORACLE:
            _ORACLE[key] = parity.oracle_nongrazing(bg[j][0], bg[j][1], slon, slat, zwn, days)
        ref, acc, ng = _ORACLE[key]
        sel = slice(56 * j, 56 * (j + 1))
        got = {n: got_all[k][:, :, sel, :] for k, n in enumerate(parity.NAMES)}
        label = f"C4 member {e} (56 sources x k=1..6, 30 d, batched) vs oracle, {'strict' if strict else 'fast'}"
        st = parity.compare(got, acc_all[:, sel, :], ref, acc, ng, label, max_outliers=0)
        parity.record(label, st)
        assert st["nongrazing"] > 0.5 * st["live"], st


# ---------------------------------------------------------------------------------------------
# termination
# ---------------------------------------------------------------------------------------------
def _violent():
    u, v = synthetic.clim(144, 73)
    u = (u * 6.0).astype(np.float32)     # exaggerated flow: the speed cut-off and pole crossings do happen
    sl, sa = synthetic.random_sources(400, 41, max_abs_lat_deg=88.0)
    return u, v, np.radians(sl), np.radians(sa), np.array([1.0, 2.0, 3.0]), 10.0


@pytest.mark.parametrize("strict", BUILDS)
def test_termination_exact_on_nongrazing_set(strict):
    """wr.py:838-850 + cal_dis (wr.py:97-112): which ray dies, at which output row and why -- identical to the oracle on
    every non-grazing ray, in the strict build (haversine as written) and in the fast build (threshold form
    sin^2(d/2) >= sin^2(cut/2) of the same test)."""
    u, v, slon, slat, zwn, days = _violent()
    if "violent" not in _ORACLE:
        _ORACLE["violent"] = parity.oracle_nongrazing(u, v, slon, slat, zwn, days)
    ref, acc, ng = _ORACLE["violent"]
    w = make_wr(u, v, zwn, slon, slat, days)
    w.ray_run(mode="numpy", inte_method="rk45", strict=strict)
    got = {n: getattr(w, n) for n in parity.NAMES}
    label = f"termination case (6 x Clim winds, 400 sources, 10 d) vs oracle, {'strict' if strict else 'fast'}"
    st = parity.compare(got, w.accepted, ref, acc, ng, label, max_outliers=0)   # includes: NaN patterns identical row by row
    live = ~np.isnan(ref["rmwn"][0])
    sel = ng & live
    died_ref = np.isnan(ref["rlon"][-1]) & live
    died_gpu = np.isnan(w.rlon[-1]) & live
    assert np.array_equal(died_ref[sel], died_gpu[sel])
    first_nan_ref = np.isnan(ref["rlon"]).argmax(axis=0)
    first_nan_gpu = np.isnan(w.rlon).argmax(axis=0)
    assert np.array_equal(first_nan_ref[sel & died_ref], first_nan_gpu[sel & died_ref])       # same row of death
    # the test is only worth something if every kind of death occurs among the non-grazing rays
    kinds = {code: int(((w.status == code) & sel).sum()) for code in (lib.RAY_ALIVE, lib.RAY_POLE, lib.RAY_SPEED, lib.RAY_NAN)}
    st.update(died_nongrazing=int((died_ref & sel).sum()), status_counts_nongrazing=kinds,
              died_all_live=int(died_ref.sum()), died_mask_agreement_all_live=float((died_ref == died_gpu)[live].mean()))
    parity.record(label, st)
    assert kinds[lib.RAY_SPEED] >= 5 and kinds[lib.RAY_NAN] + kinds[lib.RAY_POLE] >= 3, kinds
    assert ((w.status == lib.RAY_ALIVE) == (~np.isnan(w.rlon[-1]) & live)).all()
