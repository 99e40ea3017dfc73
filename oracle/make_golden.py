"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, read-only) in the build container through oracle/refshim.py.

    python -m oracle.make_golden            # from the repo root; ~2 minutes

The reference ships no golden vectors, and it cannot travel to the GPU box,
so these fixtures are how both the C oracle (CPU tests) and the CUDA path (GPU tests) are pinned
to the reference's own outputs.  Inputs are regenerated from `synthetic.py` (frozen, seeded); each
fixture stores the sha256 of the float32 winds it was made from so that drift is detected.

"Non-grazing" rays (BASELINE.json north_star; SURVEY.md section 7 hard part 1): the reference is
run four more times with every source longitude / latitude moved by +-1 ulp; a ray is non-grazing
when all five reference runs agree to < 1e-7 degrees at every output time, have identical NaN
patterns and identical accepted-step counts (one perturbation alone misses ~0.2 % of the sensitive
rays; four missed none of 6 184 in the experiment recorded in DESIGN.md).  Parity (< 1e-6 degrees, equal step counts) is demanded on that
set only; the complement is reported, not asserted.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rwrt_b200  # noqa: E402  (alias module for the hyphenated package directory)
from rwrt_b200 import synthetic  # noqa: E402
from oracle import refshim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
NAMES = ("rlon", "rlat", "rzwn", "rmwn", "ramp", "rug", "rvg")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def nongrazing_mask(ref, per):
    """(3, nsource, nzwn) bool from two reference runs (baseline, 1-ulp perturbed)."""
    ok = np.ones(ref["accepted"].shape, dtype=bool)
    for n in NAMES:
        ok &= np.all(np.isnan(ref[n]) == np.isnan(per[n]), axis=0)
    with np.errstate(invalid="ignore"):
        for n in ("rlon", "rlat"):
            d = np.abs(ref[n] - per[n])
            d = np.where(np.isnan(d), 0.0, d)
            ok &= np.degrees(d.max(axis=0)) < 1e-7
    ok &= ref["accepted"] == per["accepted"]
    return ok


def trajectory_fixture(name, background, bg_kwargs, zwn, src_lon, src_lat, days, freq=0.0, tstep_h=2.0, rk4=False):
    u, v = getattr(synthetic, background)(**bg_kwargs)
    run = refshim.run_rk4 if rk4 else refshim.run_rk45
    w = refshim.make_wr(u, v, zwn, src_lon, src_lat, tstep_h=tstep_h, ttotal_d=days, freq=freq)
    ref = run(w)
    # four perturbed runs of the reference: every source longitude / latitude moved by +-1 ulp (radians)
    ng = np.ones(ref["accepted"].shape, dtype=bool)
    for dlon_dir, dlat_dir in ((np.inf, None), (-np.inf, None), (None, np.inf), (None, -np.inf)):
        w2 = refshim.make_wr(u, v, zwn, src_lon, src_lat, tstep_h=tstep_h, ttotal_d=days, freq=freq)
        if dlon_dir is not None:
            w2.source_lon = np.nextafter(w.source_lon, dlon_dir)
        if dlat_dir is not None:
            w2.source_lat = np.nextafter(w.source_lat, dlat_dir)
        ng &= nongrazing_mask(ref, run(w2))
    live = ~np.isnan(ref["rmwn"][0])
    print(f"{name}: slots {live.size}, live {live.sum()}, non-grazing live {np.sum(ng & live)}, "
          f"accepted {ref['accepted'].sum()}")
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        background=background, bg_kwargs=repr(bg_kwargs), u_sha=sha(u), v_sha=sha(v),
        zwn=np.asarray(zwn, "float64"), src_lon_deg=np.asarray(src_lon, "float64"),
        src_lat_deg=np.asarray(src_lat, "float64"), src_lon_rad=w.source_lon, src_lat_rad=w.source_lat,
        days=days, freq=freq, tstep_h=tstep_h, rtol=1e-6, atol=1e-6, min_step_factor=1e-3,
        fields_sha=sha(w.bs.fields), accepted=ref["accepted"], nongrazing=ng,
        **{n: ref[n] for n in NAMES})


def fields_fixture():
    out = {}
    for tag, (fn, kw) in {"jet_g1": ("jet_z", {}), "gc_g1": ("great_circle", {}),
                          "clim_g1": ("clim", dict(nlon=144, nlat=73)),
                          "clim_g2": ("clim", dict(nlon=360, nlat=181))}.items():
        u, v = getattr(synthetic, fn)(**kw)
        w = refshim.make_wr(u, v, [1.0], [10.0], [10.0])
        bs = w.bs
        out[tag + "_u_sha"] = sha(u)
        out[tag + "_v_sha"] = sha(v)
        out[tag + "_fields_sha"] = sha(bs.fields)
        out[tag + "_fields_shape"] = np.array(bs.fields.shape)
        # a deterministic sample of entries so that a failing hash can be localised
        rng = np.random.default_rng(11)
        idx = rng.integers(0, bs.fields.size, 4096)
        out[tag + "_sample_idx"] = idx
        out[tag + "_sample_val"] = bs.fields.reshape(-1)[idx]
        for nm in ("q", "uxx", "uyy", "vxx", "vyy", "uxy", "vxy", "betam", "KS"):
            out[f"{tag}_{nm}_sha"] = sha(np.asarray(getattr(bs, nm), dtype="float64"))
        out[tag + "_coslat"] = np.cos(bs.lat)
        out[tag + "_sinlat"] = np.sin(bs.lat)
    np.savez_compressed(os.path.join(GOLD, "fields.npz"), **out)
    print("fields.npz written")


def pointwise_fixture():
    """Sampler / Mercator transform / RHS / cubic roots / slot order on random points (Clim, 2.5 deg)."""
    m = refshim.load()
    u, v = synthetic.clim(144, 73)
    w = refshim.make_wr(u, v, [1.0, 2.0, 3.0], [10.0], [10.0])
    rng = np.random.default_rng(5)
    n = 4000
    plon = rng.uniform(-12.0, 25.0, n)
    plat = rng.uniform(-1.6, 1.6, n)
    plat[:10] = [np.pi / 2, -np.pi / 2, 0, 1.5533, -1.5533, 1.56, -1.56, 1.5707, 0.3, 0.3]
    plon[:6] = [0, 2 * np.pi, -1e-17, 4 * np.pi, -2 * np.pi, 6.283185307179586]
    merc18 = np.asarray(w.bs.cal_bs_mercator_point(plon, plat, mode="numpy")).T.copy()      # all 18 rows (bs.py:884-887)
    merc = merc18[:, :12].copy()
    # RHS: y = (lon, lat, k, l, A) as (5, n, 1, 1) -> wr.py:492-556
    k = rng.integers(1, 9, n).astype("float64")
    l = rng.uniform(-60, 60, n)
    l[20:30] = [100.0, -100.0, 99.999, 150.0, np.nan, 0.0, 1e-3, -1e-3, 50.0, -50.0]
    amp = rng.uniform(0.2, 3.0, n)
    y = np.stack([plon, plat, k, l, amp])
    dy, _ = w.diffun_numpy(y.reshape(5, n, 1, 1).copy())
    dy = dy[:5].reshape(5, n).T.copy()
    # cubic roots in np.roots (numba -> zgeev) order, and slot-ordered l from cal_ky
    polys, roots = [], []
    for i in range(3000):
        kk = float(rng.integers(1, 9))
        U = rng.uniform(-20, 60)
        V = 0.0 if i % 11 == 0 else (rng.uniform(-3, 3) if i % 7 else rng.uniform(-30, 30))
        Qx = rng.uniform(-30, 30)
        Qy = rng.uniform(-50, 300)
        c = np.array([kk ** 3 * (U - Qy / kk ** 2), kk ** 2 * V + Qx, kk * U, V])
        deg = 3
        while deg > 0 and abs(c[deg]) == 0:
            deg -= 1
        p = np.zeros(4)
        p[: deg + 1] = c[: deg + 1][::-1]
        r = m["bs"].roots_numba(p[: deg + 1] + 0j)
        rr = np.full(3, np.nan + 0j)
        rr[: len(r)] = r
        polys.append(np.concatenate([[deg], p]))
        roots.append(rr)
    fu = rng.uniform(-20, 60, 3000)
    fv = np.where(rng.uniform(size=3000) < 0.1, 0.0, rng.uniform(-3, 3, 3000))
    fqx = rng.uniform(-30, 30, 3000)
    fqy = rng.uniform(-50, 300, 3000)
    ky = {}
    for kz, fr in ((1.0, 0.0), (3.0, 0.0), (6.0, 0.0), (2.0, 2 * np.pi / (30 * 86400.0)), (5.0, -2 * np.pi / (20 * 86400.0))):
        ml, _ = m["bs"].cal_ky(fu, fv, fqx, fqy, np.array([fr]), kz, mode="numpy", root_method="numpy")
        ky[f"ky_k{int(kz)}_{'s' if fr == 0 else ('p' if fr > 0 else 'n')}"] = np.asarray(ml)
        ky[f"kyarg_k{int(kz)}_{'s' if fr == 0 else ('p' if fr > 0 else 'n')}"] = np.array([kz, fr])
    np.savez_compressed(os.path.join(GOLD, "pointwise.npz"), u_sha=sha(u), v_sha=sha(v), plon=plon, plat=plat,
                        merc=merc, merc18=merc18, y=y.T.copy(), dy=dy, polys=np.array(polys), roots=np.array(roots),
                        fu=fu, fv=fv, fqx=fqx, fqy=fqy, **ky)
    print("pointwise.npz written")


def ky_frac_fixture():
    """cal_ky (bs.py:985-1040) with FRACTIONAL zonal wavenumbers: zwn**3 goes through libm pow there (bs.py:1006),
    so the leading coefficient is only reproduced by a correctly rounded cube (CUDA pow is 2 ulp)."""
    m = refshim.load()
    rng = np.random.default_rng(77)
    n = 2000
    fu = rng.uniform(-20, 60, n)
    fv = np.where(rng.uniform(size=n) < 0.1, 0.0, rng.uniform(-3, 3, n))
    fqx = rng.uniform(-30, 30, n)
    fqy = rng.uniform(-50, 300, n)
    out = dict(fu=fu, fv=fv, fqx=fqx, fqy=fqy)
    for j, (kz, fr) in enumerate(((1.5, 0.0), (2.7, 0.0), (0.8, 0.0), (3.3, 2 * np.pi / (30 * 86400.0)),
                                  (6.1, -2 * np.pi / (20 * 86400.0)), (1.0 / 3.0, 0.0))):
        ml, _ = m["bs"].cal_ky(fu, fv, fqx, fqy, np.array([fr]), np.float64(kz), mode="numpy", root_method="numpy")
        out[f"ky_{j}"] = np.asarray(ml)
        out[f"kyarg_{j}"] = np.array([kz, fr])
    np.savez_compressed(os.path.join(GOLD, "ky_frac.npz"), **out)
    print("ky_frac.npz written")


def kat_fixture():
    """The solver alone on the demo ODEs of rkf45.py:772-882, with the reference's own RK45 class."""
    m = refshim.load()
    RK45 = m["rkf45"].RK45
    out = {}

    def solve(fun, y0, t_end, rtol, atol, gmin):
        y = np.array(y0, dtype="float64").reshape(-1, 1)
        s = RK45(fun, 0.0, y, t_end, rtol=rtol, atol=atol, Global_Minstep=gmin)
        acc = 0
        while s.status == "running":
            t_before = s.t.copy()
            s.step()
            acc += int((s.t != t_before).sum())
        return s.y[:, 0].copy(), acc

    def lorenz(t, y):
        return np.array([10.0 * (y[1] - y[0]), y[0] * (28.0 - y[2]) - y[1], y[0] * y[1] - (8.0 / 3.0) * y[2]])

    def poly(t, y):
        return np.array([2.0 * y[1], 1.0 + 0 * y[1]])

    def expo(t, y):
        return np.array([np.exp(0.1 * y[1]), 1.0 + 0 * y[1]])

    for tag, kind, fun, y0, t_end, rtol, atol in (
            ("lorenz", 0, lorenz, [1.0, 1.0, 1.0], 1.0, 1e-10, 1e-12),
            ("lorenz_loose", 0, lorenz, [-8.0, 7.0, 27.0], 2.0, 1e-6, 1e-6),
            ("poly", 1, poly, [0.0, 0.0], 10.0, 1e-10, 1e-12),
            ("expo", 2, expo, [0.0, 0.0], 10.0, 1e-10, 1e-12)):
        y, acc = solve(fun, y0, t_end, rtol, atol, 1e-9)
        out[tag] = np.concatenate([[kind, t_end, rtol, atol, 1e-9, acc], y0, y])
        print(tag, y, acc)
    np.savez_compressed(os.path.join(GOLD, "kat.npz"), **out)


FIXTURES = {}     # name -> zero-argument maker (filled by _register below)


def check(names=None, verbose=True):
    """`python -m oracle.make_golden --check [names]`: regenerate fixtures from the live reference into a scratch
    directory and compare them with the committed files, array by array, bit for bit.  Returns the list of
    differing (file, key) pairs (empty = the committed fixtures ARE the reference's outputs)."""
    import tempfile
    global GOLD
    committed = GOLD
    bad = []
    with tempfile.TemporaryDirectory() as tmp:
        GOLD = tmp
        try:
            for name in (names or list(FIXTURES)):
                FIXTURES[name]()
                new = np.load(os.path.join(tmp, name + ".npz"), allow_pickle=False)
                old = np.load(os.path.join(committed, name + ".npz"), allow_pickle=False)
                if sorted(new.files) != sorted(old.files):
                    bad.append((name, "<key set>"))
                    continue
                for k in new.files:
                    a, b = new[k], old[k]
                    same = a.shape == b.shape and a.dtype == b.dtype and (
                        np.array_equal(a, b, equal_nan=True) if a.dtype.kind in "fc" else np.array_equal(a, b))
                    if not same:
                        bad.append((name, k))
                if verbose:
                    print(f"check {name}: {'OK' if not [x for x in bad if x[0] == name] else 'DIFFERS'} ({len(new.files)} arrays)")
        finally:
            GOLD = committed
    return bad


def _register():
    F = FIXTURES
    F["fields"] = fields_fixture
    F["pointwise"] = pointwise_fixture
    F["ky_frac"] = ky_frac_fixture
    F["kat"] = kat_fixture
    # C1: BASELINE.json configs[0] -- single source, k = 1..3, Jet-Z on 2.5 deg, 30 days
    F["c1_jet"] = lambda: trajectory_fixture("c1_jet", "jet_z", {}, [1.0, 2.0, 3.0], [120.0], [30.0], 30.0)
    # great-circle known-answer case (manual section 2): solid-body rotation, k = 1..5
    F["great_circle"] = lambda: trajectory_fixture("great_circle", "great_circle", {}, [1.0, 2.0, 3.0, 4.0, 5.0], [0.0] * 5,
                                                   [0.0, 5.0, 10.0, 15.0, 20.0], 30.0)

    def clim_g1():   # Clim on 2.5 deg: seeded random sources, k = 1..4, 12 days (cubic roots, zonally varying flow)
        sl, sa = synthetic.random_sources(16, 7)
        trajectory_fixture("clim_g1", "clim", dict(nlon=144, nlat=73), [1.0, 2.0, 3.0, 4.0], sl, sa, 12.0)

    def clim_g1_freq():   # non-stationary waves: period 30 days (omega enters the t=0 cubic only, bs.py:1005)
        sl, sa = synthetic.random_sources(10, 19, max_abs_lat_deg=60.0)
        trajectory_fixture("clim_g1_freq", "clim", dict(nlon=144, nlat=73), [1.0, 3.0, 5.0], sl, sa, 8.0,
                           freq=2 * np.pi / (30 * 86400.0))

    def clim_g2():   # 1 degree grid, short
        sl, sa = synthetic.random_sources(10, 23)
        trajectory_fixture("clim_g2", "clim", dict(nlon=360, nlat=181), [2.0, 4.0, 6.0], sl, sa, 6.0)

    def rk4_clim_g1():
        sl, sa = synthetic.random_sources(14, 31)
        trajectory_fixture("rk4_clim_g1", "clim", dict(nlon=144, nlat=73), [1.0, 2.0, 4.0], sl, sa, 10.0, rk4=True)

    F["clim_g1"], F["clim_g1_freq"], F["clim_g2"] = clim_g1, clim_g1_freq, clim_g2
    # the fixed-step RK4 path (inte_method='', main_wr.py's default): wr.py:702-765
    F["rk4_jet"] = lambda: trajectory_fixture("rk4_jet", "jet_z", {}, [1.0, 2.0, 3.0], [120.0, 60.0], [30.0, 20.0], 20.0, rk4=True)
    F["rk4_clim_g1"] = rk4_clim_g1


_register()


def main():
    os.makedirs(GOLD, exist_ok=True)
    if not refshim.available():
        raise SystemExit("reference not mounted at /root/reference")
    if "--check" in sys.argv:
        names = [a for a in sys.argv[sys.argv.index("--check") + 1:] if not a.startswith("-")]
        bad = check(names or None)
        if bad:
            raise SystemExit(f"committed fixtures differ from the live reference: {bad}")
        print("all checked fixtures are bit-identical to the live reference's outputs")
        return
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    for name, make in FIXTURES.items():
        if not only or name in only:
            make()


if __name__ == "__main__":
    main()
