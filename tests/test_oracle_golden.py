"""CPU tests: pin the C oracle (oracle/rwrt_oracle.c) to outputs of the UNMODIFIED reference.

The golden fixtures under tests/golden/ were produced by oracle/make_golden.py, which imports the
reference from /root/reference in the build container.  Nothing here needs the reference or a GPU.
"""
import hashlib

import numpy as np
import pytest

import parity
import rwrt_b200  # noqa: F401
from rwrt_b200 import synthetic
from oracle import oracle


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


FIELD_CASES = {"jet_g1": ("jet_z", {}), "gc_g1": ("great_circle", {}),
               "clim_g1": ("clim", dict(nlon=144, nlat=73)), "clim_g2": ("clim", dict(nlon=360, nlat=181))}


@pytest.mark.parametrize("tag", list(FIELD_CASES))
def test_bs_ready_bit_exact(tag):
    """bs.py:318-407: the 18 stacked fields and the 9 diagnostic planes are bit-identical."""
    g = parity.load("fields")
    fn, kw = FIELD_CASES[tag]
    u, v = getattr(synthetic, fn)(**kw)
    assert sha(u) == str(g[tag + "_u_sha"]) and sha(v) == str(g[tag + "_v_sha"])
    _, lat = synthetic.grid_coords(*u.shape)
    fields, diag = oracle.bs_ready(u, v, lat, True, g[tag + "_coslat"], g[tag + "_sinlat"], want_diag=True)
    assert tuple(g[tag + "_fields_shape"]) == fields.shape
    np.testing.assert_array_equal(fields.reshape(-1)[g[tag + "_sample_idx"]], g[tag + "_sample_val"])
    assert sha(fields) == str(g[tag + "_fields_sha"])
    for k, nm in enumerate(("q", "uxx", "uyy", "vxx", "vyy", "uxy", "vxy", "betam", "KS")):
        assert sha(diag[k]) == str(g[f"{tag}_{nm}_sha"]), nm


def test_bs_ready_no_wrap_column():
    u, v = synthetic.jet_z()
    _, lat = synthetic.grid_coords(*u.shape)
    a = oracle.bs_ready(u, v, lat, True)
    b = oracle.bs_ready(u, v, lat, False)
    assert b.shape == (144, 73, 18) and a.shape == (145, 73, 18)
    np.testing.assert_array_equal(a[:-1], b)
    np.testing.assert_array_equal(a[-1], b[0])


def _clim_g1_fields():
    u, v = synthetic.clim(144, 73)
    lon, lat = synthetic.grid_coords(144, 73)
    return oracle.bs_ready(u, v, lat, True), lon, lat


def test_sampler_and_mercator():
    """bs.py:781-887 + interpolation.py:77-135 incl. poles, the 0.0175 mask, lon wrap edge cases."""
    g = parity.load("pointwise")
    F, lon, lat = _clim_g1_fields()
    got = oracle.sample_mercator(F, lon, lat, g["plon"], g["plat"])
    ref = g["merc"]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    # libm flavours differ in the last ulp of sin/cos/tan (NumPy SIMD vs glibc): relative 1e-11
    ok = np.isfinite(ref)
    scale = np.maximum(np.abs(ref[ok]), 1e-6)
    assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-11
    assert np.mean(got[ok] == ref[ok]) > 0.99


def test_rhs():
    """wr.py:492-556 diffun_numpy incl. |lat| >= pi/2 and |l| >= 100 masks."""
    g = parity.load("pointwise")
    F, lon, lat = _clim_g1_fields()
    got = oracle.rhs(F, lon, lat, g["y"])
    ref = g["dy"]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = np.isfinite(ref)
    scale = np.maximum(np.abs(ref[ok]), 1e-12)
    assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-9


def test_roots_in_lapack_order():
    """np.roots through numba (bs.py:38-40): the real roots, in order, bit for bit."""
    g = parity.load("pointwise")
    polys, roots = g["polys"], g["roots"]
    n3 = 0
    for row, ref in zip(polys, roots):
        deg = int(row[0])
        p = row[1: deg + 2]
        got = oracle.roots(p)
        ref = ref[:deg]
        assert len(got) == deg
        rr = ref.real[np.abs(ref.imag) < 1e-8]
        gr = got.real[np.abs(got.imag) < 1e-8]
        assert np.array_equal(rr, gr), (p, ref, got)
        n3 += len(rr) == 3
        # complex pairs may come out conjugated (signed-zero path), values must still agree
        np.testing.assert_allclose(np.sort(got.real), np.sort(ref.real), rtol=1e-13, atol=0)
        np.testing.assert_allclose(np.sort(np.abs(got.imag)), np.sort(np.abs(ref.imag)), rtol=1e-13, atol=0)
    assert n3 > 200  # the order-dependent three-real-root case is well covered


def test_cal_ky_slot_order():
    """bs.py:985-1040 + 942-982: slot layout (incl. NaN slots and |l| > 100 removal) is identical."""
    g = parity.load("pointwise")
    for key in [k for k in g.files if k.startswith("ky_")]:
        kz, fr = g["kyarg_" + key[3:]]
        got = oracle.cal_ky(g["fu"], g["fv"], g["fqx"], g["fqy"], fr, kz)
        assert np.array_equal(got, g[key], equal_nan=True), key


def test_cal_ky_fractional_wavenumbers():
    """bs.py:1006: zwn**3 through libm pow -- fractional k; slot layout and values bit for bit."""
    g = parity.load("ky_frac")
    for key in [k for k in g.files if k.startswith("ky_")]:
        kz, fr = g["kyarg_" + key[3:]]
        got = oracle.cal_ky(g["fu"], g["fv"], g["fqx"], g["fqy"], fr, kz)
        assert np.array_equal(got, g[key], equal_nan=True), key
        assert np.isfinite(g[key]).sum() > 500


@pytest.mark.parametrize("tag", ["lorenz", "lorenz_loose", "poly", "expo"])
def test_solver_known_answers(tag):
    """rkf45.py:772-882 demo ODEs through the reference's RK45 class vs the oracle's controller."""
    row = parity.load("kat")[tag]
    kind, t_end, rtol, atol, gmin, acc_ref = row[:6]
    n = (len(row) - 6) // 2
    y0, y_ref = row[6:6 + n], row[6 + n:]
    y, acc = oracle.kat(int(kind), y0, t_end, rtol, atol, gmin)
    assert acc == int(acc_ref)
    np.testing.assert_allclose(y, y_ref, rtol=1e-12)
    if tag == "poly":      # x(10) = 100 exactly: error magnitudes quoted at rkf45.py:856-858
        assert abs(y[0] - 100.0) < 1e-10
    if tag == "expo":      # x(10) = 10 (e - 1)
        assert abs(y[0] - 10.0 * (np.e - 1.0)) < 1e-8


TRAJ = ["c1_jet", "great_circle", "clim_g1", "clim_g1_freq", "clim_g2"]


def run_oracle(g):
    u, v = parity.background(g)
    lon, lat = synthetic.grid_coords(*u.shape)
    F = oracle.bs_ready(u, v, lat, True)
    assert sha(F) == str(g["fields_sha"])
    nt = g["rlon"].shape[0]
    out = oracle.ray_init(F, lon, lat, g["src_lon_rad"], g["src_lat_rad"], g["zwn"], float(g["freq"]), nt)
    init = out[:, 0].copy()
    acc, att = oracle.ray_run_rk45(F, lon, lat, out, float(g["tstep_h"]) * 3600.0, float(g["days"]) * 86400.0,
                                   float(g["rtol"]), float(g["atol"]), float(g["min_step_factor"]))
    return out, init, acc, att


@pytest.mark.parametrize("name", TRAJ)
def test_initial_state_bit_exact(name):
    """wr.py:344-395: all seven t=0 planes, incl. the LAPACK-order-dependent slot layout."""
    g = parity.load(name)
    _, init, _, _ = run_oracle(g)
    for k, n in enumerate(parity.NAMES):
        assert np.array_equal(init[k], g[n][0], equal_nan=True), n


@pytest.mark.parametrize("name", TRAJ)
def test_trajectories_vs_reference(name):
    """wr.py:767-887: < 1e-6 deg and identical accepted-step counts on the non-grazing rays."""
    g = parity.load(name)
    out, _, acc, att = run_oracle(g)
    got = {n: out[k] for k, n in enumerate(parity.NAMES)}
    ref = {n: g[n] for n in parity.NAMES}
    st = parity.compare(got, acc, ref, g["accepted"], g["nongrazing"], name)
    assert st["nongrazing"] >= 0.4 * st["live"]
    assert (att >= acc).all()


def test_great_circle_known_answer():
    """Manual section 2: in solid-body rotation stationary rays follow great circles and
    K = sqrt(k^2 + l^2) at the source is the same for every k (SURVEY.md probe6)."""
    g = parity.load("great_circle")
    out, init, acc, _ = run_oracle(g)
    l0 = init[3]                     # (3, ns, nz)
    k0 = init[2]
    assert np.isnan(l0[0]).all()     # v == 0: quadratic, slot 0 empty
    assert (l0[1][np.isfinite(l0[1])] < 0).all() and (l0[2][np.isfinite(l0[2])] > 0).all()
    K = np.sqrt(k0[2] ** 2 + l0[2] ** 2)
    for s in range(K.shape[0]):
        kk = K[s][np.isfinite(K[s])]
        assert np.ptp(kk) < 1e-9 * kk.mean()
    # great circle through the source with initial heading: all output points stay on one plane
    lonr, latr = out[0], out[1]
    for ir in (1, 2):
        for s in range(lonr.shape[2]):
            for iz in range(lonr.shape[3]):
                lo, la = lonr[:, ir, s, iz], latr[:, ir, s, iz]
                ok = np.isfinite(lo)
                if ok.sum() < 20:
                    continue
                p = np.stack([np.cos(la[ok]) * np.cos(lo[ok]), np.cos(la[ok]) * np.sin(lo[ok]), np.sin(la[ok])], 1)
                nrm = np.cross(p[0], p[5])
                nrm /= np.linalg.norm(nrm)
                assert np.max(np.abs(p @ nrm)) < 5e-3


@pytest.mark.parametrize("name", ["rk4_jet", "rk4_clim_g1"])
def test_rk4_path_vs_reference(name):
    """wr.py:702-765 (inte_method='', the reference's default): fixed-step RK4 with freeze-on-failure;
    NaN-root slots turn all-NaN at row 1 (they are not masked by wr.py:508-510)."""
    g = parity.load(name)
    u, v = parity.background(g)
    lon, lat = synthetic.grid_coords(*u.shape)
    F = oracle.bs_ready(u, v, lat, True)
    nt = g["rlon"].shape[0]
    out = oracle.ray_init(F, lon, lat, g["src_lon_rad"], g["src_lat_rad"], g["zwn"], float(g["freq"]), nt)
    oracle.ray_run_rk4(F, lon, lat, out, float(g["tstep_h"]) * 3600.0)
    got = {n: out[k] for k, n in enumerate(parity.NAMES)}
    ref = {n: g[n] for n in parity.NAMES}
    parity.compare(got, None, ref, g["accepted"], g["nongrazing"], name)
    dead = np.isnan(g["rmwn"][0])
    assert np.isnan(out[0][1:, dead]).all()            # NaN-root slots: everything NaN from row 1
    # with a fixed step there is no accept/reject chaos: most rays come out bit for bit
    a, b = out[0].reshape(nt, -1), g["rlon"].reshape(nt, -1)
    same = np.array([(a[:, i] == b[:, i])[np.isfinite(b[:, i])].all() for i in range(a.shape[1])])
    assert same.mean() > 0.6
