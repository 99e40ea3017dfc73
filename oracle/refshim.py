"""TEST INFRASTRUCTURE — runs the UNMODIFIED reference (read-only, /root/reference) in the
build container to (a) pin the C oracle in `oracle/rwrt_oracle.c` and (b) generate the golden
fixtures under `tests/golden/` (see `oracle/make_golden.py`).

/root/reference does not exist on the GPU box, so nothing in `-m gpu` tests, smoke() or
bench.py imports this module; it is used by `make_golden.py` and by the optional
`tests/test_oracle_vs_reference.py` (skipped when the reference is absent).

Recipe = SURVEY.md Appendix E:
  * stub packages `netCDF4` and `matplotlib.pyplot` (imported at bs.py:11-12, wr.py:15,
    rkf45.py:2; neither is installed here and neither is used on the RK45 path),
  * NUMBA_CACHE_DIR redirected (the reference tree is read-only; it uses cache=True),
  * background arrays injected directly into `BS` (bypassing `loadbs_ncfile`, bs.py:202-262),
  * per-ray accepted-step counter = wrapper around `rkf45.RungeKutta._step_impl`
    (rkf45.py:375): a ray made an accepted step in a pass iff its t changed and its state
    was not NaN on entry.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

REFERENCE_DIR = os.environ.get("RWRT_REFERENCE_DIR", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")
_mods = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "wr.py"))


def load():
    """Import the reference modules (once). Returns a dict of modules."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_DIR}")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/rwrt_numba_cache")
    os.makedirs(os.environ["NUMBA_CACHE_DIR"], exist_ok=True)
    for p in (REFERENCE_DIR, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the reference's top-level module names are generic (bs, wr, wn, constants ...): import them
    # under their own names but keep handles here so callers never rely on sys.path order again
    import importlib
    mods = {}
    for name in ("constants", "interpolation", "bs", "wn", "rkf45", "wr"):
        mods[name] = importlib.import_module(name)
    mods["wr"].progress_bar = lambda *a, **k: None
    _mods = mods
    return mods


class StepCounter:
    """Counts accepted steps per ray by wrapping RungeKutta._step_impl (rkf45.py:375-514)."""

    def __init__(self, rk_module):
        self.rk = rk_module
        self.accepted = None
        self.passes = 0
        self._orig = None

    def __enter__(self):
        orig = self.rk.RungeKutta._step_impl
        self._orig = orig
        counter = self

        def wrapped(solver):
            t_before = np.array(solver.t, copy=True)
            nan_before = np.isnan(np.mean(solver.y, axis=0))
            out = orig(solver)
            moved = (solver.t != t_before) & ~nan_before
            if counter.accepted is None:
                counter.accepted = np.zeros(moved.shape, dtype=np.int64)
            counter.accepted += moved
            counter.passes += 1
            return out

        self.rk.RungeKutta._step_impl = wrapped
        return self

    def __exit__(self, *exc):
        self.rk.RungeKutta._step_impl = self._orig
        return False


def make_wr(u, v, zwn, src_lon_deg, src_lat_deg, *, tstep_h=2.0, ttotal_d=30.0, freq=0.0,
            rtol=1e-6, atol=1e-6, min_step_factor=1e-3, xcyclic=True, lon=None, lat=None):
    """Build a reference WR with injected background (u, v: (nlon, nlat) float32)."""
    m = load()
    nlon, nlat = u.shape
    zwn = np.asarray(zwn, dtype="float64")
    w = m["wr"].WR(len(zwn), len(src_lon_deg), tstep_h * 3600.0, ttotal_d * 86400.0, freq,
                   nx=nlon, ny=nlat, rtol=rtol, atol=atol, MinStepFactor=min_step_factor)
    w.bs.lat = (-m["constants"].pi * 0.5 + np.arange(nlat) * w.bs.dy) if lat is None else np.asarray(lat, "float64")
    w.bs.lon = (np.arange(nlon) * w.bs.dx) if lon is None else np.asarray(lon, "float64")
    w.bs.u = np.asarray(u, dtype="float32")
    w.bs.v = np.asarray(v, dtype="float32")
    w.bs.ready(xcyclic=xcyclic)
    w.set_zwn(zwn)
    w.set_source_array(np.asarray(src_lon_deg, "float64"), np.asarray(src_lat_deg, "float64"))
    return w


def run_rk45(w, quiet=True):
    """`WR.ray_run(mode='numpy', root_method='numpy', inte_method='rk45')` with step counting.

    Returns dict with the 7 result arrays (nt,3,nsource,nzwn) and `accepted` (3,nsource,nzwn) int64.
    """
    m = load()
    with StepCounter(m["rkf45"]) as sc:
        ctx = contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()
        with ctx:
            w.ray_run(mode="numpy", root_method="numpy", inte_method="rk45")
    acc = sc.accepted
    if acc is None:
        acc = np.zeros(3 * w.nsource * w.nzwn, dtype=np.int64)
    out = {k: getattr(w, k).copy() for k in ("rlon", "rlat", "rzwn", "rmwn", "ramp", "rug", "rvg")}
    out["accepted"] = acc.reshape(3, w.nsource, w.nzwn)
    out["passes"] = sc.passes
    return out


def run_rk4(w, quiet=True):
    """`WR.ray_run(mode='numpy', root_method='numpy', inte_method='')`: the fixed-step RK4 path (wr.py:702-765).
    `accepted` is the number of output intervals (every slot takes nt-1 fixed steps)."""
    ctx = contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()
    with ctx:
        w.ray_run(mode="numpy", root_method="numpy", inte_method="")
    out = {k: getattr(w, k).copy() for k in ("rlon", "rlat", "rzwn", "rmwn", "ramp", "rug", "rvg")}
    out["accepted"] = np.full((3, w.nsource, w.nzwn), w.nt - 1, dtype=np.int64)
    out["passes"] = w.nt - 1
    return out
