"""TEST INFRASTRUCTURE — ctypes front end of the C restatement in oracle/rwrt_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It is the checker, never the product: nothing under
rossby-wave-ray-tracing_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_oracle

_lib = None

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int64)


def _p(a, t=_dp):
    return None if a is None else a.ctypes.data_as(t)


def lib():
    global _lib
    if _lib is None:
        path = build_oracle.build()
        L = C.CDLL(path)
        L.orc_bs_ready.restype = C.c_int
        L.orc_bs_ready.argtypes = [C.c_int, C.c_int, _fp, _fp, _dp, _dp, _dp, C.c_int, _dp, _dp]
        L.orc_rhs.restype = None
        L.orc_rhs.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _dp, _dp]
        L.orc_sample_mercator.restype = None
        L.orc_sample_mercator.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _dp, _dp, _dp]
        L.orc_ray_run_rk45.restype = C.c_int
        L.orc_ray_run_rk45.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_int64, C.c_int, C.c_double, C.c_double,
                                       C.c_double, C.c_double, C.c_double, C.c_double, _dp, _ip, _ip]
        L.orc_ray_run_rk4.restype = C.c_int
        L.orc_ray_run_rk4.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_int64, C.c_int, C.c_double, C.c_double, _dp]
        L.orc_kat.restype = C.c_int64
        L.orc_kat.argtypes = [C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_double, C.c_double, _dp]
        L.orc_ray_init.restype = C.c_int
        L.orc_ray_init.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _dp, _dp, C.c_int, _dp, C.c_double,
                                   C.c_int, _dp]
        L.orc_roots.restype = C.c_int
        L.orc_roots.argtypes = [_dp, C.c_int, _dp, _dp]
        L.orc_cal_ky.restype = None
        L.orc_cal_ky.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def bs_ready(u, v, lat, xcyclic=True, coslat=None, sinlat=None, want_diag=False):
    """bs.py:318-407.  u, v (nlon, nlat) float32 -> fields (nlon+xcyclic, nlat, 18) [, diag (9, nlon, nlat)]."""
    u = np.ascontiguousarray(u, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    nlon, nlat = u.shape
    lat = _f64(lat)
    fields = np.empty((nlon + (1 if xcyclic else 0), nlat, 18), dtype=np.float64)
    diag = np.empty((9, nlon, nlat), dtype=np.float64) if want_diag else None
    cl = None if coslat is None else _f64(coslat)
    sl = None if sinlat is None else _f64(sinlat)
    rc = lib().orc_bs_ready(nlon, nlat, _p(u, _fp), _p(v, _fp), _p(lat), _p(cl), _p(sl), int(bool(xcyclic)),
                            _p(fields), _p(diag))
    if rc != 0:
        raise MemoryError("orc_bs_ready")
    return (fields, diag) if want_diag else fields


def rhs(fields, lon, lat, y):
    """One RHS evaluation per ray; y (n, 5) -> (n, 5)."""
    fields = _f64(fields); lon = _f64(lon); lat = _f64(lat); y = _f64(y)
    dy = np.empty_like(y)
    lib().orc_rhs(fields.shape[0], fields.shape[1], _p(fields), _p(lon), _p(lat), y.shape[0], _p(y), _p(dy))
    return dy


def sample_mercator(fields, lon, lat, plon, plat):
    """bs.py:781-887: 12 Mercator quantities (fmu..fmqyy) at n points -> (n, 12)."""
    fields = _f64(fields); lon = _f64(lon); lat = _f64(lat); plon = _f64(plon); plat = _f64(plat)
    out = np.empty((plon.shape[0], 12), dtype=np.float64)
    lib().orc_sample_mercator(fields.shape[0], fields.shape[1], _p(fields), _p(lon), _p(lat), plon.shape[0],
                              _p(plon), _p(plat), _p(out))
    return out


def ray_init(fields, lon, lat, src_lon, src_lat, zwn, freq, nt):
    """wr.py:344-395: returns out (7, nt, 3, nsource, nzwn) with row 0 filled, the rest NaN."""
    fields = _f64(fields); lon = _f64(lon); lat = _f64(lat)
    src_lon = _f64(src_lon); src_lat = _f64(src_lat); zwn = _f64(zwn)
    ns, nz = src_lon.shape[0], zwn.shape[0]
    out = np.full((7, nt, 3, ns, nz), np.nan, dtype=np.float64)
    lib().orc_ray_init(fields.shape[0], fields.shape[1], _p(fields), _p(lon), _p(lat), ns, _p(src_lon),
                       _p(src_lat), nz, _p(zwn), float(freq), nt, _p(out))
    return out


def ray_run_rk45(fields, lon, lat, out, tstep, ttotal, rtol=1e-6, atol=1e-6, min_step_factor=1e-3, cut_off=0.1):
    """wr.py:767-887 in place on `out` (7, nt, 3, nsource, nzwn) whose row 0 holds the initial state.

    Returns (accepted, attempts), each (3, nsource, nzwn) int64.
    """
    fields = _f64(fields); lon = _f64(lon); lat = _f64(lat)
    assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape[0] == 7
    nt = out.shape[1]
    shp = out.shape[2:]
    N = int(np.prod(shp))
    acc = np.zeros(N, dtype=np.int64)
    att = np.zeros(N, dtype=np.int64)
    lib().orc_ray_run_rk45(fields.shape[0], fields.shape[1], _p(fields), _p(lon), _p(lat), N, nt, float(tstep),
                           float(ttotal), float(rtol), float(atol), float(min_step_factor), float(cut_off),
                           _p(out), _p(acc, _ip), _p(att, _ip))
    return acc.reshape(shp), att.reshape(shp)


def ray_run_rk4(fields, lon, lat, out, tstep, cut_off=0.1):
    """wr.py:702-765 (fixed-step RK4, the reference's default inte_method='') in place on `out`."""
    fields = _f64(fields); lon = _f64(lon); lat = _f64(lat)
    assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape[0] == 7
    nt = out.shape[1]
    N = int(np.prod(out.shape[2:]))
    lib().orc_ray_run_rk4(fields.shape[0], fields.shape[1], _p(fields), _p(lon), _p(lat), N, nt, float(tstep),
                          float(cut_off), _p(out))
    return out


def kat(kind, y0, t_end, rtol, atol, gmin):
    """Solver known-answer ODEs (rkf45.py:772-882): 0 Lorenz-63, 1 dx/dt=2t, 2 dx/dt=exp(0.1 t) (both as (x,t))."""
    y0 = _f64(y0)
    y = np.empty_like(y0)
    acc = lib().orc_kat(int(kind), y0.shape[0], _p(y0), float(t_end), float(rtol), float(atol), float(gmin), _p(y))
    return y, int(acc)


def roots(p):
    """np.roots order emulation for a real polynomial (descending coefficients, degree <= 3)."""
    p = _f64(p)
    deg = p.shape[0] - 1
    wr = np.zeros(3); wi = np.zeros(3)
    n = lib().orc_roots(_p(p), deg, _p(wr), _p(wi))
    return (wr + 1j * wi)[:n]


def cal_ky(fu, fv, fqx, fqy, freq, k):
    """bs.py:985-1040 -> (n, 3) slot-ordered meridional wavenumbers."""
    fu = _f64(fu); fv = _f64(fv); fqx = _f64(fqx); fqy = _f64(fqy)
    out = np.empty((fu.shape[0], 3), dtype=np.float64)
    lib().orc_cal_ky(fu.shape[0], _p(fu), _p(fv), _p(fqx), _p(fqy), float(freq), float(k), _p(out))
    return out
