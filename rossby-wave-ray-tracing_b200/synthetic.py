"""Frozen synthetic background states and source sets.

All winds are generated in float64 and cast to float32, exactly as the reference
stores what it reads from a NetCDF file (`bs.py:80-81, 210-211`, read_dtype='float32').
Arrays are indexed [i_lon, j_lat] like the reference's `BS.u`, `BS.v` (`bs.py:245-247`).
Coordinates use the reference's "lon/lat missing" branch (`bs.py:240-243`):
    lat = -pi/2 + j*pi/(nlat-1),  lon = i*2*pi/nlon      (float64 radians)

Nothing here touches the GPU; it only manufactures inputs for tests and bench.py.
"""
from __future__ import annotations

import numpy as np

from .constants import pi

GRIDS = {
    "G1": (144, 73),      # 2.5 degree
    "G2": (360, 181),     # 1 degree
    "G3": (1440, 721),    # 0.25 degree
}


def grid_coords(nlon: int, nlat: int):
    """f64 lon, lat in radians as the reference builds them (`bs.py:77-78, 240-243`)."""
    dx = np.array([2.0 * pi / nlon], dtype="float64")
    dy = np.array([pi / (nlat - 1)], dtype="float64")
    lat = -pi * 0.5 + np.arange(nlat) * dy
    lon = np.arange(nlon) * dx
    return lon, lat


def _mesh(nlon, nlat):
    lon, lat = grid_coords(nlon, nlat)
    return lon[:, None], lat[None, :]


def great_circle(nlon=144, nlat=73):
    """Solid-body super-rotation u = 15 cos(phi), v = 0 (manual's great-circle example)."""
    lam, phi = _mesh(nlon, nlat)
    u = 15.0 * np.cos(phi) + 0.0 * lam
    v = np.zeros_like(u)
    return u.astype("float32"), v.astype("float32")


def jet_z(nlon=144, nlat=73):
    """Jet-Z (config C1): u = 15 cos(phi) + 30 exp(-((phi-35deg)/10deg)^2), v = 0."""
    lam, phi = _mesh(nlon, nlat)
    d = pi / 180.0
    u = 15.0 * np.cos(phi) + 30.0 * np.exp(-((phi - 35.0 * d) / (10.0 * d)) ** 2) + 0.0 * lam
    v = np.zeros_like(u)
    return u.astype("float32"), v.astype("float32")


def clim(nlon=360, nlat=181, seed=20250327, jitter=0.0):
    """'Clim' climatological-like background (configs C2, C3, C5; C4 with jitter).

    u = 12 cos(phi) + 28 e^{-((phi-35)/12)^2} + 24 e^{-((phi+40)/12)^2}
        + sum_{m=1..3} a_m cos(m lam + alpha_m) e^{-((phi-45)/15)^2}
    v = sum_{m=1..3} b_m sin(m lam + beta_m) cos^2(phi) sin(2 phi)
    a_m ~ U(2,6), b_m ~ U(0.5,2), phases ~ U(0,2pi) from default_rng(seed).
    `jitter` (ensemble members): multiplies each of the three jet/solid-body amplitudes
    by (1 + jitter * N(0,1)) drawn from the same generator *after* the waves.
    """
    rng = np.random.default_rng(seed)
    a = rng.uniform(2.0, 6.0, 3)
    b = rng.uniform(0.5, 2.0, 3)
    alpha = rng.uniform(0.0, 2.0 * pi, 3)
    beta = rng.uniform(0.0, 2.0 * pi, 3)
    amp = 1.0 + jitter * rng.standard_normal(3) if jitter else np.ones(3)
    lam, phi = _mesh(nlon, nlat)
    d = pi / 180.0
    u = (12.0 * amp[0] * np.cos(phi)
         + 28.0 * amp[1] * np.exp(-((phi - 35.0 * d) / (12.0 * d)) ** 2)
         + 24.0 * amp[2] * np.exp(-((phi + 40.0 * d) / (12.0 * d)) ** 2)) + 0.0 * lam
    v = np.zeros_like(u)
    env = np.exp(-((phi - 45.0 * d) / (15.0 * d)) ** 2)
    vshape = np.cos(phi) ** 2 * np.sin(2.0 * phi)
    for m in range(3):
        u = u + a[m] * np.cos((m + 1) * lam + alpha[m]) * env
        v = v + b[m] * np.sin((m + 1) * lam + beta[m]) * vshape
    return u.astype("float32"), v.astype("float32")


def ensemble_member(e: int, nlon=360, nlat=181):
    """Config C4 member e: Clim with seed 20250327+e and 10 % amplitude jitter."""
    return clim(nlon, nlat, seed=20250327 + e, jitter=0.10)


def all_gridpoint_sources(nlon: int, nlat: int, max_abs_lat_deg=88.0):
    """Sources at every grid point with |lat| <= max_abs_lat_deg (config C2 / C5), degrees."""
    lon_deg = np.arange(nlon) * (360.0 / nlon)
    lat_deg = -90.0 + np.arange(nlat) * (180.0 / (nlat - 1))
    lat_deg = lat_deg[np.abs(lat_deg) <= max_abs_lat_deg + 1e-9]
    LON, LAT = np.meshgrid(lon_deg, lat_deg, indexing="xy")   # lon fastest, like set_source_matrix
    return LON.ravel().copy(), LAT.ravel().copy()


def random_sources(n: int, seed: int, max_abs_lat_deg=75.0):
    rng = np.random.default_rng(seed)
    lon = rng.uniform(0.0, 360.0, n)
    lat = rng.uniform(-max_abs_lat_deg, max_abs_lat_deg, n)
    return lon, lat
