"""TEST INFRASTRUCTURE — NumPy definition of the ray-density / wave-ray-flux gridding (K4 checker).

PARITY UNPINNED: the reference repository ships no gridding code (its manual, section 4, describes
Fun1_threshold / Fun2_region_threshold / Fun3_WRF_calculate / WRF_universal, none of which is in the
tree), so there is nothing of the authors' to compare with.  This file freezes this project's own
definition; `rwrt_grid_accumulate` (csrc/rwrt_grid.cu) must reproduce it bit for bit:

  domain : three longitude circles [-360, 720) deg x [-90, 90] deg (manual section 4; rlon is never wrapped)
  cells  : ix = floor((lon_deg + 360) / dlon),  iy = min(floor((lat_deg + 90) / dlat), ncy - 1)
  sample : every output row it >= 1 of every slot whose lon, lat and l are finite, and 0 <= ix < ncx
  count  : += 1 (int64);  fx, fy : += rint(ug * 2^20), rint(vg * 2^20) (int64 fixed point) when ug, vg finite
"""
from __future__ import annotations

import numpy as np

RAD2DEG = 1.0 / (3.14159265358979323846264338327950288419716939937510 / 180.0)   # constants.py:14-15


def grid_shape(dlon_deg, dlat_deg):
    return int(np.ceil(1080.0 / dlon_deg)), int(np.ceil(180.0 / dlat_deg))


def accumulate(out7, dlon_deg, dlat_deg):
    """out7: (7, nt, ...) float64.  Returns count, fx, fy int64 arrays of shape (ncx, ncy)."""
    ncx, ncy = grid_shape(dlon_deg, dlat_deg)
    nt = out7.shape[1]
    o = out7.reshape(7, nt, -1)[:, 1:, :]
    lon, lat, l, ug, vg = o[0].ravel(), o[1].ravel(), o[3].ravel(), o[5].ravel(), o[6].ravel()
    ok = np.isfinite(lon) & np.isfinite(lat) & ~np.isnan(l)
    ok &= ~np.isnan(lon) & ~np.isnan(lat)
    with np.errstate(invalid="ignore"):
        fx_ = np.floor((lon * RAD2DEG + 360.0) / dlon_deg)
        fy_ = np.floor((lat * RAD2DEG + 90.0) / dlat_deg)
        ok &= (fx_ >= 0) & (fx_ < ncx) & (fy_ >= 0)
    ix = fx_[ok].astype(np.int64)
    iy = np.minimum(fy_[ok], ncy - 1).astype(np.int64)
    cell = ix * ncy + iy
    count = np.bincount(cell, minlength=ncx * ncy).astype(np.int64)
    fin = np.isfinite(ug[ok]) | np.isinf(ug[ok])
    fin = ~np.isnan(ug[ok]) & ~np.isnan(vg[ok])
    qx = np.rint(ug[ok][fin] * 1048576.0).astype(np.int64)
    qy = np.rint(vg[ok][fin] * 1048576.0).astype(np.int64)
    fx = np.zeros(ncx * ncy, dtype=np.int64)
    fy = np.zeros(ncx * ncy, dtype=np.int64)
    np.add.at(fx, cell[fin], qx)
    np.add.at(fy, cell[fin], qy)
    return count.reshape(ncx, ncy), fx.reshape(ncx, ncy), fy.reshape(ncx, ncy)
