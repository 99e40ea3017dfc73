"""`WN` — grid-wide wavenumber maps with the reference's interface (`wn.py:19-136`), GPU backed.

For every grid point and every initial zonal wavenumber: the propagating meridional wavenumbers `mwn`
(nlon, nlat, nzwn, 3), their number `rootnum` (nlon, nlat, nzwn) and the group velocities `ug`, `vg`
(nlon, nlat, nzwn, 3).  One `rwrt_ray_init` call with the grid points as sources does all of it (the t = 0
state of a ray started at a grid point IS that point's dispersion solution: `wr.py:344-395`).

Parity unpinned: the reference's `WN.cal_wave` never computes anything (`if result is None:` is inverted,
`wn.py:71-72`, and two helpers lack `self`, `wn.py:93, 118`), and the driver does not import the class
(out of scope).  What is implemented is what the method evidently intends, with the ray
path's own root solve and slot order (`cal_ky_numpy`, `bs.py:985-1040`) and, as in `wn.py:84-87`, zero
group velocity where a slot has no real root.  `ug`, `vg` use the t = 0 form `cal_ugvg_numpy` (`wn.py:209-259`).
"""
from __future__ import annotations

import numpy as np

from . import lib as _lib
from .bs import BS


class WN:
    def __init__(self, nx, ny, nzwn, freq=0.0, device=None):
        self.bs = BS(nx, ny, device=device)
        self.nzwn = int(nzwn)
        self.freq = freq if freq is not None else 0.0
        self.zwn = np.zeros(self.nzwn, dtype=float)
        self.mwn = np.zeros((nx, ny, self.nzwn, 3), dtype=float)
        self.rootnum = np.zeros((nx, ny, self.nzwn), dtype=int)
        self.ug = np.zeros((nx, ny, self.nzwn, 3), dtype=float)
        self.vg = np.zeros((nx, ny, self.nzwn, 3), dtype=float)

    def set_zwn(self, zwn_array):
        if len(zwn_array) != self.nzwn:
            raise ValueError("Length of zwn_array must equal nzwn")      # wn.py:47-48
        self.zwn[:] = np.array(zwn_array, dtype=float)

    def set_freq(self, freq):
        self.freq = freq

    def cal_wave(self):
        """wn.py:58-87 as intended: dispersion roots and group velocities at every grid point (call bs.ready first)."""
        import torch
        if self.bs._h is None:
            raise RuntimeError("call bs.ready() before cal_wave")
        nx, ny = self.bs.nlon, self.bs.nlat
        dev = torch.device("cuda", self.bs.device)
        lon = np.repeat(np.asarray(self.bs.lon, dtype=np.float64), ny)        # source index = ix * ny + iy
        lat = np.tile(np.asarray(self.bs.lat, dtype=np.float64), nx)
        d_lon, d_lat = torch.from_numpy(lon).to(dev), torch.from_numpy(lat).to(dev)
        d_zwn = torch.from_numpy(np.ascontiguousarray(self.zwn, dtype=np.float64)).to(dev)
        ns = nx * ny
        out = torch.empty((7, 1, 3, ns, self.nzwn), dtype=torch.float64, device=dev)
        with torch.cuda.device(self.bs.device):
            _lib.check(_lib.load().rwrt_ray_init(self.bs._h, ns, _lib.ptr(d_lon), _lib.ptr(d_lat), None, self.nzwn, _lib.ptr(d_zwn),
                                                 float(self.freq), 1, _lib.ptr(out), _lib.stream_ptr()), "rwrt_ray_init")
        o = out.cpu().numpy()[:, 0]                                          # (7, 3, ns, nzwn)
        mwn = np.moveaxis(o[3], 0, -1).reshape(nx, ny, self.nzwn, 3)         # -> (nlon, nlat, nzwn, 3)
        has = ~np.isnan(mwn)
        self.mwn = mwn
        self.rootnum = has.sum(axis=-1).astype(int)
        self.ug = np.where(has, np.moveaxis(o[5], 0, -1).reshape(nx, ny, self.nzwn, 3), 0.0)   # wn.py:84-87
        self.vg = np.where(has, np.moveaxis(o[6], 0, -1).reshape(nx, ny, self.nzwn, 3), 0.0)

    def clean(self):
        self.bs.clean()
        del self.mwn, self.rootnum, self.ug, self.vg
