"""`BS` — basic-state object with the reference's interface (`bs.py:69-511`), GPU backed.

Same names and argument meaning as the reference class: `BS(nlon, nlat, read_dtype, cal_dtype)`,
array injection (`bs.u`, `bs.v` float32 (nlon, nlat); `bs.lon`, `bs.lat` float64 radians),
`loadbs_ncfile`, `ready(xcyclic)`, `output`, `cal_bs_mercator_point(lon, lat, mode='numpy')`,
`clean`.  The derivative fields are computed by the CUDA stencil kernel behind `rwrt_bs_ready`
(`csrc/rwrt_bs.cu`); the NumPy attributes (`fields`, `q`, `ux`, ... `betam`, `KS`) are copies
fetched from the device on first access.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib
from .constants import pi, undef

FIELD_NAMES = ("u", "v", "ux", "uy", "vx", "vy", "qx", "qy", "qxx", "qxy", "qyx", "qyy",
               "qxxx", "qxxy", "qxyy", "qyyy", "qyxx", "qyyx")            # bs.py:349-368
DIAG_NAMES = ("q", "uxx", "uyy", "vxx", "vyy", "uxy", "vxy", "betam", "KS")


class BS:
    def __init__(self, nlon, nlat, read_dtype="float32", cal_dtype="float64", device=None):
        if cal_dtype != "float64":
            raise ValueError("cal_dtype: only float64 is supported (as in the reference, main_wr.py:21)")
        self.all_dtype = read_dtype
        self.all_dtype_ = cal_dtype
        self.nlon, self.nlat = int(nlon), int(nlat)
        self.dx = np.array([2.0 * pi / self.nlon], dtype=cal_dtype)      # bs.py:77
        self.dy = np.array([pi / (self.nlat - 1)], dtype=cal_dtype)      # bs.py:78
        self.u = np.zeros((self.nlon, self.nlat), dtype=read_dtype)
        self.v = np.zeros((self.nlon, self.nlat), dtype=read_dtype)
        self.lat = np.zeros(self.nlat, dtype=cal_dtype)
        self.lon = np.zeros(self.nlon, dtype=cal_dtype)
        self.xcyclic = True
        self.device = device
        self._h = None
        self._d = {}          # device tensors
        self._host = {}       # lazily fetched host copies

    # ---- reference accessors -------------------------------------------------------------
    def getlon(self):
        return self.lon

    def getlat(self):
        return self.lat

    # ---- device plumbing -----------------------------------------------------------------
    def _torch(self):
        import torch
        _lib.require_gpu()
        if self.device is None:
            self.device = torch.cuda.current_device()
        return torch

    def _handle(self, xcyclic, nmember=1):
        if self._h is not None and self._hkey == (bool(xcyclic), nmember):
            return self._h
        self._release()
        torch = self._torch()
        h = _lib.vp()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().rwrt_create(int(self.device), self.nlon, self.nlat, int(bool(xcyclic)), nmember,
                                               C.byref(h)), "rwrt_create")
        self._h, self._hkey = h, (bool(xcyclic), nmember)
        return h

    def _release(self):
        if getattr(self, "_h", None) is not None:
            _lib.load().rwrt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # ---- IO (bs.py:202-262) --------------------------------------------------------------
    def loadbs_ncfile(self, ncfile):
        from . import ncio
        d = ncio.read_uv(ncfile, self.nlon, self.nlat, read_dtype=self.all_dtype, dx=self.dx, dy=self.dy)
        self.u, self.v, self.lon, self.lat = d["u"], d["v"], d["lon"], d["lat"]

    # ---- BS.ready (bs.py:318-407) ----------------------------------------------------------
    def ready(self, xcyclic=False, keep_fields18=True, keep_diag=True):
        torch = self._torch()
        u = np.ascontiguousarray(self.u, dtype=np.float32)
        v = np.ascontiguousarray(self.v, dtype=np.float32)
        if u.shape != (self.nlon, self.nlat) or v.shape != (self.nlon, self.nlat):
            raise ValueError(f"u, v must have shape (nlon, nlat) = ({self.nlon}, {self.nlat})")
        self.xcyclic = bool(xcyclic)
        h = self._handle(xcyclic)
        dev = torch.device("cuda", self.device)
        lat = np.ascontiguousarray(self.lat, dtype=np.float64)
        lon = np.ascontiguousarray(self.lon, dtype=np.float64)
        d = self._d = {}
        d["u"] = torch.from_numpy(u).to(dev)
        d["v"] = torch.from_numpy(v).to(dev)
        d["lon"] = torch.from_numpy(lon).to(dev)
        d["lat"] = torch.from_numpy(lat).to(dev)
        # NumPy's cos/sin of the grid latitudes, as bs.py:270-277 uses them: makes the fields bit-identical
        d["coslat"] = torch.from_numpy(np.cos(lat)).to(dev)
        d["sinlat"] = torch.from_numpy(np.sin(lat)).to(dev)
        nxf = self.nlon + (1 if xcyclic else 0)
        d["fields18"] = torch.empty((nxf, self.nlat, 18), dtype=torch.float64, device=dev) if keep_fields18 else None
        d["diag9"] = torch.empty((9, self.nlon, self.nlat), dtype=torch.float64, device=dev) if keep_diag else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().rwrt_bs_ready(h, _lib.ptr(d["u"]), _lib.ptr(d["v"]), _lib.ptr(d["lon"]),
                                                 _lib.ptr(d["lat"]), _lib.ptr(d["coslat"]), _lib.ptr(d["sinlat"]),
                                                 _lib.ptr(d["fields18"]), _lib.ptr(d["diag9"]), _lib.stream_ptr()),
                       "rwrt_bs_ready")
        self._host = {}

    # ---- lazily fetched NumPy views of the device results ---------------------------------
    def __getattr__(self, name):
        # only called when normal lookup fails
        if name.startswith("_"):
            raise AttributeError(name)
        d = self.__dict__.get("_d", {})
        host = self.__dict__.get("_host", {})
        if name == "fields":
            if d.get("fields18") is None:
                raise AttributeError("fields: call ready() (with keep_fields18=True) first")
            if "fields" not in host:
                host["fields"] = d["fields18"].cpu().numpy()
            return host["fields"]
        if name in FIELD_NAMES[2:]:
            k = FIELD_NAMES.index(name)
            return self.fields[: self.nlon, :, k]
        if name in DIAG_NAMES:
            if d.get("diag9") is None:
                raise AttributeError(f"{name}: call ready() (with keep_diag=True) first")
            if "diag" not in host:
                host["diag"] = d["diag9"].cpu().numpy()
            return host["diag"][DIAG_NAMES.index(name)]
        raise AttributeError(name)

    # ---- BS.cal_bs_mercator_point (bs.py:513-887) ------------------------------------------
    def cal_bs_mercator_point(self, lon, lat, mode="numpy"):
        """The reference's 18 rows (fmu ... fmqyy, fmqxxx ... fmqyyx) at the given points.  When ready() was
        called with keep_fields18=False only rows 0..11 (what the ray path consumes) are computed; 12..17 are NaN."""
        if mode != "numpy":
            raise NotImplementedError("only mode='numpy' (the vectorised path) exists on the GPU engine")
        if self._h is None:
            raise RuntimeError("call ready() first")
        torch = self._torch()
        dev = torch.device("cuda", self.device)
        lon = np.atleast_1d(np.asarray(lon, dtype=np.float64))
        lat = np.atleast_1d(np.asarray(lat, dtype=np.float64))
        n = lon.shape[0]
        dl, da = torch.from_numpy(lon).to(dev), torch.from_numpy(lat).to(dev)
        out = torch.full((18, n), float("nan"), dtype=torch.float64, device=dev)
        with torch.cuda.device(self.device):
            if self._d.get("fields18") is not None:
                _lib.check(_lib.load().rwrt_sample_mercator18(self._h, _lib.ptr(self._d["fields18"]), n, _lib.ptr(dl),
                                                              _lib.ptr(da), None, _lib.ptr(out), _lib.stream_ptr()),
                           "rwrt_sample_mercator18")
            else:
                _lib.check(_lib.load().rwrt_sample_mercator(self._h, n, _lib.ptr(dl), _lib.ptr(da), None,
                                                            _lib.ptr(out), _lib.stream_ptr()), "rwrt_sample_mercator")
        return out.cpu().numpy()

    def output(self, ncfile):
        """bs.py:461-511: 23 variables (lon, lat) with the reference's names (and its radian lon/lat)."""
        from . import ncio
        fields = {"u": (np.asarray(self.u), "m/s"), "v": (np.asarray(self.v), "m/s"), "q": (self.q, "1/s")}
        for nm in ("ux", "uxx", "uy", "vx", "vxx", "vy", "qx", "qy", "qxx", "qxy", "qyx", "qyy", "qxxx", "qxxy",
                   "qxyy", "qyyy", "qyxx", "qyyx"):
            fields[nm] = (getattr(self, nm), "None")
        fields["betam"] = (self.betam, "1/(m·s)")
        fields["KS"] = (self.KS, "1/m")
        ncio.write_bsfile(ncfile, self.getlon(), self.getlat(), fields)

    def clean(self):
        self._d = {}
        self._host = {}
        self._release()


__all__ = ["BS", "FIELD_NAMES", "DIAG_NAMES", "undef"]
