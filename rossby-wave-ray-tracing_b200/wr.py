"""`WR` — wave-ray tracer with the reference's interface (`wr.py:114-977`), GPU backed.

`WR(nzwn, nsource, tstep, ttotal, freq, ..., rtol, atol, cut_off, nx, ny, ncfile, MinStepFactor)`,
`set_zwn`, `set_freq`, `set_source_array`, `set_source_matrix`, `ray_info`, `ray_run(mode,
inte_method, root_method)`, `output`, `clean` and the result arrays `rlon, rlat, rzwn, rmwn, ramp,
rug, rvg` of shape `(nt, 3, nsource, nzwn)` (radians, NaN = not propagating) behave as in the
reference.  Only `mode='numpy', inte_method='rk45'` — the path this engine replaces — is
implemented; anything else raises (there is no CPU fallback).

Additions (all optional): `accepted` / `status` per slot, `total_accepted`, `device=` and
`to_host=` (keep results on the GPU as `d_out`), `strict=` (reference operation order, no FMA).
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import lib as _lib
from .bs import BS
from .constants import day, deg2rad, hour, rad2deg, undef

NAMES = ("rlon", "rlat", "rzwn", "rmwn", "ramp", "rug", "rvg")


class WR:
    def __init__(self, nzwn, nsource, tstep=1. * hour, ttotal=20. * day, freq=0, cal_dtype="float64",
                 read_dtype="float32", rtol=1e-6, atol=1e-6, cut_off=0.1, nx=None, ny=None, ncfile=None,
                 MinStepFactor=1e-3, device=None):
        self.all_dtype = cal_dtype
        if (nx is None) or (ny is None):
            if ncfile is None:
                raise ValueError("ncfile is need")                       # wr.py:140-141
            ny, nx = self.get_defualt_nlon_and_nlat(ncfile)
        self.bs = BS(nx, ny, read_dtype=read_dtype, cal_dtype=cal_dtype, device=device)
        self.tstep = np.array([tstep], dtype=self.all_dtype)             # wr.py:147
        self.ttotal = ttotal
        self.freq = np.array([freq], dtype=self.all_dtype)
        self.nzwn = int(nzwn)
        self.zwn = np.zeros(self.nzwn, dtype=self.all_dtype)
        self.nsource = int(nsource)
        self.source_lon = np.zeros(self.nsource, dtype=self.all_dtype)
        self.source_lat = np.zeros(self.nsource, dtype=self.all_dtype)
        self.nt = int(self.ttotal / self.tstep[0]) + 1                   # wr.py:157
        self.rtol = rtol
        self.atol = atol
        self.cut_off_factor = cut_off
        self.cut_off = cut_off * self.tstep / 3600.                      # wr.py:170
        self.MinStepFactor = MinStepFactor
        self.device = device
        self.d_out = None            # torch (7, nt, 3, nsource, nzwn) on the GPU after ray_run
        self.accepted = None         # (3, nsource, nzwn) int64 accepted RK45 steps per slot
        self.status = None           # (3, nsource, nzwn) int8, lib.RAY_*
        self.total_accepted = 0
        self._host = None
        self._host_t = None
        self._host_stale = False     # d_out holds results that the host planes do not (ray_run(..., to_host=False))
        self.pin_results = False     # result planes in pinned host memory (see _host_planes)

    # ---- result arrays: allocated on first use (the reference allocates them in __init__) ----
    def _host_planes(self, fill=True):
        """The seven result planes on the host.  Pageable NumPy memory by default: the delivery from the GPU writes it
        with host threads (rwrt_ray_run_rk45_d2h), so pinning 23 GB (seconds) buys nothing; `pin_results = True`
        before the first use gives pinned memory (faster `to_host()` after the RK4 path or repeated runs).
        fill=False skips the NaN pre-fill of wr.py:160-167 when a run is about to overwrite every element."""
        if self._host is None:
            shape = (7, self.nt, 3, self.nsource, self.nzwn)
            self._host_t = None
            if self.pin_results:
                try:
                    import torch
                    self._host_t = torch.empty(shape, dtype=torch.float64, pin_memory=torch.cuda.is_available())
                    self._host = self._host_t.numpy()
                except Exception:
                    self._host_t = None
            if self._host_t is None:
                self._host = np.empty(shape, dtype=np.float64)
            if fill:
                self._host.fill(undef)                                   # wr.py:160-167
        return self._host

    def __getattr__(self, name):
        if name in NAMES:
            # results that exist only on the GPU (ray_run(to_host=False), dist.run_sharded's default) are fetched on
            # first access -- never the NaN pre-fill of wr.py:160-167 while d_out holds a newer run
            if self.__dict__.get("_host_stale") and self.__dict__.get("d_out") is not None:
                self.to_host()
            return self._host_planes()[NAMES.index(name)]
        raise AttributeError(name)

    def get_defualt_nlon_and_nlat(self, ncfile):                         # wr.py:173-212 (sic)
        from . import ncio
        return ncio.grid_shape(ncfile)

    def set_zwn(self, zwn_array):
        if len(zwn_array) != self.nzwn:
            raise ValueError("Length of zwn_array must equal nzwn")      # wr.py:216-217
        self.zwn[:] = np.array(zwn_array, dtype=self.all_dtype)

    def set_freq(self, freq):
        self.freq = np.array([freq], dtype=self.all_dtype)

    def set_source_array(self, lon_list, lat_list):
        if len(lon_list) != self.nsource or len(lat_list) != self.nsource:
            raise ValueError("Source list length mismatch nsource")      # wr.py:229-230
        self.source_lon = np.array(lon_list, dtype=self.all_dtype) * (1.0 * deg2rad)
        self.source_lat = np.array(lat_list, dtype=self.all_dtype) * (1.0 * deg2rad)

    def set_source_matrix(self, SW_lon, SW_lat, dlon, dlat, nnx, nny):
        if nnx * nny != self.nsource:
            raise ValueError("nsource != nnx * nny, matrix size mismatch!")   # wr.py:242-243
        if SW_lat + (nny - 1) * dlat > 89.0:
            raise ValueError("source latitude out of -90~90 range!")          # wr.py:245-246
        SW_lon = SW_lon % 360.0
        ix = np.arange(nnx)
        iy = np.arange(nny)
        lon_deg = (SW_lon + ix * dlon) % 360.0                                # wr.py:252-255
        lat_deg = SW_lat + iy * dlat
        self.source_lon = np.tile(lon_deg * deg2rad, nny).astype(self.all_dtype)
        self.source_lat = np.repeat(lat_deg * deg2rad, nnx).astype(self.all_dtype)

    def ray_info(self, file=None, max_sources=50):
        """wr.py:260-279; the per-source listing is capped (10^5+ sources would flood stdout)."""
        f = file or sys.stdout
        bar = "=" * 78
        print(bar, file=f)
        print(" WNWR Package: Barotropic Horizontal Rossby Wave Ray Tracing Information ", file=f)
        print(f" Shape of the Basic Flow (nlon x nlat): {self.bs.nlon} x {self.bs.nlat}", file=f)
        print(f" Initial Zonal Wave Numbers (nzwn): {self.nzwn}", file=f)
        print(" " * 15 + " ".join(f"{z:.1f}" for z in self.zwn), file=f)
        print(f" Source Locations (total {self.nsource} points):", file=f)
        for i in range(min(self.nsource, max_sources)):
            print(" " * 15 + f"{self.source_lon[i] * rad2deg:7.2f}, {self.source_lat[i] * rad2deg:7.2f}", file=f)
        if self.nsource > max_sources:
            print(" " * 15 + f"... ({self.nsource - max_sources} more)", file=f)
        print(f" Time Step (s): {self.tstep[0]:.1f}", file=f)
        print(f" Total Integration Time (day): {self.ttotal / day:.1f}", file=f)
        print(f" Total Steps (nt): {self.nt}", file=f)
        print(bar, file=f)

    # ---- the ray path ----------------------------------------------------------------------
    def _check_ready(self):
        if self.bs._h is None:
            raise RuntimeError("call bs.ready() before ray_run (main_wr.py:81)")

    def ray_initial(self, mode="numpy", root_method="numpy"):
        """wr.py:344-395 / 417-421: fills row 0 of the seven result planes on the GPU."""
        if mode != "numpy":
            raise NotImplementedError("ray_initial: only mode='numpy' exists on the GPU engine")
        self._check_ready()
        import torch
        dev = torch.device("cuda", self.bs.device)
        self._d_src_lon = torch.from_numpy(np.ascontiguousarray(self.source_lon, dtype=np.float64)).to(dev)
        self._d_src_lat = torch.from_numpy(np.ascontiguousarray(self.source_lat, dtype=np.float64)).to(dev)
        self._d_zwn = torch.from_numpy(np.ascontiguousarray(self.zwn, dtype=np.float64)).to(dev)
        self.d_out = torch.empty((7, self.nt, 3, self.nsource, self.nzwn), dtype=torch.float64, device=dev)
        with torch.cuda.device(self.bs.device):
            _lib.check(_lib.load().rwrt_ray_init(self.bs._h, self.nsource, _lib.ptr(self._d_src_lon),
                                                 _lib.ptr(self._d_src_lat), None, self.nzwn, _lib.ptr(self._d_zwn),
                                                 float(self.freq[0]), self.nt, _lib.ptr(self.d_out),
                                                 _lib.stream_ptr()), "rwrt_ray_init")

    def core_ray_run(self, mode="numpy_rk45", strict=False, to_host=False):
        """wr.py:767-887 (mode 'numpy_rk45').  to_host: also deliver the seven planes to the pinned NumPy
        arrays, overlapped with the integration (time windows, `rwrt_ray_run_rk45_d2h`)."""
        if mode != "numpy_rk45":
            raise NotImplementedError("only the RK45 path (mode='numpy', inte_method='rk45') exists on the GPU engine")
        import torch
        dev = torch.device("cuda", self.bs.device)
        N = 3 * self.nsource * self.nzwn
        self._d_acc = torch.zeros(N, dtype=torch.int64, device=dev)
        self._d_status = torch.zeros(N, dtype=torch.int8, device=dev)
        self._d_total = torch.zeros(1, dtype=torch.int64, device=dev)
        flags = _lib.MODE_STRICT if strict else _lib.MODE_FAST
        args = (self.bs._h, self.nsource, None, self.nzwn, self.nt, float(self.tstep[0]), float(self.ttotal),
                float(self.rtol), float(self.atol), float(self.MinStepFactor), float(self.cut_off_factor), flags,
                _lib.ptr(self.d_out), _lib.ptr(self._d_acc), None, _lib.ptr(self._d_status), _lib.ptr(self._d_total))
        with torch.cuda.device(self.bs.device):
            if to_host and self._host is not None and N > 0:
                _lib.check(_lib.load().rwrt_ray_run_rk45_d2h(*args, _lib.ptr(self._host), _lib.stream_ptr()),
                           "rwrt_ray_run_rk45_d2h")
                return True
            _lib.check(_lib.load().rwrt_ray_run_rk45(*args, _lib.stream_ptr()), "rwrt_ray_run_rk45")
        return False

    def core_ray_run_rk4(self, strict=False):
        """wr.py:702-765 (mode 'numpy'): fixed-step RK4, the reference's default integrator."""
        import torch
        dev = torch.device("cuda", self.bs.device)
        N = 3 * self.nsource * self.nzwn
        self._d_status = torch.zeros(N, dtype=torch.int8, device=dev)
        flags = _lib.MODE_STRICT if strict else _lib.MODE_FAST
        with torch.cuda.device(self.bs.device):
            _lib.check(_lib.load().rwrt_ray_run_rk4(self.bs._h, self.nsource, None, self.nzwn, self.nt,
                                                    float(self.tstep[0]), float(self.cut_off_factor), flags,
                                                    _lib.ptr(self.d_out), _lib.ptr(self._d_status), _lib.stream_ptr()),
                       "rwrt_ray_run_rk4")

    def ray_run(self, mode="original", inte_method="", root_method="Fortran", debug=False, debug_file=None,
                to_host=True, strict=False):
        """wr.py:897-911.  mode='numpy' with inte_method='rk45' (adaptive RK45) or anything else (fixed-step
        RK4, the reference's default) run on the GPU; the scalar mode='original' does not exist here."""
        if mode != "numpy":
            raise NotImplementedError(
                "rwrt_b200 implements ray_run(mode='numpy', ...) only; the reference's scalar 'original' mode is "
                "not part of the GPU path and there is no CPU fallback")
        if debug:
            raise NotImplementedError("debug / debug_file (Fortran parity hooks, wr.py:901-906) are not supported")
        import torch
        self.ray_initial(mode="numpy", root_method=root_method)
        shp = (3, self.nsource, self.nzwn)
        if inte_method == "rk45":                                       # wr.py:907-908
            if to_host:
                self._host_planes(fill=False)    # the result arrays exist before the run, which writes every element
            try:
                delivered = self.core_ray_run("numpy_rk45", strict=strict, to_host=to_host)
                torch.cuda.synchronize(self.bs.device)
            except Exception:
                if self._host is not None:
                    self._host.fill(undef)       # a failed run leaves "not propagating" everywhere, like wr.py:160-167
                raise
            self.total_accepted = int(self._d_total.item())
            self.accepted = self._d_acc.cpu().numpy().reshape(shp)
        else:
            delivered = False
            self.core_ray_run_rk4(strict=strict)
            torch.cuda.synchronize(self.bs.device)
            self.total_accepted = 0
            self.accepted = None
        self.status = self._d_status.cpu().numpy().reshape(shp)
        self._host_stale = not delivered
        if to_host and not delivered:
            self.to_host()

    def to_host(self):
        """Copy the seven result planes from the GPU into the (pinned) NumPy arrays."""
        import torch
        host = self._host_planes(fill=False)
        dst = self._host_t if self._host_t is not None else torch.from_numpy(host)
        if self.d_out is None:
            raise RuntimeError("no results on the GPU: call ray_run first (or the results were released by clean())")
        dst.copy_(self.d_out, non_blocking=self._host_t is not None)
        torch.cuda.synchronize(self.bs.device)
        self._host_stale = False
        return host

    def output(self, ncfile):
        """wr.py:916-959: dims (time, root, source, zwn); lon/lat written in degrees."""
        from . import ncio
        ncio.write_rayfile(ncfile, self.zwn, self.nsource, self.nt,
                           {n: getattr(self, n) for n in NAMES}, rad2deg)

    def clean(self):
        if self._host_stale and self.d_out is not None:
            import warnings
            warnings.warn("WR.clean(): the results of the last ray_run(to_host=False) were never copied to the host "
                          "and are released now", RuntimeWarning, stacklevel=2)
        self._host_stale = False
        self.bs.clean()
        self.d_out = None
        self._host = None
        self._host_t = None
        for attr in ("source_lon", "source_lat", "zwn"):
            if hasattr(self, attr):
                delattr(self, attr)
