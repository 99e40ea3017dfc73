"""Mirror of the reference's SHSF.py (`SHSF.py:10-33`): spherical-harmonic spectral filtering of a
Driscoll-Healy grid -- expand, truncate at `truncation_level`, rebuild -- on the GPU through the C ABI
(`rwrt_shsf_*`, csrc/rwrt_shsf.cu).  Same name, arguments and return value as the reference function:

    u_filtered = SHSF(ubar, truncation_level)        # (n, 2n) in -> (2L+2, 2L+2) out, like the reference

The reference calls pyshtools (SHExpandDH -> SHCoeffs.pad -> MakeGridDH); pyshtools is not vendored there and
not installed here, so parity is unpinned and what runs is the published algorithm.  Extension:
a 3-D `data` (nmember, n, sampling*n) filters a whole ensemble in one call (the reference accepts 2-D only:
SHExpandDH rejects anything else).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib


class SHSFPlan:
    """Transform tables for one (grid size, truncation level) on one device; reuse it across calls."""

    def __init__(self, nlat: int, truncation_level: int, sampling: int = 2, out_sampling: int = 1, device: int = 0):
        L = lib.load()
        lib.require_gpu()
        self._L, self._h = L, lib.vp()
        self.nlat, self.sampling, self.device = int(nlat), int(sampling), int(device)
        lib.check(L.rwrt_shsf_plan_create(device, nlat, sampling, truncation_level, out_sampling, C.byref(self._h)),
                  "rwrt_shsf_plan_create")
        a, b = lib.i32(0), lib.i32(0)
        lib.check(L.rwrt_shsf_out_shape(self._h, C.byref(a), C.byref(b)), "rwrt_shsf_out_shape")
        self.out_shape = (a.value, b.value)

    def apply_device(self, grid, out=None):
        """grid: torch float64 CUDA tensor (nmember, nlat, sampling*nlat); returns (nmember, *out_shape)."""
        import torch
        assert grid.is_cuda and grid.dtype == torch.float64 and grid.is_contiguous()
        nm = grid.shape[0]
        if out is None:
            out = torch.empty((nm,) + self.out_shape, dtype=torch.float64, device=grid.device)
        lib.check(self._L.rwrt_shsf_apply(self._h, lib.ptr(grid), nm, lib.ptr(out), lib.stream_ptr()), "rwrt_shsf_apply")
        return out

    def close(self):
        if self._h:
            self._L.rwrt_shsf_plan_destroy(self._h)
            self._h = lib.vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def SHSF(data, truncation_level, sampling=2, device=0):
    """SHSF.py:10-33.  data: (n, sampling*n) -- or (nmember, n, sampling*n) -- array on a Driscoll-Healy grid
    (row 0 = north pole); returns the filtered field on the (2L+2) x (2L+2) grid MakeGridDH produces."""
    import torch
    a = np.asarray(data, dtype=np.float64)
    if a.ndim not in (2, 3):
        raise ValueError("data must be 2-D (lat, lon) or 3-D (member, lat, lon)")
    n, nlon = a.shape[-2], a.shape[-1]
    if n % 2 or nlon != sampling * n:
        raise ValueError("grid must be (n, n) or (n, 2n) with n even (Driscoll-Healy sampling)")
    plan = SHSFPlan(n, int(truncation_level), sampling, 1, device)
    try:
        g = torch.from_numpy(np.ascontiguousarray(a.reshape((-1, n, nlon)))).to(torch.device("cuda", device))
        out = plan.apply_device(g).cpu().numpy()
    finally:
        plan.close()
    return out[0] if a.ndim == 2 else out
