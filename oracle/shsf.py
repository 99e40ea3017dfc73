"""TEST INFRASTRUCTURE -- CPU restatement of the reference's optional background pre-filter SHSF
(`SHSF.py:10-33`): spherical-harmonic analysis of a Driscoll-Healy grid, triangular truncation, synthesis.

    SHSF(data, L, sampling=2) = MakeGridDH( pad( SHExpandDH(data, sampling), L ) )

**parity unpinned.**  The reference delegates all arithmetic to the un-vendored third-party package
`pyshtools` (no version pin; not installable here, no network), so no output of the reference itself can be
produced in this container.  What is restated is the *published* algorithm pyshtools documents for these
calls (Driscoll & Healy 1994 quadrature; 4-pi normalised real harmonics without Condon-Shortley phase, the
package defaults `norm=1, csphase=1`; `MakeGridDH` default `sampling=1`, i.e. an n x n output grid with
n = 2 L + 2).  The restatement is anchored on mathematical known-answer tests instead (tests/test_shsf.py):
orthonormality of the basis under the quadrature, exact recovery of band-limited fields, agreement of the
basis functions with scipy.special.sph_harm_y, closed-form fields.

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

LD = np.longdouble


def dh_weights(n: int) -> np.ndarray:
    """Driscoll-Healy quadrature weights for  int_0^pi g(theta) sin(theta) dtheta  on theta_j = pi j / n,
    j = 0..n-1 (n even): exact for polynomials in cos(theta) of degree < n.  Sum = 2."""
    j = np.arange(n, dtype=LD)
    th = LD(np.pi) * j / n
    k = np.arange(n // 2, dtype=LD)
    s = (np.sin((2 * k[None, :] + 1) * th[:, None]) / (2 * k[None, :] + 1)).sum(axis=1)
    return (LD(4.0) / n) * np.sin(th) * s


def plm_bar(lmax: int, x: np.ndarray) -> np.ndarray:
    """4-pi normalised associated Legendre functions without Condon-Shortley phase,
    P[l, m, i] = sqrt((2 - delta_m0)(2l+1)(l-m)!/(l+m)!) P_lm(x_i), by the standard stable recurrences in
    extended precision (sin^m(theta) underflows double for m > ~300 near the poles)."""
    x = np.asarray(x, dtype=LD)
    s = np.sqrt(np.maximum(LD(1) - x * x, LD(0)))
    P = np.zeros((lmax + 1, lmax + 1, x.size), dtype=LD)
    pmm = np.ones_like(x)
    for m in range(lmax + 1):
        if m == 1:
            pmm = np.sqrt(LD(3)) * s
        elif m > 1:
            pmm = np.sqrt(LD(2 * m + 1) / LD(2 * m)) * s * pmm
        P[m, m] = pmm
        if m + 1 <= lmax:
            P[m + 1, m] = np.sqrt(LD(2 * m + 3)) * x * pmm
        for l in range(m + 2, lmax + 1):
            a = np.sqrt(LD(4 * l * l - 1) / LD(l * l - m * m))
            b = np.sqrt(LD((l - 1) ** 2 - m * m) / LD(4 * (l - 1) ** 2 - 1))
            P[l, m] = a * (x * P[l - 1, m] - b * P[l - 2, m])
    return P


def sh_expand_dh(grid: np.ndarray, sampling: int = 2) -> np.ndarray:
    """cilm[2, n/2, n/2] of an (n, sampling*n) grid: colatitude theta_j = pi j / n (row 0 = north pole),
    longitude phi_k = 2 pi k / nlon.  c_lm = (1/4pi) int f Y_lm dOmega with the DH quadrature."""
    grid = np.asarray(grid, dtype=np.float64)
    n, nlon = grid.shape
    if n % 2 or nlon != sampling * n:
        raise ValueError("DH grid must be (n, n) or (n, 2n) with n even")
    lmax = n // 2 - 1
    th = np.pi * np.arange(n) / n
    w = dh_weights(n)
    P = plm_bar(lmax, np.cos(th.astype(LD)))
    phi = 2 * np.pi * np.arange(nlon) / nlon
    cilm = np.zeros((2, lmax + 1, lmax + 1))
    for m in range(lmax + 1):
        fc = (grid * np.cos(m * phi)[None, :]).sum(axis=1).astype(LD) * (LD(2 * np.pi) / nlon)
        fs = (grid * np.sin(m * phi)[None, :]).sum(axis=1).astype(LD) * (LD(2 * np.pi) / nlon)
        for l in range(m, lmax + 1):
            cilm[0, l, m] = float((w * P[l, m] * fc).sum() / (4 * LD(np.pi)))
            cilm[1, l, m] = float((w * P[l, m] * fs).sum() / (4 * LD(np.pi)))
    return cilm


def pad(cilm: np.ndarray, lmax_new: int) -> np.ndarray:
    """SHCoeffs.pad: truncate or zero-pad to degree lmax_new."""
    out = np.zeros((2, lmax_new + 1, lmax_new + 1))
    k = min(lmax_new, cilm.shape[1] - 1) + 1
    out[:, :k, :k] = cilm[:, :k, :k]
    return out


def make_grid_dh(cilm: np.ndarray, sampling: int = 1) -> np.ndarray:
    """(n, sampling*n) grid with n = 2 lmax + 2 from 4-pi normalised coefficients."""
    lmax = cilm.shape[1] - 1
    n = 2 * lmax + 2
    nlon = sampling * n
    th = np.pi * np.arange(n) / n
    P = plm_bar(lmax, np.cos(th.astype(LD)))
    phi = 2 * np.pi * np.arange(nlon) / nlon
    out = np.zeros((n, nlon), dtype=LD)
    for m in range(lmax + 1):
        a = np.einsum("l,li->i", cilm[0, m:, m].astype(LD), P[m:, m])
        b = np.einsum("l,li->i", cilm[1, m:, m].astype(LD), P[m:, m])
        out += a[:, None] * np.cos(m * phi)[None, :] + b[:, None] * np.sin(m * phi)[None, :]
    return out.astype(np.float64)


def shsf(data: np.ndarray, truncation_level: int, sampling: int = 2) -> np.ndarray:
    """SHSF.py:10-33."""
    return make_grid_dh(pad(sh_expand_dh(data, sampling), truncation_level))
