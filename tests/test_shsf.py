"""SHSF (reference SHSF.py:10-33).  pyshtools, which does all of the reference's arithmetic here, is
neither vendored nor installable in this container: **parity unpinned**.  The CPU tests therefore pin the oracle
restatement (oracle/shsf.py) on mathematical known answers of the published algorithm; the GPU tests compare the
CUDA path (rwrt_shsf_*, through the C ABI) with that oracle on the same seeded grids.  Tolerance: 1e-11 absolute on
fields of order 10 (plain fp64 sums of <= 720 terms; no reference tolerance exists for this row)."""
import numpy as np
import pytest

import rwrt_b200  # noqa: F401
from oracle import shsf as O


def random_coeffs(lmax, seed):
    rng = np.random.default_rng(seed)
    c = np.zeros((2, lmax + 1, lmax + 1))
    for l in range(lmax + 1):
        c[0, l, :l + 1] = rng.normal(size=l + 1) / (1 + l)
        c[1, l, 1:l + 1] = rng.normal(size=l) / (1 + l)
    return c


def synth_on(c, n, nlon):
    """direct evaluation of sum c_lm Y_lm on the (n, nlon) DH grid -- independent of make_grid_dh's grid choice"""
    lmax = c.shape[1] - 1
    th = np.pi * np.arange(n) / n
    P = O.plm_bar(lmax, np.cos(th.astype(np.longdouble))).astype(float)
    phi = 2 * np.pi * np.arange(nlon) / nlon
    g = np.zeros((n, nlon))
    for m in range(lmax + 1):
        a = np.einsum("l,li->i", c[0, m:, m], P[m:, m])
        b = np.einsum("l,li->i", c[1, m:, m], P[m:, m])
        g += a[:, None] * np.cos(m * phi) + b[:, None] * np.sin(m * phi)
    return g


# ----------------------------------------------------------------------------------------- oracle known answers (CPU)
def test_dh_weights_are_exact_for_degree_below_n():
    from numpy.polynomial import legendre as Lg
    for n in (8, 36, 180):
        w = O.dh_weights(n)
        th = np.pi * np.arange(n) / n
        assert abs(float(w.sum()) - 2.0) < 1e-15
        for l in range(n):
            got = float((w * Lg.legval(np.cos(th), [0] * l + [1])).sum())
            assert abs(got - (2.0 if l == 0 else 0.0)) < 5e-15, (n, l)


def test_basis_matches_scipy_and_is_orthonormal_under_the_quadrature():
    from scipy.special import sph_harm_y
    n, lmax = 36, 17
    th = np.pi * np.arange(n) / n
    P = O.plm_bar(lmax, np.cos(th.astype(np.longdouble)))
    for l, m in ((0, 0), (5, 0), (7, 4), (17, 17), (16, 1)):
        y = sph_harm_y(l, m, th, 0.0).real            # orthonormal, Condon-Shortley phase
        ref = np.sqrt(4 * np.pi) * (np.sqrt(2.0) if m else 1.0) * y * (-1.0) ** m
        assert np.abs(P[l, m].astype(float) - ref).max() < 5e-14, (l, m)
    w = O.dh_weights(n)
    for m in (0, 3, 11):
        G = np.einsum("j,lj,kj->lk", w, P[m:, m], P[m:, m]).astype(float)
        assert np.abs(G - (2.0 if m == 0 else 4.0) * np.eye(lmax + 1 - m)).max() < 1e-13, m


@pytest.mark.parametrize("sampling", [1, 2])
def test_expand_recovers_band_limited_coefficients(sampling):
    n = 36
    c = random_coeffs(n // 2 - 1, 1)
    g = synth_on(c, n, sampling * n)
    assert np.abs(O.sh_expand_dh(g, sampling) - c).max() < 1e-13
    assert np.abs(O.make_grid_dh(c, 1) - synth_on(c, n, n)).max() < 1e-13


def test_shsf_closed_forms():
    n = 24
    th = np.pi * np.arange(n) / n
    phi = 2 * np.pi * np.arange(2 * n) / (2 * n)
    f = 3.0 + 2.0 * np.cos(th)[:, None] + 0 * phi[None, :] + np.sin(th)[:, None] ** 2 * np.cos(2 * phi)[None, :]
    L = 4
    out = O.shsf(f, L)
    assert out.shape == (2 * L + 2, 2 * L + 2)            # MakeGridDH's default sampling = 1 (SHSF.py:31)
    th2 = np.pi * np.arange(2 * L + 2) / (2 * L + 2)
    ph2 = 2 * np.pi * np.arange(2 * L + 2) / (2 * L + 2)
    exact = 3.0 + 2.0 * np.cos(th2)[:, None] + np.sin(th2)[:, None] ** 2 * np.cos(2 * ph2)[None, :]
    assert np.abs(out - exact).max() < 1e-13              # degrees 0, 1, 2 survive a truncation at 4 unchanged
    out1 = O.shsf(f, 1)                                   # truncation at 1 removes the degree-2 term
    th1 = np.pi * np.arange(4) / 4
    assert np.abs(out1 - (3.0 + 2.0 * np.cos(th1)[:, None] + np.zeros((4, 4)))).max() < 1e-13
    # truncation above the resolved band zero-pads (SHCoeffs.pad): same field on a finer grid
    big = O.shsf(f, 20)
    assert big.shape == (42, 42)
    th3, ph3 = np.pi * np.arange(42) / 42, 2 * np.pi * np.arange(42) / 42
    assert np.abs(big - (3.0 + 2.0 * np.cos(th3)[:, None] + np.sin(th3)[:, None] ** 2 * np.cos(2 * ph3)[None, :])).max() < 1e-12


def test_facade_mirrors_the_reference_signature():
    import inspect
    from rwrt_b200.SHSF import SHSF
    assert list(inspect.signature(SHSF).parameters)[:3] == ["data", "truncation_level", "sampling"]
    assert inspect.signature(SHSF).parameters["sampling"].default == 2          # SHSF.py:10
    with pytest.raises(ValueError):
        SHSF(np.zeros((5, 10)), 2)                                              # odd n is not a DH grid


# ----------------------------------------------------------------------------------------- CUDA path vs oracle (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("n,L,sampling", [(36, 9, 2), (72, 35, 2), (72, 50, 1), (180, 44, 2), (180, 89, 2)])
def test_gpu_shsf_matches_oracle(n, L, sampling):
    from rwrt_b200.SHSF import SHSF
    rng = np.random.default_rng(n + L)
    th = np.pi * np.arange(n) / n
    # a wind-like field: jets + waves + grid-scale noise (not band limited: the quadrature's aliasing is part of the answer)
    phi = 2 * np.pi * np.arange(sampling * n) / (sampling * n)
    f = (20 * np.sin(th)[:, None] ** 2 + 8 * np.cos(3 * phi)[None, :] * np.sin(th)[:, None] ** 3 +
         rng.normal(size=(n, sampling * n)))
    got = SHSF(f, L, sampling)
    ref = O.shsf(f, L, sampling)
    assert got.shape == ref.shape == (2 * L + 2, 2 * L + 2)
    assert np.abs(got - ref).max() < 1e-11


@pytest.mark.gpu
def test_gpu_shsf_ensemble_linearity_and_projection():
    from rwrt_b200.SHSF import SHSF
    n, L = 72, 20
    rng = np.random.default_rng(3)
    batch = rng.normal(size=(5, n, 2 * n))
    out = SHSF(batch, L)
    assert out.shape == (5, 2 * L + 2, 2 * L + 2)
    for k in range(5):                                      # a batch is the same as its members one by one (bit for bit)
        assert np.array_equal(out[k], SHSF(batch[k], L))
    lin = SHSF(2.0 * batch[0] - 3.0 * batch[1], L)
    assert np.abs(lin - (2.0 * out[0] - 3.0 * out[1])).max() < 1e-12
    # projection: a field that is band limited at L passes unchanged (evaluated on the output grid)
    c = random_coeffs(L, 9)
    g = synth_on(c, n, 2 * n)
    assert np.abs(SHSF(g, L) - synth_on(c, 2 * L + 2, 2 * L + 2)).max() < 1e-12
