"""CPU test: the restated LAPACK pieces behind np.roots' eigenvalue ORDER (oracle/zgeev3.inc; the same
algorithm is in csrc/rwrt_init.cu) against the real library in this environment, called through the
`scipy.linalg.cython_lapack` capsules: ZLARFG, ZLADIV and ZLAHQR bit for bit, and ZGEEV's eigenvalue
order on companion matrices of physically shaped cubics (bs.py:1005-1011)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle

cl = pytest.importorskip("scipy.linalg.cython_lapack")
sl = pytest.importorskip("scipy.linalg.lapack")

C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
C.pythonapi.PyCapsule_GetName.restype = C.c_char_p
C.pythonapi.PyCapsule_GetName.argtypes = [C.py_object]


def fptr(name):
    cap = cl.__pyx_capi__[name]
    return C.pythonapi.PyCapsule_GetPointer(cap, C.pythonapi.PyCapsule_GetName(cap))


class Z(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


dp = C.POINTER(C.c_double)


def test_zladiv_and_zlarfg_bit_exact():
    lib = oracle.lib()
    zladiv = C.CFUNCTYPE(Z, C.POINTER(Z), C.POINTER(Z))(fptr("zladiv"))
    rng = np.random.default_rng(0)
    for i in range(2000):
        x = Z(rng.normal(), rng.normal() if i % 3 else 0.0)
        y = Z(rng.normal(), rng.normal() if i % 2 else 0.0)
        r = zladiv(C.byref(x), C.byref(y))
        o = np.zeros(2)
        xa, ya = np.array([x.re, x.im]), np.array([y.re, y.im])
        lib.orc_dbg_zladiv(xa.ctypes.data_as(dp), ya.ctypes.data_as(dp), o.ctypes.data_as(dp))
        assert (r.re, r.im) == (o[0], o[1])
    for i in range(2000):
        alpha = complex(rng.normal(), rng.normal() if i % 2 else 0.0)
        x = np.array([complex(rng.normal(), 0.0)])
        a_ref, x_ref, tau_ref = sl.zlarfg(2, alpha, x.copy())
        a = np.array([alpha.real, alpha.imag]); xx = np.array([x[0].real, x[0].imag]); t = np.zeros(2)
        lib.orc_dbg_zlarfg2(a.ctypes.data_as(dp), xx.ctypes.data_as(dp), t.ctypes.data_as(dp))
        assert (a_ref.real, a_ref.imag, x_ref[0].real, x_ref[0].imag, tau_ref.real, tau_ref.imag) == \
            (a[0], a[1], xx[0], xx[1], t[0], t[1])


def test_zlahqr_bit_exact():
    lib = oracle.lib()
    lib.orc_dbg_zlahqr.restype = C.c_int
    ip, vp = C.POINTER(C.c_int), C.c_void_p
    zlahqr = C.CFUNCTYPE(None, ip, ip, ip, ip, ip, vp, ip, vp, ip, ip, vp, ip, ip)(fptr("zlahqr"))
    rng = np.random.default_rng(1)
    for n in (2, 3):
        for _ in range(1500):
            H = np.zeros((n, n))
            H[0, :] = rng.normal(size=n) * rng.choice([1, 10, 100])
            for j in range(1, n):
                H[j, j - 1] = 1.0
            A = np.asfortranarray(H.astype(complex)); w = np.zeros(n, complex); z = np.zeros((1, 1), complex)
            one, nn, f, info = C.c_int(1), C.c_int(n), C.c_int(0), C.c_int(0)
            zlahqr(C.byref(f), C.byref(f), C.byref(nn), C.byref(one), C.byref(nn), A.ctypes.data, C.byref(nn),
                   w.ctypes.data, C.byref(one), C.byref(nn), z.ctypes.data, C.byref(one), C.byref(info))
            hin = np.ascontiguousarray(H.astype(complex)).view(np.float64).copy(); wo = np.zeros(2 * n)
            lib.orc_dbg_zlahqr(n, hin.ctypes.data_as(dp), wo.ctypes.data_as(dp))
            assert np.array_equal(w.view(float), wo)


def test_zgeev_order_on_dispersion_cubics():
    rng = np.random.default_rng(2)
    three = 0
    for i in range(3000):
        k = float(rng.integers(1, 9)); U = rng.uniform(-20, 60)
        V = 0.0 if i % 11 == 0 else (rng.uniform(-3, 3) if i % 7 else rng.uniform(-30, 30))
        Qx, Qy = rng.uniform(-30, 30), rng.uniform(-50, 300)
        c = np.array([k ** 3 * (U - Qy / k ** 2), k ** 2 * V + Qx, k * U, V])
        deg = 3
        while deg > 0 and c[deg] == 0:
            deg -= 1
        p = c[: deg + 1][::-1].copy()
        A = np.zeros((deg, deg), complex, order="F")
        A[0, :] = -p[1:] / p[0]
        for j in range(1, deg):
            A[j, j - 1] = 1.0
        ref = sl.zgeev(A, compute_vl=0, compute_vr=0)[0]
        got = oracle.roots(p)
        rr = ref.real[np.abs(ref.imag) < 1e-8]; gr = got.real[np.abs(got.imag) < 1e-8]
        assert np.array_equal(rr, gr), (p, ref, got)
        three += len(rr) == 3
    assert three > 200
