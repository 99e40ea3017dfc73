// rwrt_init.cu -- K2: t = 0 ray state (WR.ray_initial_numpy, wr.py:344-395) and the point sampler.
//
// Per (source, k): bilinear sample + Mercator transform at the source, the dispersion cubic
//     V l^3 + k(U - c) l^2 + (k^2 V + Qx) l + k^3 (U - c) - k Qy = 0,  c = omega a / k   (bs.py:1005-1011)
// its real roots (|Im| < 1e-8, bs.py:1030), the reference's slot shuffle (bs.py:942-982) and the t = 0
// group velocity (wn.py:209-259).
//
// The slot shuffle depends on the ORDER in which numpy.roots (numba -> LAPACK zgeev on the
// companion matrix, bs.py:38-40) returns the eigenvalues, so the eigenvalue iteration below follows
// the published LAPACK algorithm for n <= 3 operation by operation: ZGEBAL('B') power-of-two
// balancing, ZGEHRD (identity for a companion matrix), ZLAHQR single-shift QR with Wilkinson /
// exceptional shifts, Ahues-Tisseur deflation, ZLARFG reflectors, ZLADIV (Baudin-Smith).  This
// file is compiled with -fmad=false so the arithmetic is the same sequence of IEEE operations.
#include "rwrt_internal.h"

namespace rwrt {
namespace {

struct zc {
    double re, im;
};
__device__ __forceinline__ zc zmk(double re, double im) { return zc{re, im}; }
__device__ __forceinline__ zc zadd(zc a, zc b) { return zmk(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ zc zsub(zc a, zc b) { return zmk(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ zc zmul(zc a, zc b) { return zmk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ zc zrmul(double s, zc a) { return zmk(s * a.re, s * a.im); }
__device__ __forceinline__ zc zrdiv(zc a, double s) { return zmk(a.re / s, a.im / s); }
__device__ __forceinline__ zc zconj(zc a) { return zmk(a.re, -a.im); }
__device__ __forceinline__ zc zneg(zc a) { return zmk(-a.re, -a.im); }
__device__ __forceinline__ double zabs1(zc a) { return fabs(a.re) + fabs(a.im); }
// hypot as glibc 2.35+ computes it on x86-64 (the libm behind the reference's LAPACK `ABS(complex)` and csqrt): the
// naive sqrt(x^2 + y^2) followed by the correction step of Borges, "An improved algorithm for hypot(a, b)" (2019),
// non-FMA form.  It is NOT correctly rounded (0.15 % of random inputs differ from the rounded exact value) and CUDA's
// hypot() differs from it in the last bit more often than that, which moved 4 % of the cubic roots by an ulp against the
// reference's own vectors.  Checked bit for bit against libm on 2e7 random inputs; this unit is built with -fmad=false.
// (Magnitudes here are O(1e-3 .. 1e6): the library's rescaling branches for |x| > 2^511 / < 2^-459 cannot be reached.)
__device__ __forceinline__ double hypot_libm(double x, double y)
{
    double ax = fabs(x), ay = fabs(y);
    if (ax < ay) { const double t = ax; ax = ay; ay = t; }
    if (!(ax == ax) || !(ay == ay) || ax > 1.0e300 || (ay < 1.0e-130 && ay != 0.0)) return hypot(x, y);
    if (ay <= ax * 5.5511151231257827e-17) return ax + ay;          // 2^-54: y does not reach the last bit of x
    double h = sqrt(ax * ax + ay * ay), t1, t2;
    if (h <= 2.0 * ay) {
        const double d = h - ay;
        t1 = ax * (2.0 * d - ax);
        t2 = (d - 2.0 * (ax - ay)) * d;
    } else {
        const double d = h - ax;
        t1 = 2.0 * d * (ax - 2.0 * ay);
        t2 = (4.0 * d - ay) * ay + d * d;
    }
    h -= (t1 + t2) / (2.0 * h);
    return h;
}
__device__ __forceinline__ double zabs(zc a) { return hypot_libm(a.re, a.im); }

// complex square root (C99 csqrt semantics for finite arguments)
__device__ zc zsqrt(zc z)
{
    double re = z.re, im = z.im;
    if (im == 0.0) {
        if (re < 0.0) return zmk(0.0, copysign(sqrt(-re), im));
        return zmk(fabs(sqrt(re)), copysign(0.0, im));
    }
    if (re == 0.0) {
        double r = sqrt(0.5 * fabs(im));
        return zmk(r, copysign(r, im));
    }
    double d = hypot_libm(re, im), r, s;
    if (re > 0) {
        r = sqrt(0.5 * (d + re));
        s = 0.5 * (im / r);
    } else {
        s = sqrt(0.5 * (d - re));
        r = fabs(0.5 * (im / s));
    }
    return zmk(r, copysign(s, im));
}

// LAPACK DLADIV (robust complex division)
__device__ __forceinline__ double dladiv2(double a, double b, double c, double d, double r, double t)
{
    if (r != 0.0) {
        double br = b * r;
        if (br != 0.0) return (a + br) * t;
        return a * t + (b * t) * r;
    }
    return (a + d * (b / c)) * t;
}
__device__ __forceinline__ void dladiv1(double a, double b, double c, double d, double &p, double &q)
{
    double r = d / c;
    double t = 1.0 / (c + d * r);
    p = dladiv2(a, b, c, d, r, t);
    a = -a;
    q = dladiv2(b, a, c, d, r, t);
}
__device__ zc zladiv(zc x, zc y)
{
    double aa = x.re, bb = x.im, cc = y.re, dd = y.im;
    double ab = fmax(fabs(aa), fabs(bb)), cd = fmax(fabs(cc), fabs(dd)), s = 1.0;
    const double ov = 1.79769313486231570815e+308, un = 2.22507385850720138309e-308, eps = 1.11022302462515654042e-16;
    const double bs = 2.0, be = bs / (eps * eps);
    if (ab >= 0.5 * ov) { aa *= 0.5; bb *= 0.5; s *= 2.0; }
    if (cd >= 0.5 * ov) { cc *= 0.5; dd *= 0.5; s *= 0.5; }
    if (ab <= un * bs / eps) { aa *= be; bb *= be; s /= be; }
    if (cd <= un * bs / eps) { cc *= be; dd *= be; s *= be; }
    double p, q;
    if (fabs(dd) <= fabs(cc)) dladiv1(aa, bb, cc, dd, p, q);
    else { dladiv1(bb, aa, dd, cc, p, q); q = -q; }
    return zmk(p * s, q * s);
}

__device__ __forceinline__ double dlapy3(double x, double y, double z)
{
    double xa = fabs(x), ya = fabs(y), za = fabs(z);
    double w = fmax(xa, fmax(ya, za));
    if (w == 0.0) return xa + ya + za;
    return w * sqrt((xa / w) * (xa / w) + (ya / w) * (ya / w) + (za / w) * (za / w));
}

// ZLARFG with n = 2
__device__ zc zlarfg2(zc &alpha, zc &x)
{
    double xnorm = zabs(x);
    double alphr = alpha.re, alphi = alpha.im;
    if (xnorm == 0.0 && alphi == 0.0) return zmk(0.0, 0.0);
    double beta = -copysign(dlapy3(alphr, alphi, xnorm), alphr);
    zc tau = zmk((beta - alphr) / beta, -alphi / beta);
    zc sc = zladiv(zmk(1.0, 0.0), zsub(alpha, zmk(beta, 0.0)));
    x = zmul(sc, x);
    alpha = zmk(beta, 0.0);
    return tau;
}

// 3x3 complex matrix kept in registers/local memory; H(a,b) with a,b in 0..2
struct Mat3 {
    zc e[9];
    __device__ __forceinline__ zc &operator()(int a, int b) { return e[a * 3 + b]; }
};

// ZGEBAL('B') scaling loop (no isolated rows/columns for an irreducible companion matrix)
__device__ void zgebal_scale(int n, Mat3 &h)
{
    const double radix = 2.0, factor = 0.95;
    const double sfmin1 = 2.22507385850720138309e-308 / 2.22044604925031308085e-16, sfmax1 = 1.0 / sfmin1;
    const double sfmin2 = sfmin1 * radix, sfmax2 = 1.0 / sfmin2;
    double scale[3] = {1.0, 1.0, 1.0};
    bool noconv = true;
    while (noconv) {
        noconv = false;
        for (int i = 0; i < n; ++i) {
            double c2 = 0, r2 = 0, ca = 0, ra = 0;
            for (int j = 0; j < n; ++j) {
                zc cji = h(j, i), cij = h(i, j);
                c2 += cji.re * cji.re + cji.im * cji.im;
                r2 += cij.re * cij.re + cij.im * cij.im;
                ca = fmax(ca, zabs1(cji));
                ra = fmax(ra, zabs1(cij));
            }
            double c = sqrt(c2), r = sqrt(r2);
            if (c == 0.0 || r == 0.0) continue;
            double g = r / radix, f = 1.0, s = c + r;
            while (c < g && fmax(f, fmax(c, ca)) < sfmax2 && fmin(r, fmin(g, ra)) > sfmin2) {
                f *= radix; c *= radix; ca *= radix; r /= radix; g /= radix; ra /= radix;
            }
            g = c / radix;
            while (g >= r && fmax(r, ra) < sfmax2 && fmin(fmin(f, c), fmin(g, ca)) > sfmin2) {
                f /= radix; c /= radix; g /= radix; ca /= radix; r *= radix; ra *= radix;
            }
            if ((c + r) >= factor * s) continue;
            if (f < 1.0 && scale[i] < 1.0 && f * scale[i] <= sfmin1) continue;
            if (f > 1.0 && scale[i] > 1.0 && scale[i] >= sfmax1 / f) continue;
            g = 1.0 / f;
            scale[i] *= f;
            noconv = true;
            for (int j = 0; j < n; ++j) h(i, j) = zrmul(g, h(i, j));
            for (int j = 0; j < n; ++j) h(j, i) = zrmul(f, h(j, i));
        }
    }
}

// ZLAHQR, eigenvalues only, on the whole n x n Hessenberg matrix (0-based indices)
__device__ int zlahqr3(int n, Mat3 &h, zc *w)
{
    const double safmin = 2.22507385850720138309e-308, ulp = 2.22044604925031308085e-16;
    const double dat1 = 3.0 / 4.0;
    const int kexsh = 10;
    if (n == 0) return 0;
    if (n == 1) { w[0] = h(0, 0); return 0; }
    const int ilo = 0, ihi = n - 1;
    if (ilo <= ihi - 2) h(ihi, ihi - 2) = zmk(0, 0);
    for (int i = ilo + 1; i <= ihi; ++i) {
        if (h(i, i - 1).im != 0.0) {
            zc sc = zrdiv(h(i, i - 1), zabs1(h(i, i - 1)));
            sc = zrdiv(zconj(sc), zabs(sc));
            h(i, i - 1) = zmk(zabs(h(i, i - 1)), 0.0);
            for (int j = i; j <= ihi; ++j) h(i, j) = zmul(sc, h(i, j));
            int top = min(ihi, i + 1);
            for (int j = ilo; j <= top; ++j) h(j, i) = zmul(zconj(sc), h(j, i));
        }
    }
    const int nh = ihi - ilo + 1;
    const double smlnum = safmin * ((double)nh / ulp);
    const int itmax = 30 * max(10, nh);
    int kdefl = 0;
    int i = ihi;
    while (i >= ilo) {
        int l = ilo;
        bool converged = false;
        for (int its = 0; its <= itmax; ++its) {
            int k;
            for (k = i; k >= l + 1; --k) {
                if (zabs1(h(k, k - 1)) <= smlnum) break;
                double tst = zabs1(h(k - 1, k - 1)) + zabs1(h(k, k));
                if (tst == 0.0) {
                    if (k - 2 >= ilo) tst += fabs(h(k - 1, k - 2).re);
                    if (k + 1 <= ihi) tst += fabs(h(k + 1, k).re);
                }
                if (fabs(h(k, k - 1).re) <= ulp * tst) {
                    double a1 = zabs1(h(k, k - 1)), a2 = zabs1(h(k - 1, k));
                    double ab = fmax(a1, a2), ba = fmin(a1, a2);
                    double d1 = zabs1(h(k, k)), d2 = zabs1(zsub(h(k - 1, k - 1), h(k, k)));
                    double aa = fmax(d1, d2), bb = fmin(d1, d2);
                    double s = aa + ab;
                    if (ba * (ab / s) <= fmax(smlnum, ulp * (bb * (aa / s)))) break;
                }
            }
            l = k;
            if (l > ilo) h(l, l - 1) = zmk(0, 0);
            if (l >= i) { converged = true; break; }
            ++kdefl;
            const int i1 = l, i2 = i;
            zc t;
            if (kdefl % (2 * kexsh) == 0) {
                double s = dat1 * fabs(h(i, i - 1).re);
                t = zadd(zmk(s, 0), h(i, i));
            } else if (kdefl % kexsh == 0) {
                double s = dat1 * fabs(h(l + 1, l).re);
                t = zadd(zmk(s, 0), h(l, l));
            } else {
                t = h(i, i);
                zc u = zmul(zsqrt(h(i - 1, i)), zsqrt(h(i, i - 1)));
                double s = zabs1(u);
                if (s != 0.0) {
                    zc x = zrmul(0.5, zsub(h(i - 1, i - 1), t));
                    double sx = zabs1(x);
                    s = fmax(s, zabs1(x));
                    zc xs = zrdiv(x, s), us = zrdiv(u, s);
                    zc y = zrmul(s, zsqrt(zadd(zmul(xs, xs), zmul(us, us))));
                    if (sx > 0.0) {
                        zc xn = zrdiv(x, sx);
                        if (xn.re * y.re + xn.im * y.im < 0.0) y = zneg(y);
                    }
                    t = zsub(t, zmul(u, zladiv(u, zadd(x, y))));
                }
            }
            int m;
            zc v0, v1;
            for (m = i - 1; m >= l + 1; --m) {
                zc h11 = h(m, m), h22 = h(m + 1, m + 1);
                zc h11s = zsub(h11, t);
                double h21 = h(m + 1, m).re;
                double s = zabs1(h11s) + fabs(h21);
                h11s = zrdiv(h11s, s);
                h21 = h21 / s;
                v0 = h11s;
                v1 = zmk(h21, 0);
                double h10 = h(m, m - 1).re;
                if (fabs(h10) * fabs(h21) <= ulp * (zabs1(h11s) * (zabs1(h11) + zabs1(h22)))) break;
            }
            if (m < l + 1) {
                m = l;
                zc h11s = zsub(h(l, l), t);
                double h21 = h(l + 1, l).re;
                double s = zabs1(h11s) + fabs(h21);
                v0 = zrdiv(h11s, s);
                v1 = zmk(h21 / s, 0);
            }
            for (k = m; k <= i - 1; ++k) {
                if (k > m) { v0 = h(k, k - 1); v1 = h(k + 1, k - 1); }
                zc t1 = zlarfg2(v0, v1);
                if (k > m) { h(k, k - 1) = v0; h(k + 1, k - 1) = zmk(0, 0); }
                zc v2 = v1;
                double t2 = zmul(t1, v2).re;
                for (int j = k; j <= i2; ++j) {
                    zc sum = zadd(zmul(zconj(t1), h(k, j)), zrmul(t2, h(k + 1, j)));
                    h(k, j) = zsub(h(k, j), sum);
                    h(k + 1, j) = zsub(h(k + 1, j), zmul(sum, v2));
                }
                int top = min(k + 2, i);
                for (int j = i1; j <= top; ++j) {
                    zc sum = zadd(zmul(t1, h(j, k)), zrmul(t2, h(j, k + 1)));
                    h(j, k) = zsub(h(j, k), sum);
                    h(j, k + 1) = zsub(h(j, k + 1), zmul(sum, zconj(v2)));
                }
                if (k == m && m > l) {
                    zc temp = zsub(zmk(1, 0), t1);
                    temp = zrdiv(temp, zabs(temp));
                    h(m + 1, m) = zmul(h(m + 1, m), zconj(temp));
                    if (m + 2 <= i) h(m + 2, m + 1) = zmul(h(m + 2, m + 1), temp);
                    for (int j = m; j <= i; ++j) {
                        if (j != m + 1) {
                            for (int jj = j + 1; jj <= i2; ++jj) h(j, jj) = zmul(temp, h(j, jj));
                            for (int jj = i1; jj < j; ++jj) h(jj, j) = zmul(zconj(temp), h(jj, j));
                        }
                    }
                }
            }
            zc temp = h(i, i - 1);
            if (temp.im != 0.0) {
                double rtemp = zabs(temp);
                h(i, i - 1) = zmk(rtemp, 0);
                temp = zrdiv(temp, rtemp);
                for (int jj = i + 1; jj <= i2; ++jj) h(i, jj) = zmul(zconj(temp), h(i, jj));
                for (int jj = i1; jj < i; ++jj) h(jj, i) = zmul(temp, h(jj, i));
            }
        }
        if (!converged) return i + 1;
        w[i] = h(i, i);
        kdefl = 0;
        i = l - 1;
    }
    return 0;
}

// np.roots of a real polynomial, descending coefficients p[0..deg], p[0] != 0, deg <= 3
__device__ int poly_roots(const double *p, int deg, double *wr, double *wi)
{
    int n = deg, tz = 0;
    while (n > 0 && p[n] == 0.0) { --n; ++tz; }  // trailing zeros -> zero roots appended at the end
    Mat3 h;
    zc w[3];
    for (int a = 0; a < 9; ++a) h.e[a] = zmk(0, 0);
    for (int j = 0; j < n; ++j) {  // first row -p[1:] / p[0] as numba's complex division evaluates it
        double ar = -p[j + 1], ai = -0.0, ratio = 0.0 / p[0], denom = p[0] + 0.0 * ratio;
        h(0, j) = zmk((ar + ai * ratio) / denom, (ai - ar * ratio) / denom);
    }
    for (int j = 1; j < n; ++j) h(j, j - 1) = zmk(1.0, 0.0);
    if (n >= 1) {
        if (n > 1) zgebal_scale(n, h);
        zlahqr3(n, h, w);
    }
    int cnt = 0;
    for (int i = 0; i < n; ++i) { wr[cnt] = w[i].re; wi[cnt] = w[i].im; ++cnt; }
    for (int i = 0; i < tz; ++i) { wr[cnt] = 0.0; wi[cnt] = 0.0; ++cnt; }
    return cnt;
}

// bs.py:942-982
__device__ void change_roots_order(double *mwn, int deg)
{
    double t;
#define RWRT_SWAP(a, b) (t = mwn[a], mwn[a] = mwn[b], mwn[b] = t)
    if (deg == 3) {
        int best = 1;
        for (int i = 1; i < 3; ++i)
            if (mwn[i] >= 0. && mwn[i] < mwn[best]) { RWRT_SWAP(i, best); best = i; }
        if (mwn[0] < 0) RWRT_SWAP(0, 1);
        if ((mwn[1] < 0 && mwn[2] < 0 && mwn[1] < mwn[2]) || (mwn[1] > 0 && mwn[2] < 0.)) RWRT_SWAP(1, 2);
    } else if (deg == 2) {
        if (!(mwn[0] == mwn[0] && mwn[0] > 0)) RWRT_SWAP(0, 1);  // the loop breaks on its first iteration
    } else if (deg == 1) {
        for (int i = 0; i < 3; ++i) {
            bool fin = (mwn[i] == mwn[i]);
            if (fin && mwn[i] >= 0 && i != 0) RWRT_SWAP(i, 0);
            else if (fin && mwn[i] <= 0. && i != 2) RWRT_SWAP(i, 1);
        }
    }
#undef RWRT_SWAP
    for (int i = 0; i < 3; ++i)
        if (mwn[i] == mwn[i] && fabs(mwn[i]) > 100.) mwn[i] = qnan();
    t = mwn[0]; mwn[0] = mwn[2]; mwn[2] = t;  // return mwn[::-1]
}

// k^3 rounded once.  The reference evaluates zwn**3 through libm pow (bs.py:1006), which is correctly rounded in all
// but the rarest cases (< 0.52 ulp); CUDA's pow() is specified to 2 ulp, which for a fractional zonal wavenumber could
// move the leading cubic coefficient by an ulp and with it near-degenerate roots across the |Im| < 1e-8 filter.
// Exact product in double-double (error-free transformations through fma), one final rounding.
__device__ __forceinline__ double cube_rn(double k)
{
    const double hi = k * k, lo = fma(k, k, -hi);       // k^2 = hi + lo exactly
    const double qh = hi * k, ql = fma(hi, k, -qh);     // hi k = qh + ql exactly
    return qh + fma(lo, k, ql);
}

// bs.py:985-1040 for one (source, k)
__device__ void cal_ky_point(double fu, double fv, double fqx, double fqy, double freq, double k, double *l)
{
    l[0] = l[1] = l[2] = qnan();
    if (k == 0) return;
    double ps = freq / k * REARTH;
    double coeff[4];
    coeff[0] = cube_rn(k) * (fu - ps - (fqy / (k * k)));
    coeff[1] = (k * k) * fv + fqx;
    coeff[2] = k * (fu - ps);
    coeff[3] = fv;
    int deg = 3;
    while (deg > 0 && fabs(coeff[deg]) == 0) --deg;
    if (deg < 1) return;
    double p[4];
    for (int i = 0; i <= deg; ++i) p[i] = coeff[deg - i];
    double wr[3], wi[3];
    int nr = poly_roots(p, deg, wr, wi);
    double mwn[3] = {qnan(), qnan(), qnan()};
    int cnt = 0;
    for (int i = 0; i < nr; ++i)
        if (fabs(wi[i]) < DELT) {
            if (cnt < 3) mwn[cnt] = wr[i];
            ++cnt;
        }
    change_roots_order(mwn, cnt);
    l[0] = mwn[0];
    l[1] = mwn[1];
    l[2] = mwn[2];
}

__global__ void __launch_bounds__(128) ray_init_kernel(Grid g, int nsource, const double *src_lon, const double *src_lat,
                                                       const int32_t *src_member, int nzwn, const double *zwn,
                                                       double freq, int nt, double *out)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nsource * nzwn) return;
    int s = (int)(idx / nzwn), iz = (int)(idx % nzwn);
    const double *fb = g.fpk + (src_member ? (size_t)src_member[s] * g.member_stride : 0);
    double lon = src_lon[s], lat = src_lat[s], k = zwn[iz];
    Merc m = sample_mercator<false>(g, fb, lon, lat);  // wr.py:360-365
    double l[3];
    cal_ky_point(m.u, m.v, m.qx, m.qy, freq, k, l);
    const size_t N = (size_t)3 * nsource * nzwn, pl = (size_t)nt * N;
    for (int ir = 0; ir < 3; ++ir) {
        size_t o = ((size_t)ir * nsource + s) * nzwn + iz;
        double ug, vg;
        ugvg_numpy(m.u, m.v, m.qx, m.qy, k, l[ir], ug, vg);
        out[0 * pl + o] = lon;                           // wr.py:354-357
        out[1 * pl + o] = lat;
        out[2 * pl + o] = k;                             // wr.py:368-369
        out[3 * pl + o] = l[ir];                         // wr.py:385
        out[4 * pl + o] = (l[ir] == l[ir]) ? 1.0 : qnan();  // wr.py:388-390
        out[5 * pl + o] = ug;                            // wr.py:393-395
        out[6 * pl + o] = vg;
    }
}

__global__ void __launch_bounds__(128) sample_kernel(Grid g, long long n, const double *lon, const double *lat,
                                                     const int32_t *member, double *out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *fb = g.fpk + (member ? (size_t)member[i] * g.member_stride : 0);
    Merc m = sample_mercator<false>(g, fb, lon[i], lat[i]);
    out[0 * n + i] = m.u;
    out[1 * n + i] = m.v;
    out[2 * n + i] = m.ux;
    out[3 * n + i] = m.uy;
    out[4 * n + i] = m.vx;
    out[5 * n + i] = m.vy;
    out[6 * n + i] = m.qx;
    out[7 * n + i] = m.qy;
    out[8 * n + i] = m.qxx;
    out[9 * n + i] = m.qxy;
    out[10 * n + i] = m.qyx;
    out[11 * n + i] = m.qyy;
}

// All 18 rows of BS.cal_bs_mercator_point(mode='numpy') (bs.py:781-887) from the reference-layout fields18
// table (nlon_f, nlat, 18): rows 12..17 are the third-derivative terms the ray path never consumes.
__global__ void __launch_bounds__(128) sample18_kernel(Grid g, const double *f18, size_t member_stride18, long long n,
                                                       const double *lon, const double *lat, const int32_t *member,
                                                       double *out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double la = lat[i];
    if (!(fabs(la) <= HALFPI)) {  // bs.py:787
        for (int k = 0; k < 18; ++k) out[(size_t)k * n + i] = qnan();
        return;
    }
    const double *fb = f18 + (member ? (size_t)member[i] * member_stride18 : 0);
    int x0, y0, x1, y1;
    double sx, sy;
    locate<false>(g, lon[i], la, x0, y0, x1, y1, sx, sy);
    const int H = g.nlat;
    double wa = (1 - sx) * sy, wb = sx * sy, wc = (1 - sx) * (1 - sy), wd = sx * (1 - sy);
    const double *pa = fb + ((size_t)x0 * H + y1) * 18, *pb = fb + ((size_t)x1 * H + y1) * 18;
    const double *pc = fb + ((size_t)x0 * H + y0) * 18, *pd = fb + ((size_t)x1 * H + y0) * 18;
    double f[18];
    for (int k = 0; k < 18; ++k) f[k] = pa[k] * wa + pb[k] * wb + pc[k] * wc + pd[k] * wd;
    double cp = cos(la), sp = sin(la), tl = tan(la);
    double mask = (fabs(cp) <= 0.0175) ? 0.0 : 1.0;  // bs.py:858-860
    cp = cp * mask + (1 - mask) * 1e-6;
    double o[18];
    o[8] = f[8] * mask;                               // fmqxx
    o[10] = f[9] * cp * mask;                         // fmqyx
    o[9] = o[10] * mask;                              // fmqxy
    o[11] = (f[11] * cp - f[7] * sp) * cp * mask;     // fmqyy
    o[12] = f[12] * mask;                             // fmqxxx
    o[13] = f[13] * cp * mask;                        // fmqxxy
    o[14] = (f[14] * cp - f[9] * sp) * cp * mask;     // fmqxyy
    o[15] = f[15] * mask;                             // fmqyyy
    o[16] = f[16] * cp * mask;                        // fmqyxx
    o[17] = (f[17] * cp - f[9] * sp) * cp * mask;     // fmqyyx
    o[6] = f[6] * mask;                               // fmqx
    o[7] = f[7] * cp * mask;                          // fmqy
    o[2] = f[2] / cp * mask;                          // fmux
    o[4] = f[4] / cp * mask;                          // fmvx
    o[3] = (f[3] + tl * f[0]) * mask;                 // fmuy
    o[5] = (f[5] + tl * f[1]) * mask;                 // fmvy
    o[0] = f[0] / cp * mask;                          // fmu
    o[1] = f[1] / cp * mask;                          // fmv
    for (int k = 0; k < 18; ++k) out[(size_t)k * n + i] = o[k];
}

}  // namespace

void launch_sample18(const Grid &g, const double *f18, long long n, const double *lon, const double *lat,
                     const int32_t *member, double *out, cudaStream_t s)
{
    if (n > 0)
        sample18_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g, f18, (size_t)g.nlon_f * g.nlat * 18, n, lon, lat,
                                                                    member, out);
}

void launch_ray_init(const Grid &g, int nsource, const double *src_lon, const double *src_lat, const int32_t *src_member,
                     int nzwn, const double *zwn, double freq, int nt, double *out7, cudaStream_t s)
{
    long long n = (long long)nsource * nzwn;
    if (n > 0)
        ray_init_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g, nsource, src_lon, src_lat, src_member, nzwn, zwn,
                                                                   freq, nt, out7);
}

void launch_sample_mercator(const Grid &g, long long n, const double *lon, const double *lat, const int32_t *member,
                            double *out, cudaStream_t s)
{
    if (n > 0) sample_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g, n, lon, lat, member, out);
}

}  // namespace rwrt
