// rwrt_device.cuh -- device functions shared by the sampler, ray-init and RK45 kernels.
//
// Reference semantics being reproduced (file:line into yinan-codes/Rossby-wave-ray-tracing):
//   interpolation.py:77-85, 103-135   index computation + bilinear blend (clamped indices)
//   bs.py:513-519, 781-887            lon wrap, Mercator transform, |cos(lat)| <= 0.0175 mask
//   wn.py:266-294                     group velocity ("extent" form)
//   wr.py:44-82, 492-556              ray equations + |lat| >= pi/2, |l| >= 100 masks
//
// Layout: the 11 fields the ray equations consume (of the reference's 18) are packed per grid
// point as 12 doubles = 96 B (u,v,ux,uy,vx,vy,qx,qy,qxx,qxy,qyy,pad), points ordered [i_lon][j_lat]
// so the two latitude corners of a cell are 192 contiguous bytes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rwrt {

constexpr double PI = 3.14159265358979323846264338327950288419716939937510;  // constants.py:13
constexpr double TWOPI = 2 * PI;
constexpr double HALFPI = 0.5 * PI;
constexpr double REARTH = 6.3712e6;                                           // constants.py:16
constexpr double OMEGA = 7.2921e-5;                                           // constants.py:17
constexpr double DELT = 1.0e-8;                                               // constants.py:28
constexpr int NPK = 12;  // doubles per packed grid point

struct Grid {
    const double *fpk;     // (nmember, nlon_f, nlat, NPK)
    size_t member_stride;  // doubles between members
    int nlon_f, nlat;      // nlon_f = nlon + 1 when xcyclic (wrap column, bs.py:370-372)
    double lon0, dlon, lat0, dlat;  // lon[0], lon[1]-lon[0], lat[0], lat[1]-lat[0] (interpolation.py:78-79)
    double rdlon, rdlat;            // 1/dlon, 1/dlat (fast build only)
};

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// Python / numpy float remainder by 2*pi (result has the divisor's sign).  fmod is exact, and so is this
// FMA form of it: for the right integer quotient q the remainder a - q*2pi is representable, so one
// fused multiply-add returns it without rounding; a quotient that is off by one (rounding of a/2pi)
// shows up as a remainder outside [0, 2pi) and is corrected.  Bit-identical to fmod(), without its loop.
__device__ __forceinline__ double pymod_2pi(double a)
{
    double aa = fabs(a);
    double r;
    if (aa < 1.0e9) {
        double q = floor(aa * (1.0 / TWOPI));
        r = fma(-q, TWOPI, aa);
        if (r < 0.0) r = fma(-(q - 1.0), TWOPI, aa);
        else if (r >= TWOPI) r = fma(-(q + 1.0), TWOPI, aa);
    } else {
        r = fmod(aa, TWOPI);  // huge, inf or NaN arguments (kept inline: out of line it costs 0.3 %, the call's register constraints)
    }
    if (a < 0.0 && r != 0.0) r = TWOPI - r;  // np.remainder: fmod(a, b) + b when the signs differ
    return r;
}

#ifndef RWRT_LON_DIRECT
#define RWRT_LON_DIRECT 1
#endif

#ifndef RWRT_PYMOD_NOINLINE
#define RWRT_PYMOD_NOINLINE 1        /* 86 instructions (fmod's loop) leave the stage loop's body: 38.09 -> 37.84 ms on C2 */
#endif
// the general routine out of line, for the one-in-10^6 longitudes that leave locate_cell's direct path
static __device__ __noinline__ double pymod_2pi_wrapped(double a)
{
    double r = pymod_2pi(a);
    return r >= TWOPI ? 0.0 : r;
}

// np.maximum: NaN if either operand is NaN
__device__ __forceinline__ double np_maximum(double a, double b) { return (a != a || b != b) ? qnan() : fmax(a, b); }

// np.maximum(|a|, |b|) (the error scale, rkf45.py:443): the bit patterns of non-negative doubles order like integers and
// every NaN pattern lies above +inf, so the integer maximum of the patterns IS the NaN-propagating maximum -- two integer
// compares and two selects where fmax on doubles is emulated with ~9 instructions per component on sm_100.
__device__ __forceinline__ double np_absmax(double a, double b)
{
    const long long ia = __double_as_longlong(a) & 0x7fffffffffffffffLL, ib = __double_as_longlong(b) & 0x7fffffffffffffffLL;
    return __longlong_as_double(ia > ib ? ia : ib);
}

// x / sqrt(5) (rkf45.py:29-31), correctly rounded without the division routine (~18 instructions): with c = RN(sqrt 5),
// y = RN(1 / c), q = RN(x y) and the exact remainder r = x - q c, RN(q + r y) is the correctly rounded quotient
// (Markstein's final division step; checked against x / c on 2e5 random arguments with exact rational arithmetic).
// inf is passed through (the remainder would be NaN); NaN stays NaN.
__device__ __forceinline__ double div_sqrt5(double x)
{
    const double c = 2.23606797749978969641, y = 0.44721359549995793928;   // 0x1.1e3779b97f4a8p+1, 0x1.c9f25c5bfedd9p-2
    const double q = x * y;
    const double r = fma(-q, c, x);
    const double res = fma(r, y, q);
    return (fabs(x) == __longlong_as_double(0x7ff0000000000000LL)) ? x : res;
}

// Bilinear sample of the packed fields (interpolation.py:103-135).  f[0..10].
// interpolation.py:77-85 + 112-130: fractional indices, floor -> int32, clamp, weights from the CLAMPED indices
template <bool FAST>
__device__ __forceinline__ void locate(const Grid &g, double lon, double lat, int &x0, int &y0, int &x1, int &y1,
                                       double &sx, double &sy)
{
#if RWRT_LON_DIRECT
    // bs.py:519 + interpolation.py:80 on the SIGNED longitude: with q = floor(lon / 2pi) the remainder lon - q 2pi is
    // what np.remainder returns, for either sign, in ONE rounding: for lon >= 0 it is fmod's exact result; for lon < 0
    // NumPy computes 2pi - fmod(|lon|, 2pi) = lon + (m + 1) 2pi (m = floor(|lon| / 2pi)), i.e. the same real number
    // rounded once, and q = -(m + 1).  A quotient that is off by one, an exact multiple of 2pi, a result that rounds
    // to 2pi itself, huge / inf / NaN arguments all leave [0, 2pi) and take the general routine (NaN / inf give NaN either way and skip it).  7 instructions
    // instead of 35 on the path that every stage of every ray walks.
    double l2 = fma(-floor(lon * (1.0 / TWOPI)), TWOPI, lon);
    if (!(l2 >= 0.0 && l2 < TWOPI) && l2 == l2) {                          // (NaN: lon is NaN or +-inf, and so is the remainder)
        l2 = pymod_2pi(lon);
        if (l2 >= TWOPI) l2 = 0.0;
    }
#else
    double l1 = pymod_2pi(lon);            // bs.py:519
    double l2 = (l1 >= TWOPI) ? 0.0 : l1;  // interpolation.py:80: a second % 2pi only moves an exact 2pi to 0
#endif
    double x, y;
    if (FAST) {
        x = (l2 - g.lon0) * g.rdlon;
        y = (lat - g.lat0) * g.rdlat;
    } else {
        x = (l2 - g.lon0) / g.dlon;
        y = (lat - g.lat0) / g.dlat;
    }
    const int W = g.nlon_f, H = g.nlat;
    int x0u = __double2int_rd(x), y0u = __double2int_rd(y);  // floor -> int32 (NaN -> 0 after the clamp either way)
    x0 = min(max(x0u, 0), W - 1);
    y0 = min(max(y0u, 0), H - 1);
    x1 = min(max(x0u + 1, 0), W - 1);
    y1 = min(max(y0u + 1, 0), H - 1);
    sx = x - (double)x0;
    sy = y - (double)y0;
}

// The same for the cooperative kernels, which stage a cell by its (x0, y0) point and only need to know WHETHER the
// far index was clamped onto the near one: x1 = clamp(x0u + 1) equals x0 = clamp(x0u) exactly when the unclamped index
// lies outside [0, W - 2] -- one unsigned compare per axis instead of a second min / max chain and an equality test.
// clamp: bit 30 = x1 == x0, bit 31 = y1 == y0.
template <bool FAST, bool NANSKIP>
__device__ __forceinline__ void locate_cell(const Grid &g, double lon, double lat, int &x0, int &y0, unsigned &clamp,
                                            double &sx, double &sy)
{
#if RWRT_LON_DIRECT
    double l2 = fma(-floor(lon * (1.0 / TWOPI)), TWOPI, lon);             // see locate()
    // (a NaN here means lon is NaN or +-inf, and then the remainder is NaN as well.  NANSKIP: no call for it -- in the
    // fixed-step RK4 kernel, where the slots without a root run along as NaN, half of the lanes would take the general
    // routine in every evaluation (42.6 -> 39.6 ms per C2 pass); the adaptive kernel drops a NaN ray at once and
    // does without the third compare on its common path (36.97 -> 36.85 ms))
    if (!(l2 >= 0.0 && l2 < TWOPI) && (!NANSKIP || l2 == l2)) {
#if RWRT_PYMOD_NOINLINE
        l2 = pymod_2pi_wrapped(lon);
#else
        l2 = pymod_2pi(lon);
        if (l2 >= TWOPI) l2 = 0.0;
#endif
    }
#else
    double l1 = pymod_2pi(lon);
    double l2 = (l1 >= TWOPI) ? 0.0 : l1;
#endif
    double x, y;
    if (FAST) {
        x = (l2 - g.lon0) * g.rdlon;
        y = (lat - g.lat0) * g.rdlat;
    } else {
        x = (l2 - g.lon0) / g.dlon;
        y = (lat - g.lat0) / g.dlat;
    }
    const int W = g.nlon_f, H = g.nlat;
    const int x0u = __double2int_rd(x), y0u = __double2int_rd(y);
    x0 = min(max(x0u, 0), W - 1);
    y0 = min(max(y0u, 0), H - 1);
    clamp = ((unsigned)x0u > (unsigned)(W - 2) ? 0x40000000u : 0u) | ((unsigned)y0u > (unsigned)(H - 2) ? 0x80000000u : 0u);
    sx = x - (double)x0;
    sy = y - (double)y0;
}

// Bilinear sample of the packed fields (interpolation.py:103-135).  f[0..10].
template <bool FAST>
__device__ __forceinline__ void sample_packed(const Grid &g, const double *__restrict__ fb, double lon, double lat,
                                              double *f)
{
    int x0, y0, x1, y1;
    double sx, sy;
    locate<FAST>(g, lon, lat, x0, y0, x1, y1, sx, sy);
    const int H = g.nlat;
    double wa = (1 - sx) * sy, wb = sx * sy, wc = (1 - sx) * (1 - sy), wd = sx * (1 - sy);
    const double2 *pa = reinterpret_cast<const double2 *>(fb + ((size_t)x0 * H + y1) * NPK);
    const double2 *pb = reinterpret_cast<const double2 *>(fb + ((size_t)x1 * H + y1) * NPK);
    const double2 *pc = reinterpret_cast<const double2 *>(fb + ((size_t)x0 * H + y0) * NPK);
    const double2 *pd = reinterpret_cast<const double2 *>(fb + ((size_t)x1 * H + y0) * NPK);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double2 a = __ldg(pa + k), b = __ldg(pb + k), c = __ldg(pc + k), d = __ldg(pd + k);
        f[2 * k] = a.x * wa + b.x * wb + c.x * wc + d.x * wd;
        if (2 * k + 1 < 11) f[2 * k + 1] = a.y * wa + b.y * wb + c.y * wc + d.y * wd;
    }
}

// Mercator quantities (bs.py:856-883).  Only what the ray path consumes.
struct Merc {
    double u, v, ux, uy, vx, vy, qx, qy, qxx, qxy, qyx, qyy;
    double cosl, sinl;
};

// FAST-build math: branch-free sin/cos for |x| <= pi/2 (one quadrant fold by Cody-Waite, then the fdlibm
// kernel polynomials on [-pi/4, pi/4]; < 1 ulp), with the coefficients as constant-bank operands.  The
// library sincos costs 12 % of the integrator's instructions (slow-path branch, coefficient table loads).
__constant__ double cSinCos[12] = {
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};

__device__ __forceinline__ void sincos_halfpi(double x, double &sn, double &cs)
{
    const double q = rint(x * 0.63661977236758134308);            // -1, 0 or 1 for |x| <= pi/2
    double r = fma(-q, 1.57079632679489655800e+00, x);
    r = fma(-q, 6.12323399573676603587e-17, r);
    const double z = r * r;
    double ps = fma(z, cSinCos[5], cSinCos[4]);
    double pc = fma(z, cSinCos[11], cSinCos[10]);
    ps = fma(z, ps, cSinCos[3]);
    pc = fma(z, pc, cSinCos[9]);
    ps = fma(z, ps, cSinCos[2]);
    pc = fma(z, pc, cSinCos[8]);
    ps = fma(z, ps, cSinCos[1]);
    pc = fma(z, pc, cSinCos[7]);
    ps = fma(z, ps, cSinCos[0]);
    pc = fma(z, pc, cSinCos[6]);
    const double sr = fma(r * z, ps, r);                           // sin(r)
    const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));          // cos(r)
    // quadrant fix-up without selects: q is -1, 0 or 1, so every product below is exact (x0, x1, x-1)
    const double nq2 = fma(-q, q, 1.0);                            // 1 - q^2: 1 in the centre, 0 in the folded quadrants
    sn = fma(q, cr, nq2 * sr);
    cs = fma(-q, sr, nq2 * cr);
}

// FAST-build sine of a small angle (the speed cut-off's sin(dlat/2), sin(dlon/2)): the fdlibm kernel polynomial
// on |x| <= pi/4, the library routine beyond (rare: such a ray has moved > 90 degrees in one output interval).
// Out-of-line library calls for the fast build's rare paths (a masked output point, a ray that moved > 90 degrees in
// one output interval): inlined, libdevice's sin / cos / atan2 put ~1 000 instructions between the hot parts of the
// integrator's attempt loop.
#ifndef RWRT_COLD_OOL
#define RWRT_COLD_OOL 1
#endif
#if RWRT_COLD_OOL
#define RWRT_COLD static __device__ __noinline__
#else
#define RWRT_COLD static __device__ __forceinline__
#endif
RWRT_COLD double sin_cold(double x) { return sin(x); }
RWRT_COLD double cos_cold(double x) { return cos(x); }
__device__ __forceinline__ double sin_small(double x)
{
    if (!(fabs(x) <= 0.78539816339744830962)) return sin_cold(x);
    const double z = x * x;
    double ps = fma(z, cSinCos[5], cSinCos[4]);
    ps = fma(z, ps, cSinCos[3]);
    ps = fma(z, ps, cSinCos[2]);
    ps = fma(z, ps, cSinCos[1]);
    ps = fma(z, ps, cSinCos[0]);
    return fma(x * z, ps, x);
}

// FAST-build reciprocal: hardware seed (MUFU.RCP64H, ~2^-20) + two Newton steps, no special-case branch
// (0 -> inf, like the division it replaces; the operands here are never denormal).
__device__ __forceinline__ double rcp_fast(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// FAST-build err^(-1/5) for the step controller (rkf45.py:453-455, 474-475): float seed + two Newton
// steps y <- y (6 - a y^5) / 5 in double (quadratic: 2^-21 -> 2^-42 -> full), instead of the 166-instruction pow().
__device__ __forceinline__ double pow_m02(double a)
{
    // a lies in [5e-6, 2000] (step_factor): neither the logarithm's argument nor the exponential's needs the denormal
    // / range handling of __log2f and exp2f (6 instructions); the bare MUFU results are the same bits there
    float la, ex;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(la) : "f"((float)a));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-0.2f * la));
    double y = (double)ex;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        double y2 = y * y;
        double y5 = y2 * y2 * y;
        y = y * (fma(-a, y5, 6.0) * 0.2);
    }
    return y;
}

// cos / sin / tan of the latitude as bs.py:856-857, 876-877 evaluate them (FAST: one sincos and a division)
struct Trig {
    double cp, sp, tl, rc;  // rc = 1/cos (FAST only)
};
template <bool FAST>
__device__ __forceinline__ Trig trig_of_lat(double lat)
{
    Trig t;
    t.rc = 0.0;
    if (FAST) {
        sincos_halfpi(lat, t.sp, t.cp);
        t.rc = rcp_fast(t.cp);
        t.tl = t.sp * t.rc;
    } else {
        t.cp = cos(lat);
        t.sp = sin(lat);
        t.tl = tan(lat);
    }
    return t;
}

// bs.py:856-883 applied to the 11 interpolated fields f at a latitude with |lat| <= pi/2
template <bool FAST>
__device__ __forceinline__ Merc mercator_from_fields(const double *f, const Trig &tg)
{
    Merc m;
    double fu = f[0], fv = f[1], fux = f[2], fuy = f[3], fvx = f[4], fvy = f[5], fqx = f[6], fqy = f[7], fqxx = f[8],
           fqxy = f[9], fqyy = f[10];
    const double cp = tg.cp, sp = tg.sp, tl = tg.tl;
    m.cosl = cp;
    m.sinl = sp;
    if (fabs(cp) <= 0.0175) {  // bs.py:858-883 with mask = 0: every Mercator quantity is multiplied by zero
        // (cos -> 1e-6 there, so nothing overflows and the products are exact zeros unless a field is NaN)
        double z = (fu + fv + fux + fuy + fvx + fvy + fqx + fqy + fqxx + fqxy + fqyy) * 0.0;
        m.u = m.v = m.ux = m.uy = m.vx = m.vy = m.qx = m.qy = m.qxx = m.qxy = m.qyx = m.qyy = z;
        return m;
    }
    // mask = 1: cos_phi * 1 + 0 * 1e-6 = cos_phi, and the trailing "* mask" factors are exact
    m.qxx = fqxx;
    m.qyx = fqxy * cp;
    m.qxy = m.qyx;
    m.qyy = (fqyy * cp - fqy * sp) * cp;
    m.qx = fqx;
    m.qy = fqy * cp;
    m.uy = fuy + tl * fu;
    m.vy = fvy + tl * fv;
    if (FAST) {
        const double rc = tg.rc;
        m.ux = fux * rc;
        m.vx = fvx * rc;
        m.u = fu * rc;
        m.v = fv * rc;
    } else {
        m.ux = fux / cp;
        m.vx = fvx / cp;
        m.u = fu / cp;
        m.v = fv / cp;
    }
    return m;
}

template <bool FAST>
__device__ __forceinline__ Merc sample_mercator(const Grid &g, const double *__restrict__ fb, double lon, double lat)
{
    if (!(fabs(lat) <= HALFPI)) {  // bs.py:787: rows outside +-pi/2 (or NaN) stay NaN
        Merc m;
        double n = qnan();
        m.u = m.v = m.ux = m.uy = m.vx = m.vy = m.qx = m.qy = m.qxx = m.qxy = m.qyx = m.qyy = n;
        m.cosl = n;
        m.sinl = n;
        return m;
    }
    double f[12];
    sample_packed<FAST>(g, fb, lon, lat, f);
    return mercator_from_fields<FAST>(f, trig_of_lat<FAST>(lat));
}

// wn.py:266-294
__device__ __forceinline__ void ugvg_extent(double fu, double fv, double fqx, double fqy, double k, double l, double &ug,
                                            double &vg)
{
    double kap = l / k;
    double kap2 = kap * kap;
    double kap1 = 1.0 + kap2;
    double KK = k * k * kap1;
    double den = KK * kap1;
    ug = fu + (((1. - kap2) * fqy) - (2. * kap * fqx)) / den;
    vg = fv + ((2. * kap * fqy) + ((1. - kap2) * fqx)) / den;
}

// wn.py:209-259 (t = 0 only)
__device__ __forceinline__ void ugvg_numpy(double fu, double fv, double fqx, double fqy, double k, double l, double &ug,
                                           double &vg)
{
    if (k == 0) { ug = 0; vg = 0; return; }
    double nans = (l * 0) * (fu * fqx * fqy * 0) + 1;
    if (nans != nans) nans = 0;
    double kap1 = k * k - l * l;
    double kap2 = 2 * k * l;
    double KK2 = k * k + l * l;
    double kk4 = KK2 * KK2;
    ug = (fu + (kap1 * fqy - kap2 * fqx) / kk4) * nans;
    vg = (fv + (kap1 * fqx + kap2 * fqy) / kk4) * nans;
}

// One right-hand side: wr.py:784-789 -> 492-556 -> 44-82.  y = (lon, lat, k, l, A).
// FAST: same formulas with the shared denominators inverted once (1 reciprocal instead of 7 divisions
// in the wavenumber algebra, multiplication by 1/a instead of 5 divisions) -- differences are at the
// level of one rounding per operation, like FMA contraction.
// FAST build: everything the ray equations need from (k, l) alone, so that it can be evaluated while
// the field gather is still in flight.
struct WaveRecip {
    double kap, kap2, rkk, rkap1, rden;
};
__device__ __forceinline__ WaveRecip wave_recip(double kx, double ky)
{
    WaveRecip w;
    double k2 = kx * kx, l2 = ky * ky;
    double K2 = k2 + l2;                 // = kk = kx^2 (1 + kap^2)
    double r = rcp_fast(kx * K2 * K2);   // one reciprocal for 1/kx, 1/kk, 1/kap1, 1/(kk kap1)
    double rk = r * K2 * K2;             // 1/kx
    w.kap = ky * rk;
    w.kap2 = w.kap * w.kap;
    w.rkk = r * kx * K2;                 // 1/kk
    w.rkap1 = k2 * w.rkk;                // 1/kap1 = kx^2 / kk
    w.rden = w.rkk * w.rkap1;            // 1/(kk kap1)
    return w;
}

template <bool FAST>
__device__ __forceinline__ void rhs_from_merc(const Merc &m, const double *y, const WaveRecip &w, double *dy,
                                              double &ugo, double &vgo)
{
    double kx = y[2], ky = y[3], amp = y[4];
    if (FAST) {
        const double kap = w.kap, kap2 = w.kap2, rkk = w.rkk, rkap1 = w.rkap1, rden = w.rden;
        double ug = m.u + (((1. - kap2) * m.qy) - (2. * kap * m.qx)) * rden;
        double vg = m.v + ((2. * kap * m.qy) + ((1. - kap2) * m.qx)) * rden;
        double dzwn = -kx * ((m.ux + kap * m.vx) + (kap * m.qxx - m.qyx) * rkk);
        double dmwn = -kx * ((m.uy + kap * m.vy) + (kap * m.qxy - m.qyy) * rkk);
        double damp1 = 2.0 * (m.ux + m.vy + kap * (m.vx + m.uy)) * rkap1;
        double damp2 = 2.0 * (kap * (m.qxx - m.qyy) + (kap2 - 1.0) * m.qxy) * rden;
        double damp3 = -2.0 * m.sinl * m.v;
        double damp = damp1 + damp2 + damp3;
        const double ra = 1.0 / REARTH;
        dy[0] = ug * ra;
        dy[1] = (vg * m.cosl) * ra;
        dy[2] = dzwn * ra;
        dy[3] = dmwn * ra;
        dy[4] = (damp * amp) * ra;
        ugo = ug;
        vgo = vg;
        return;
    }
    double ug, vg;
    ugvg_extent(m.u, m.v, m.qx, m.qy, kx, ky, ug, vg);
    double kap = ky / kx;
    double kap2 = kap * kap;
    double kap1 = 1 + kap * kap;
    double kk = kx * kx * kap1;
    double dzwn = -kx * ((m.ux + kap * m.vx) + (kap * m.qxx - m.qyx) / kk);
    double dmwn = -kx * ((m.uy + kap * m.vy) + (kap * m.qxy - m.qyy) / kk);
    double damp1 = 2.0 * (m.ux + m.vy + kap * (m.vx + m.uy)) / kap1;
    double damp2 = 2.0 * (kap * (m.qxx - m.qyy) + (kap2 - 1.0) * m.qxy) / (kk * kap1);
    double damp3 = -2.0 * m.sinl * m.v;
    double damp = damp1 + damp2 + damp3;
    dy[0] = ug / REARTH;
    dy[1] = (vg * m.cosl) / REARTH;
    dy[2] = dzwn / REARTH;
    dy[3] = dmwn / REARTH;
    dy[4] = (damp * amp) / REARTH;
    ugo = ug;
    vgo = vg;
}

// The same right-hand side in two pieces, for the pipelined integrator: rhs_position gives d lon/dt, d lat/dt and
// (ug, vg) -- all the next stage's position needs --, rhs_wavenumber the other three components.  Operation by
// operation identical to rhs_from_merc.
template <bool FAST>
__device__ __forceinline__ void rhs_position(const Merc &m, const double *y, const WaveRecip &w, double *dy, double &ugo,
                                             double &vgo)
{
    double ug, vg;
    if (FAST) {
        const double kap = w.kap, kap2 = w.kap2, rden = w.rden;
        ug = m.u + (((1. - kap2) * m.qy) - (2. * kap * m.qx)) * rden;
        vg = m.v + ((2. * kap * m.qy) + ((1. - kap2) * m.qx)) * rden;
        const double ra = 1.0 / REARTH;
        dy[0] = ug * ra;
        dy[1] = (vg * m.cosl) * ra;
    } else {
        ugvg_extent(m.u, m.v, m.qx, m.qy, y[2], y[3], ug, vg);
        dy[0] = ug / REARTH;
        dy[1] = (vg * m.cosl) / REARTH;
    }
    ugo = ug;
    vgo = vg;
}
template <bool FAST>
__device__ __forceinline__ void rhs_wavenumber(const Merc &m, const double *y, const WaveRecip &w, double *dy)
{
    double kx = y[2], ky = y[3], amp = y[4];
    if (FAST) {
        const double kap = w.kap, kap2 = w.kap2, rkk = w.rkk, rkap1 = w.rkap1, rden = w.rden;
        double dzwn = -kx * ((m.ux + kap * m.vx) + (kap * m.qxx - m.qyx) * rkk);
        double dmwn = -kx * ((m.uy + kap * m.vy) + (kap * m.qxy - m.qyy) * rkk);
        double damp1 = 2.0 * (m.ux + m.vy + kap * (m.vx + m.uy)) * rkap1;
        double damp2 = 2.0 * (kap * (m.qxx - m.qyy) + (kap2 - 1.0) * m.qxy) * rden;
        double damp3 = -2.0 * m.sinl * m.v;
        double damp = damp1 + damp2 + damp3;
        const double ra = 1.0 / REARTH;
        dy[2] = dzwn * ra;
        dy[3] = dmwn * ra;
        dy[4] = (damp * amp) * ra;
        return;
    }
    double kap = ky / kx;
    double kap2 = kap * kap;
    double kap1 = 1 + kap * kap;
    double kk = kx * kx * kap1;
    double dzwn = -kx * ((m.ux + kap * m.vx) + (kap * m.qxx - m.qyx) / kk);
    double dmwn = -kx * ((m.uy + kap * m.vy) + (kap * m.qxy - m.qyy) / kk);
    double damp1 = 2.0 * (m.ux + m.vy + kap * (m.vx + m.uy)) / kap1;
    double damp2 = 2.0 * (kap * (m.qxx - m.qyy) + (kap2 - 1.0) * m.qxy) / (kk * kap1);
    double damp3 = -2.0 * m.sinl * m.v;
    double damp = damp1 + damp2 + damp3;
    dy[2] = dzwn / REARTH;
    dy[3] = dmwn / REARTH;
    dy[4] = (damp * amp) / REARTH;
}

// ugo / vgo: the group velocity at y (m/s), which is also what the reference stores at output times
// (wr.py:856-865 re-samples at the same point with the same formula).  masked is set when the
// |lat| / |l| mask fired (then ugo / vgo are not computed here).
template <bool FAST>
__device__ __forceinline__ void ray_rhs(const Grid &g, const double *__restrict__ fb, const double *y, double *dy,
                                        double &ugo, double &vgo, bool &masked)
{
    double lon = y[0], lat = y[1], ky = y[3];
    bool err = (fabs(lat) >= HALFPI) || (fabs(ky) >= 100.0);  // wr.py:508-510
    masked = err;
    if (err) {
        double n = qnan();
        dy[0] = dy[1] = dy[2] = dy[3] = dy[4] = n;  // wr.py:552-553
        ugo = vgo = n;
        return;
    }
    Merc m = sample_mercator<FAST>(g, fb, lon, lat);
    WaveRecip w;
    if (FAST) w = wave_recip(y[2], ky);
    rhs_from_merc<FAST>(m, y, w, dy, ugo, vgo);
}

template <bool FAST>
__device__ __forceinline__ void ray_rhs(const Grid &g, const double *__restrict__ fb, const double *y, double *dy)
{
    double ug, vg;
    bool masked;
    ray_rhs<FAST>(g, fb, y, dy, ug, vg, masked);
}

// wr.py:97-112
__device__ __forceinline__ double haversine(double lon_c, double lat_c, double lon_p, double lat_p)
{
    double dlon = lon_c - lon_p, dlat = lat_c - lat_p;
    double s1 = sin(dlat / 2.0), s2 = sin(dlon / 2.0);
    double a = s1 * s1 + cos(lat_p) * cos(lat_c) * (s2 * s2);
    double c = 2 * atan2(sqrt(a), sqrt(1.0 - a));
    return fabs(c);
}


// Ray-density / wave-ray-flux gridding of one output sample (this project's definition, oracle/gridding.py; the
// reference ships no gridding code -- the manual's tool would read rlon, rlat, rug, rvg as wr.py:947-959 writes
// them).  Every operation is an individually rounded IEEE op (intrinsics: no FMA contraction in any build), so the
// cell index equals NumPy's floor((lon * rad2deg + 360) / dlon) bit for bit wherever this is inlined.
// Returns the flattened cell ix * ncy + iy, or -1 when the sample is not counted.
__device__ __forceinline__ long long grid_cell(double lon, double lat, double l, double dlon, double dlat, int ncx,
                                               int ncy)
{
    if (!(lon == lon) || !(lat == lat) || !(l == l)) return -1;
    const double rad2deg = 1.0 / (PI / 180.0);  // constants.py:14-15
    const double x = __ddiv_rn(__dadd_rn(__dmul_rn(lon, rad2deg), 360.0), dlon);
    const double y = __ddiv_rn(__dadd_rn(__dmul_rn(lat, rad2deg), 90.0), dlat);
    const double fxd = floor(x), fyd = floor(y);
    if (!(fxd >= 0.0) || !(fxd < (double)ncx) || !(fyd >= 0.0)) return -1;
    const int ix = (int)fxd, iy = (int)fmin(fyd, (double)(ncy - 1));
    return (long long)ix * ncy + iy;
}
// count += 1; flux += rint(ug|vg * 2^20) as int64 fixed point: integer sums are order independent, so the grids are
// bit-identical for any sharding of the rays over GPUs and any accumulation order.  Results unused: RED, not ATOM.
__device__ __forceinline__ void grid_add(unsigned long long *count, unsigned long long *fx, unsigned long long *fy,
                                         long long c, double ug, double vg)
{
    atomicAdd(count + c, 1ULL);
    if (ug == ug && vg == vg) {
        atomicAdd(fx + c, (unsigned long long)llrint(ug * 1048576.0));
        atomicAdd(fy + c, (unsigned long long)llrint(vg * 1048576.0));
    }
}

// FAST build of the speed cut-off test |2 atan2(sqrt a, sqrt(1-a))| >= cut (wr.py:97-112, 844-850):
// the central angle is monotonic in a on [0, pi], so compare a with sin^2(cut/2); cos of both latitudes
// are already known (cos_c from the last stage, cos_p cached from the previous output).
RWRT_COLD double haversine_cold(double lon_c, double lat_c, double lon_p, double lat_p) { return haversine(lon_c, lat_c, lon_p, lat_p); }
__device__ __forceinline__ bool speed_cutoff_fast(double lon_c, double lat_c, double lon_p, double lat_p, double cos_c,
                                                  double cos_p, double sin2_half_cut)
{
    double s1 = sin_small((lat_c - lat_p) * 0.5), s2 = sin_small((lon_c - lon_p) * 0.5);
    double a = s1 * s1 + cos_p * cos_c * (s2 * s2);
    return a >= sin2_half_cut;
}

}  // namespace rwrt
