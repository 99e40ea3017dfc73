/*
 * rwrt_oracle.c -- TEST INFRASTRUCTURE.  NOT PART OF THE PRODUCT.
 *
 * Plain-C, single-thread-per-ray CPU restatement of the reference's RK45 ray-integration
 * path (yinan-codes/Rossby-wave-ray-tracing).  It exists only so that tests/, the smoke
 * check and bench.py's cpu_baseline leg can check / time the CUDA path against an
 * independent implementation of the same algorithm.  Nothing under
 * rossby-wave-ray-tracing_b200/ may include, link or call this file.
 *
 * Parity pin: this restatement is checked against outputs of the UNMODIFIED reference run
 * in the build container (oracle/refshim.py -> oracle/make_golden.py -> tests/golden/ *.npz;
 * tests/test_oracle_golden.py) and against the solver known-answer ODEs of rkf45.py's
 * __main__ demo.  The reference ships no golden vectors of its own.
 *
 * Every function cites the reference file:line it follows.  Compile WITHOUT fp contraction
 * and WITHOUT fast-math (oracle/build_oracle.py: gcc -O2 -ffp-contract=off), because the
 * reference evaluates every expression as separate IEEE-754 double operations (NumPy /
 * numba without fastmath).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RWRT_PI 3.14159265358979323846264338327950288419716939937510 /* constants.py:13 */
#define RWRT_REARTH 6.3712e6                                          /* constants.py:16 */
#define RWRT_OMEGA 7.2921e-5                                          /* constants.py:17 */
#define RWRT_DELT 1.0e-8                                              /* constants.py:28 */
#define NF 18                                                         /* bs.py:349-368 */

/* ------------------------------------------------------------------------------------ */
/* Basic state: finite differences, absolute vorticity, smoothing, stacking             */
/* ------------------------------------------------------------------------------------ */

/* bs.py:121-131 gradient_x: centred, periodic in i.  Arrays are [i_lon][j_lat]. */
static void grad_x(int nx, int ny, double dx, const double *f, double *o)
{
    double den = 2.0 * dx;
    for (int i = 0; i < nx; ++i) {
        int ip = (i + 1) % nx, im = (i + nx - 1) % nx;
        for (int j = 0; j < ny; ++j)
            o[i * ny + j] = (f[ip * ny + j] - f[im * ny + j]) / den;
    }
}

/* bs.py:133-142 gradient_y: centred inside, one-sided at the two pole rows. */
static void grad_y(int nx, int ny, double dy, const double *f, double *o)
{
    double den = 2.0 * dy;
    for (int i = 0; i < nx; ++i) {
        const double *r = f + (size_t)i * ny;
        double *w = o + (size_t)i * ny;
        for (int j = 1; j < ny - 1; ++j) w[j] = (r[j + 1] - r[j - 1]) / den;
        w[0] = (r[1] - r[0]) / dy;
        w[ny - 1] = (r[ny - 1] - r[ny - 2]) / dy;
    }
}

/* bs.py:144-154 gradient_xx: 3-point periodic. */
static void grad_xx(int nx, int ny, double dx, const double *f, double *o)
{
    double den = dx * dx;
    for (int i = 0; i < nx; ++i) {
        int ip = (i + 1) % nx, im = (i + nx - 1) % nx;
        for (int j = 0; j < ny; ++j)
            o[i * ny + j] = (f[ip * ny + j] - 2.0 * f[i * ny + j] + f[im * ny + j]) / den;
    }
}

/* bs.py:156-166 gradient_yy: 3-point inside; pole rows copy their neighbour. */
static void grad_yy(int nx, int ny, double dy, const double *f, double *o)
{
    double den = dy * dy;
    for (int i = 0; i < nx; ++i) {
        const double *r = f + (size_t)i * ny;
        double *w = o + (size_t)i * ny;
        for (int j = 1; j < ny - 1; ++j) w[j] = (r[j + 1] - 2.0 * r[j] + r[j - 1]) / den;
        w[0] = w[1];
        w[ny - 1] = w[ny - 2];
    }
}

/* bs.py:168-195 gradient_xy: 4-point cross difference, periodic in i, copy at pole rows. */
static void grad_xy(int nx, int ny, double dx, double dy, const double *f, double *o)
{
    double den = 4.0 * dx * dy;
    for (int i = 0; i < nx; ++i) {
        int ip = (i + 1) % nx, im = (i + nx - 1) % nx;
        double *w = o + (size_t)i * ny;
        for (int j = 1; j < ny - 1; ++j)
            w[j] = (f[ip * ny + j + 1] - f[ip * ny + j - 1] - f[im * ny + j + 1] + f[im * ny + j - 1]) / den;
        w[0] = w[1];
        w[ny - 1] = w[ny - 2];
    }
}

/* Same stencil on float32 input: bs.py:334-335 pass self.u / self.v (float32) straight into
 * gradient_xy, which does not cast (bs.py:168-176), so the differences are formed in float32
 * and only the division by the float64 (4 dx dy) promotes. */
static void grad_xy_f32(int nx, int ny, double dx, double dy, const float *f, double *o)
{
    double den = 4.0 * dx * dy;
    for (int i = 0; i < nx; ++i) {
        int ip = (i + 1) % nx, im = (i + nx - 1) % nx;
        double *w = o + (size_t)i * ny;
        for (int j = 1; j < ny - 1; ++j) {
            float d = f[ip * ny + j + 1] - f[ip * ny + j - 1];
            d = d - f[im * ny + j + 1];
            d = d + f[im * ny + j - 1];
            w[j] = (double)d / den;
        }
        w[0] = w[1];
        w[ny - 1] = w[ny - 2];
    }
}

/* bs.py:291-305 smth9: Jacobi 9-point smoother on the window [1:-2, 1:-2]; the convolution is
 * scipy.ndimage.convolve(mode='constant'): a running sum over the 3x3 kernel in row-major order. */
static void smth9(int nx, int ny, double *f, double *tmp)
{
    const double p = 0.5, q = 0.25;
    const double kq = q / 4, kp = p / 4, kc = -(p + q);
    const double w[3][3] = {{kq, kp, kq}, {kp, kc, kp}, {kq, kp, kq}};
    memcpy(tmp, f, sizeof(double) * (size_t)nx * ny);
    for (int i = 1; i <= nx - 3; ++i)
        for (int j = 1; j <= ny - 3; ++j) {
            double s = 0.0;
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) s += w[a][b] * tmp[(i + a - 1) * ny + (j + b - 1)];
            f[i * ny + j] = tmp[i * ny + j] + s;
        }
}

/*
 * bs.py:264-279 calc_absolute_vorticity + bs.py:318-407 ready.
 *   u, v      : (nlon, nlat) float32
 *   lat       : (nlat) float64 radians;  coslat/sinlat: optional tables (NULL -> libm).  The
 *               tests pass numpy's cos/sin so that this check is independent of libm flavour.
 *   fields    : ((nlon + xcyclic), nlat, 18) float64, order u,v,ux,uy,vx,vy,qx,qy,qxx,qxy,qyx,
 *               qyy,qxxx,qxxy,qxyy,qyyy,qyxx,qyyx (bs.py:349-372)
 *   diag      : optional 9 planes (nlon, nlat): q,uxx,uyy,vxx,vyy,uxy,vxy,betam,KS
 */
int orc_bs_ready(int nx, int ny, const float *u, const float *v, const double *lat, const double *coslat_in,
                 const double *sinlat_in, int xcyclic, double *fields, double *diag)
{
    size_t n = (size_t)nx * ny;
    double dx = 2.0 * RWRT_PI / nx;       /* bs.py:77 */
    double dy = RWRT_PI / (ny - 1);       /* bs.py:78 */
    enum { U, V, UX, UY, VX, VY, QX, QY, QXX, QXY, QYX, QYY, QXXX, QXXY, QXYY, QYYY, QYXX, QYYX,
           Q, UXX, UYY, VXX, VYY, UXY, VXY, UCOS, UCOSY, VXQ, TMP, NPLANE };
    double *P = (double *)malloc(sizeof(double) * n * NPLANE);
    double *cl = (double *)malloc(sizeof(double) * ny * 2), *sl = cl + ny;
    if (!P || !cl) { free(P); free(cl); return -1; }
#define PL(k) (P + (size_t)(k) * n)
    for (int j = 0; j < ny; ++j) {
        cl[j] = coslat_in ? coslat_in[j] : cos(lat[j]);
        sl[j] = sinlat_in ? sinlat_in[j] : sin(lat[j]);
    }
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j) {
            PL(U)[i * ny + j] = (double)u[i * ny + j];
            PL(V)[i * ny + j] = (double)v[i * ny + j];
            PL(UCOS)[i * ny + j] = (double)u[i * ny + j] * cl[j]; /* bs.py:270 */
        }
    grad_y(nx, ny, dy, PL(UCOS), PL(UCOSY));
    grad_x(nx, ny, dx, PL(V), PL(VXQ));
    for (int i = 0; i < nx; ++i) { /* bs.py:276-279 */
        double *q = PL(Q) + (size_t)i * ny;
        for (int j = 1; j < ny - 1; ++j)
            q[j] = (PL(VXQ)[i * ny + j] - PL(UCOSY)[i * ny + j]) / cl[j] + 2.0 * RWRT_OMEGA * sl[j] * RWRT_REARTH;
        q[0] = q[1];
        q[ny - 1] = q[ny - 2];
    }
    grad_x(nx, ny, dx, PL(U), PL(UX));    /* bs.py:322-343, in the reference's order */
    grad_y(nx, ny, dy, PL(U), PL(UY));
    grad_x(nx, ny, dx, PL(V), PL(VX));
    grad_y(nx, ny, dy, PL(V), PL(VY));
    grad_x(nx, ny, dx, PL(Q), PL(QX));
    grad_y(nx, ny, dy, PL(Q), PL(QY));
    grad_xx(nx, ny, dx, PL(U), PL(UXX));
    grad_yy(nx, ny, dy, PL(U), PL(UYY));
    grad_xx(nx, ny, dx, PL(V), PL(VXX));
    grad_yy(nx, ny, dy, PL(V), PL(VYY));
    grad_xx(nx, ny, dx, PL(Q), PL(QXX));
    grad_yy(nx, ny, dy, PL(Q), PL(QYY));
    grad_xy_f32(nx, ny, dx, dy, u, PL(UXY));
    grad_xy_f32(nx, ny, dx, dy, v, PL(VXY));
    grad_xy(nx, ny, dx, dy, PL(Q), PL(QXY));
    memcpy(PL(QYX), PL(QXY), sizeof(double) * n); /* bs.py:337: copy of the UNSMOOTHED qxy */
    grad_x(nx, ny, dx, PL(QXX), PL(QXXX));
    grad_y(nx, ny, dy, PL(QXX), PL(QXXY));
    grad_y(nx, ny, dy, PL(QXY), PL(QXYY));
    grad_y(nx, ny, dy, PL(QYY), PL(QYYY));
    grad_x(nx, ny, dx, PL(QXY), PL(QYXX));
    grad_x(nx, ny, dx, PL(QYY), PL(QYYX));
    smth9(nx, ny, PL(QXX), PL(TMP));      /* bs.py:345-347 */
    smth9(nx, ny, PL(QYY), PL(TMP));
    smth9(nx, ny, PL(QXY), PL(TMP));
    if (fields) {
        int nxo = nx + (xcyclic ? 1 : 0);
        for (int i = 0; i < nxo; ++i) {
            int is = i % nx; /* bs.py:370-372 wrap column = copy of column 0 */
            for (int j = 0; j < ny; ++j)
                for (int k = 0; k < NF; ++k) fields[((size_t)i * ny + j) * NF + k] = PL(k)[is * ny + j];
        }
    }
    if (diag) {
        const int src[7] = {Q, UXX, UYY, VXX, VYY, UXY, VXY};
        for (int k = 0; k < 7; ++k) memcpy(diag + k * n, PL(src[k]), sizeof(double) * n);
        double *bm = diag + 7 * n, *ks = diag + 8 * n;
        for (int i = 0; i < nx; ++i)
            for (int j = 0; j < ny; ++j) {
                size_t a = (size_t)i * ny + j;
                if (j == 0 || j == ny - 1) { bm[a] = NAN; ks[a] = NAN; continue; }
                double uu = (double)u[a];
                /* bs.py:379-383 */
                bm[a] = (2 * RWRT_OMEGA * (cl[j] * cl[j]) +
                         (-cl[j] * PL(UYY)[a] + sl[j] * PL(UY)[a] + uu / cl[j]) / RWRT_REARTH) / RWRT_REARTH;
                /* bs.py:393-404 */
                if (bm[a] > 0 && uu > 0) ks[a] = sqrt(bm[a] * cl[j] / uu) * RWRT_REARTH;
                else ks[a] = NAN;
            }
    }
#undef PL
    free(P);
    free(cl);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Field sampling + Mercator transform                                                  */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    int nlon_f;            /* first dimension of fields (nlon + 1 when xcyclic) */
    int nlat;
    const double *fields;  /* (nlon_f, nlat, 18) */
    double lon0, dlon;     /* lon[0], lon[1]-lon[0]   interpolation.py:78 */
    double lat0, dlat;     /* lat[0], lat[1]-lat[0]   interpolation.py:79 */
} orc_grid;

/* numpy float remainder (np.remainder / Python %): result takes the sign of the divisor. */
static double py_mod(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

static int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* int32 cast of floor(x) the way numpy does it on x86-64 (cvttsd2si: NaN/overflow -> INT_MIN) */
static int floor_i32(double x)
{
    double f = floor(x);
    if (!(f >= -2147483648.0 && f <= 2147483647.0)) return INT32_MIN;
    return (int)f;
}

/*
 * bs.py:513-519 + 781-887 cal_bs_mercator_point(mode='numpy'), interpolation.py:77-85 and 103-135.
 * Returns the 12 Mercator quantities the ray equations use:
 *   m[0..11] = fmu, fmv, fmux, fmuy, fmvx, fmvy, fmqx, fmqy, fmqxx, fmqxy, fmqyx, fmqyy
 */
static void sample_mercator(const orc_grid *g, double lon, double lat, double *m)
{
    const double twopi = 2 * RWRT_PI;
    if (!(fabs(lat) <= 0.5 * RWRT_PI)) { /* bs.py:787: out-of-range (or NaN) rows stay NaN */
        for (int k = 0; k < 12; ++k) m[k] = NAN;
        return;
    }
    double l1 = py_mod(lon, twopi);  /* bs.py:519 */
    double l2 = py_mod(l1, twopi);   /* interpolation.py:80 */
    double x = (l2 - g->lon0) / g->dlon;
    double y = (lat - g->lat0) / g->dlat;
    int W = g->nlon_f, H = g->nlat;
    int x0u = floor_i32(x), y0u = floor_i32(y);
    int x1 = clipi((int)((int64_t)x0u + 1 > INT32_MAX ? INT32_MAX : x0u + 1), 0, W - 1);
    int y1 = clipi(y0u + 1, 0, H - 1);
    int x0 = clipi(x0u, 0, W - 1), y0 = clipi(y0u, 0, H - 1);
    const double *a = g->fields + ((size_t)x0 * H + y1) * NF;
    const double *b = g->fields + ((size_t)x1 * H + y1) * NF;
    const double *c = g->fields + ((size_t)x0 * H + y0) * NF;
    const double *d = g->fields + ((size_t)x1 * H + y0) * NF;
    double sx = x - x0, sy = y - y0;
    double wa = (1 - sx) * sy, wb = sx * sy, wc = (1 - sx) * (1 - sy), wd = sx * (1 - sy);
    double f[12];
    for (int k = 0; k < 12; ++k) f[k] = a[k] * wa + b[k] * wb + c[k] * wc + d[k] * wd;
    double fu = f[0], fv = f[1], fux = f[2], fuy = f[3], fvx = f[4], fvy = f[5], fqx = f[6], fqy = f[7], fqxx = f[8],
           fqxy = f[9], fqyy = f[11];
    double cp = cos(lat), sp = sin(lat);
    double mask = (fabs(cp) <= 0.0175) ? 0.0 : 1.0; /* bs.py:858-860 */
    cp = cp * mask + (1 - mask) * 1e-6;
    double tl = tan(lat);
    m[8] = fqxx * mask;                              /* fmqxx bs.py:862 */
    m[10] = fqxy * cp * mask;                        /* fmqyx */
    m[9] = m[10] * mask;                             /* fmqxy */
    m[11] = (fqyy * cp - fqy * sp) * cp * mask;      /* fmqyy */
    m[6] = fqx * mask;                               /* fmqx  bs.py:872 */
    m[7] = fqy * cp * mask;                          /* fmqy */
    m[2] = fux / cp * mask;                          /* fmux  bs.py:875 */
    m[4] = fvx / cp * mask;                          /* fmvx */
    m[3] = (fuy + tl * fu) * mask;                   /* fmuy */
    m[5] = (fvy + tl * fv) * mask;                   /* fmvy */
    m[0] = fu / cp * mask;                           /* fmu   bs.py:880 */
    m[1] = fv / cp * mask;                           /* fmv */
}

/* wn.py:266-294 core_cal_ugvg_extent */
static void ugvg_extent(double fu, double fv, double fqx, double fqy, double k, double l, double *ug, double *vg)
{
    double kap = l / k;
    double kap2 = kap * kap;
    double kap1 = 1.0 + kap2;
    double KK = k * k * kap1;
    double den = KK * kap1;
    *ug = fu + (((1. - kap2) * fqy) - (2. * kap * fqx)) / den;
    *vg = fv + ((2. * kap * fqy) + ((1. - kap2) * fqx)) / den;
}

/* wn.py:209-259 cal_ugvg_numpy (t = 0 only, wr.py:393) */
static void ugvg_numpy(double fu, double fv, double fqx, double fqy, double k, double l, double *ug, double *vg)
{
    if (k == 0) { *ug = 0; *vg = 0; return; }
    double nans = (l * 0) * (fu * fqx * fqy * 0) + 1; /* wn.py:232-233 */
    if (isnan(nans)) nans = 0;
    double kap1 = k * k - l * l;
    double kap2 = 2 * k * l;
    double KK2 = k * k + l * l;
    double kk4 = KK2 * KK2;
    *ug = (fu + (kap1 * fqy - kap2 * fqx) / kk4) * nans;
    *vg = (fv + (kap1 * fqx + kap2 * fqy) / kk4) * nans;
}

/*
 * One right-hand-side evaluation: wr.py:784-789 fun -> wr.py:492-556 diffun_numpy ->
 * bs.py:781-887 -> wn.py:266-294 -> wr.py:44-82 core_diffun.   y = (lon, lat, k, l, A).
 */
static void rhs(const orc_grid *g, const double *y, double *dy)
{
    double lon = y[0], lat = y[1], kx = y[2], ky = y[3], amp = y[4];
    int err = (fabs(lat) >= 0.5 * RWRT_PI) || (fabs(ky) >= 100.0); /* wr.py:508-510 */
    if (err) ky = NAN;                                              /* wr.py:513-514 */
    double m[12];
    sample_mercator(g, lon, lat, m);
    double fmu = m[0], fmv = m[1], fmux = m[2], fmuy = m[3], fmvx = m[4], fmvy = m[5], fmqx = m[6], fmqy = m[7],
           fmqxx = m[8], fmqxy = m[9], fmqyx = m[10], fmqyy = m[11];
    double ug, vg;
    ugvg_extent(fmu, fmv, fmqx, fmqy, kx, ky, &ug, &vg);
    /* wr.py:53-78 */
    double kap = ky / kx;
    double kap2 = kap * kap;
    double kap1 = 1 + kap * kap;
    double kk = kx * kx * kap1;
    double dzwn = -kx * ((fmux + kap * fmvx) + (kap * fmqxx - fmqyx) / kk);
    double dmwn = -kx * ((fmuy + kap * fmvy) + (kap * fmqxy - fmqyy) / kk);
    double damp1 = 2.0 * (fmux + fmvy + kap * (fmvx + fmuy)) / kap1;
    double damp2 = 2.0 * (kap * (fmqxx - fmqyy) + (kap2 - 1.0) * fmqxy) / (kk * kap1);
    double damp3 = -2.0 * sin(lat) * fmv;
    double damp = damp1 + damp2 + damp3;
    dy[0] = ug / RWRT_REARTH;
    dy[1] = (vg * cos(lat)) / RWRT_REARTH;
    dy[2] = dzwn / RWRT_REARTH;
    dy[3] = dmwn / RWRT_REARTH;
    dy[4] = (damp * amp) / RWRT_REARTH;
    if (err) for (int c = 0; c < 5; ++c) dy[c] = NAN; /* wr.py:552-553 */
}

void orc_rhs(int nlon_f, int nlat, const double *fields, const double *lon, const double *lat, int n, const double *y,
             double *dy)
{
    orc_grid g = {nlon_f, nlat, fields, lon[0], lon[1] - lon[0], lat[0], lat[1] - lat[0]};
    for (int i = 0; i < n; ++i) rhs(&g, y + 5 * i, dy + 5 * i);
}

/* 12 Mercator quantities at n points; out is (n, 12) */
void orc_sample_mercator(int nlon_f, int nlat, const double *fields, const double *lon, const double *lat, int n,
                         const double *plon, const double *plat, double *out)
{
    orc_grid g = {nlon_f, nlat, fields, lon[0], lon[1] - lon[0], lat[0], lat[1] - lat[0]};
    for (int i = 0; i < n; ++i) sample_mercator(&g, plon[i], plat[i], out + 12 * i);
}

/* ------------------------------------------------------------------------------------ */
/* Dormand-Prince 5(4) with the reference's controller                                   */
/* ------------------------------------------------------------------------------------ */

/* rkf45.py:604-615 */
static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};

/* rkf45.py:29-31: np.linalg.norm(x, axis=0) / 5**0.5 -- squares summed row by row */
static double rms5(const double *x)
{
    double s = x[0] * x[0];
    for (int c = 1; c < 5; ++c) s = s + x[c] * x[c];
    return sqrt(s) / sqrt(5.0);
}

typedef void (*rhs_fn)(const void *ctx, const double *y, double *dy);

static void rhs_grid(const void *ctx, const double *y, double *dy) { rhs((const orc_grid *)ctx, y, dy); }

/* rkf45.py:34-99 select_initial_step (direction = +1, order = 4) */
static double initial_step(rhs_fn fn, const void *ctx, const double *y0, const double *f0, double rtol, double atol,
                           int n)
{
    double scale[8], a[8], y1[8], f1[8];
    /* the reference norm is over n components (n == 5 on the ray path; the KAT ODEs use n <= 5) */
    double d0s = 0, d1s = 0;
    for (int c = 0; c < n; ++c) {
        scale[c] = atol + fabs(y0[c]) * rtol;
        a[c] = y0[c] / scale[c];
        d0s = (c == 0) ? a[c] * a[c] : d0s + a[c] * a[c];
    }
    for (int c = 0; c < n; ++c) {
        a[c] = f0[c] / scale[c];
        d1s = (c == 0) ? a[c] * a[c] : d1s + a[c] * a[c];
    }
    double rn = sqrt((double)n);
    double d0 = sqrt(d0s) / rn, d1 = sqrt(d1s) / rn;
    double h0 = 0.01 * d0 / d1;
    if (d0 < 1e-5) h0 = 1e-6;
    if (d1 < 1e-5) h0 = 1e-6;
    for (int c = 0; c < n; ++c) y1[c] = y0[c] + h0 * 1.0 * f0[c];
    fn(ctx, y1, f1);
    double d2s = 0;
    for (int c = 0; c < n; ++c) {
        a[c] = (f1[c] - f0[c]) / scale[c];
        d2s = (c == 0) ? a[c] * a[c] : d2s + a[c] * a[c];
    }
    double d2 = (sqrt(d2s) / rn) / h0;
    int m1 = !(d1 > 1e-15), m2 = !(d2 > 1e-15);
    double mx = isnan(d1) ? d2 : (isnan(d2) ? d1 : (d1 > d2 ? d1 : d2)); /* np.nanmax */
    double h1 = pow(0.01 / mx, 1.0 / (4 + 1));
    if (m1 && m2) { double t = h0 * 1e-3; h1 = (1e-6 > t || isnan(t)) ? (isnan(t) ? t : 1e-6) : t; } /* np.maximum */
    double hh = 100 * h0;
    if (isnan(hh) || isnan(h1)) return NAN; /* np.min propagates NaN */
    return hh < h1 ? hh : h1;
}

typedef struct {
    double t, h_abs;
    double y[8];
    int n;
} rk_state;

/*
 * Advance one ray from its current t to t_bound: the per-ray view of repeated
 * OdeSolver.step (rkf45.py:222-253) -> RungeKutta._step_impl (rkf45.py:375-514) passes.
 * Returns the number of accepted steps (a step whose entry state was non-NaN; the
 * definition of the metric) and adds the attempts to *attempts.
 */
static int64_t advance_to(rhs_fn fn, const void *ctx, rk_state *s, double t_bound, double rtol, double atol,
                          double gmin, int64_t *attempts)
{
    int n = s->n;
    int64_t acc = 0;
    double K[7][8], ys[8], yn[8];
    for (;;) {
        /* rkf45.py:400-403: NaN state -> parked at t_bound; t == t_bound -> waits */
        double my = 0;
        for (int c = 0; c < n; ++c) my += s->y[c];
        if (isnan(my)) { s->t = t_bound; return acc; }
        if (s->t == t_bound) return acc;
        fn(ctx, s->y, K[0]);                                  /* rkf45.py:378 */
        double h_abs = s->h_abs;                              /* rkf45.py:383-387: clamp once per pass */
        if (!(h_abs >= gmin) && !isnan(h_abs)) h_abs = gmin;
        int rejected = 0;
        double t_new, h;
        for (;;) {
            /* a NaN step (rkf45.py:423-425 aborts only when EVERY pending ray has one) flows through:
             * t_new = NaN, the stages are NaN, the error norm is NaN -> 0 -> accepted (rkf45.py:446) */
            t_new = s->t + h_abs;                             /* rkf45.py:426-435 */
            if (t_new - t_bound > 0) t_new = t_bound;
            h = t_new - s->t;
            h_abs = fabs(h);
            ++*attempts;
            for (int st = 1; st < 6; ++st) {                  /* rkf45.py:306-311 */
                for (int c = 0; c < n; ++c) {
                    double acc_ = K[0][c] * RK_A[st][0];
                    for (int j = 1; j < st; ++j) acc_ = acc_ + K[j][c] * RK_A[st][j];
                    ys[c] = s->y[c] + acc_ * h;
                }
                fn(ctx, ys, K[st]);
            }
            for (int c = 0; c < n; ++c) {                     /* rkf45.py:316 */
                double acc_ = K[0][c] * RK_B[0];
                for (int j = 1; j < 6; ++j) acc_ = acc_ + K[j][c] * RK_B[j];
                yn[c] = s->y[c] + h * acc_;
            }
            fn(ctx, yn, K[6]);                                /* rkf45.py:317-319 */
            double ssum = 0;
            for (int c = 0; c < n; ++c) {                     /* rkf45.py:442-446 */
                double ay = fabs(s->y[c]), an = fabs(yn[c]);
                double mx = (isnan(ay) || isnan(an)) ? NAN : (ay > an ? ay : an); /* np.maximum */
                double scale = atol + mx * rtol;
                double e = K[0][c] * RK_E[0];
                for (int j = 1; j < 7; ++j) e = e + K[j][c] * RK_E[j];
                double r = (h * e) / scale;
                ssum = (c == 0) ? r * r : ssum + r * r;
            }
            double err = sqrt(ssum) / sqrt((double)n);
            if (isnan(err)) err = 0;
            if (err < 1) {                                    /* rkf45.py:449-471 */
                double factor = 0.9 * pow(err, -0.2);
                if (!(factor < 10)) factor = 10;              /* np.minimum(10, .) incl. inf */
                if (err == 0) factor = 10;
                if (rejected && factor > 1.0) factor = 1.0;
                h_abs *= factor;
                break;
            }
            double factor = 0.9 * pow(err, -0.2);             /* rkf45.py:474-479 */
            if (!(factor > 0.2)) factor = 0.2;
            h_abs *= factor;
            rejected = 1;
        }
        s->t = isnan(t_new) ? t_bound : t_new;                /* rkf45.py:503 */
        s->h_abs = h_abs;
        for (int c = 0; c < n; ++c) s->y[c] = yn[c];
        ++acc;
    }
}

/* wr.py:97-112 cal_dis */
static double haversine(double lon_c, double lat_c, double lon_p, double lat_p)
{
    double dlon = lon_c - lon_p, dlat = lat_c - lat_p;
    double s1 = sin(dlat / 2.0), s2 = sin(dlon / 2.0);
    double a = s1 * s1 + cos(lat_p) * cos(lat_c) * (s2 * s2);
    double c = 2 * atan2(sqrt(a), sqrt(1.0 - a));
    return fabs(c);
}

/*
 * wr.py:767-887 core_ray_run_rk45 for every ray slot, one ray at a time (rays are independent:
 * SURVEY.md probe9).  out = 7 planes (rlon,rlat,rzwn,rmwn,ramp,rug,rvg), each (nt, N); row 0 must
 * hold the initial state (wr.py:768-781 read it from there); rows 1.. are written here.
 * accepted (N) int64, attempts (N) int64 (may be NULL).
 */
int orc_ray_run_rk45(int nlon_f, int nlat, const double *fields, const double *lon, const double *lat, int64_t N,
                     int nt, double tstep, double ttotal, double rtol, double atol, double min_step_factor,
                     double cut_off_factor, double *out, int64_t *accepted, int64_t *attempts_out)
{
    orc_grid g = {nlon_f, nlat, fields, lon[0], lon[1] - lon[0], lat[0], lat[1] - lat[0]};
    const double eps = 2.220446049250313e-16;
    if (rtol < 100 * eps) rtol = 100 * eps;                    /* rkf45.py:21-26 */
    double gmin = min_step_factor * tstep;                     /* wr.py:794, rkf45.py:362 */
    if ((tstep - 0) * 0.001 < gmin) gmin = (tstep - 0) * 0.001;
    double cut = cut_off_factor * tstep / 3600.;               /* wr.py:170 */
    size_t pl = (size_t)nt * N;
    double *rlon = out, *rlat = out + pl, *rzwn = out + 2 * pl, *rmwn = out + 3 * pl, *ramp = out + 4 * pl,
           *rug = out + 5 * pl, *rvg = out + 6 * pl;
    for (int64_t r = 0; r < N; ++r) {
        rk_state s;
        s.n = 5;
        s.t = 0;
        s.y[0] = rlon[r]; s.y[1] = rlat[r]; s.y[2] = rzwn[r]; s.y[3] = rmwn[r]; s.y[4] = ramp[r];
        double f0[5];
        rhs(&g, s.y, f0);                                      /* rkf45.py:345 */
        s.h_abs = initial_step(rhs_grid, &g, s.y, f0, rtol, atol, 5); /* rkf45.py:347-349 */
        int64_t acc = 0, att = 0;
        double plon = rlon[r], plat = rlat[r];
        for (int i = 1; i < nt; ++i) {
            double tb = i * tstep;                             /* wr.py:798-801 */
            if (i == nt - 1 && tb > ttotal) tb = ttotal;
            acc += advance_to(rhs_grid, &g, &s, tb, rtol, atol, gmin, &att);
            double *y = s.y;
            if (fabs(y[1]) >= 0.5 * RWRT_PI)                   /* wr.py:838-843 */
                for (int c = 0; c < 5; ++c) y[c] = NAN;
            double dd = haversine(y[0], y[1], plon, plat);     /* wr.py:844-850 */
            if (fabs(dd) >= cut)
                for (int c = 0; c < 5; ++c) y[c] = NAN;
            double m[12], ug, vg;
            sample_mercator(&g, y[0], y[1], m);                /* wr.py:856-865 */
            ugvg_extent(m[0], m[1], m[6], m[7], y[2], y[3], &ug, &vg);
            size_t o = (size_t)i * N + r;
            rlon[o] = y[0]; rlat[o] = y[1]; rzwn[o] = y[2]; rmwn[o] = y[3]; ramp[o] = y[4];
            rug[o] = ug; rvg[o] = vg;
            plon = y[0]; plat = y[1];
        }
        if (accepted) accepted[r] = acc;
        if (attempts_out) attempts_out[r] = att;
    }
    return 0;
}

/*
 * wr.py:702-765 core_ray_run_numpy + wr.py:583-622 rk4_step_numpy + wr.py:89-95 core_rk4_step: the
 * fixed-step RK4 path (inte_method='' -- main_wr.py's default), one ray at a time.  State is the five
 * integrated components; ug / vg of the stored rows are the fresh-sample values (wr.py:737-756), the
 * step-averaged ug / vg of core_rk4_step never reach the outputs.
 * "Freeze" semantics: if the ray is valid at the step start (mask1 false) but a later stage input hits the
 * |lat| >= pi/2 / |l| >= 100 mask, the step is not taken and the old state is kept (wr.py:608-618).
 * NaN-root slots are NOT masked (comparisons with NaN are False): their step gives NaN in all components.
 */
static int rhs_masked(const orc_grid *g, const double *y, double *dy)
{
    int err = (fabs(y[1]) >= 0.5 * RWRT_PI) || (fabs(y[3]) >= 100.0); /* wr.py:508-510 */
    rhs(g, y, dy);
    return err;
}

int orc_ray_run_rk4(int nlon_f, int nlat, const double *fields, const double *lon, const double *lat, int64_t N,
                    int nt, double tstep, double cut_off_factor, double *out)
{
    orc_grid g = {nlon_f, nlat, fields, lon[0], lon[1] - lon[0], lat[0], lat[1] - lat[0]};
    double cut = cut_off_factor * tstep / 3600.;
    double dt = tstep;
    size_t pl = (size_t)nt * N;
    double *pln[7];
    for (int k = 0; k < 7; ++k) pln[k] = out + k * pl;
    for (int64_t r = 0; r < N; ++r) {
        double y[5];
        for (int c = 0; c < 5; ++c) y[c] = pln[c][r];
        double plon = y[0], plat = y[1];
        for (int it = 0; it < nt - 1; ++it) {
            double k1[5], k2[5], k3[5], k4[5], ys[5], yn[5];
            int m1 = rhs_masked(&g, y, k1);
            for (int c = 0; c < 5; ++c) yn[c] = y[c];
            if (!m1) {
                for (int c = 0; c < 5; ++c) ys[c] = y[c] + 0.5 * dt * k1[c];
                int m2 = rhs_masked(&g, ys, k2);
                for (int c = 0; c < 5; ++c) ys[c] = y[c] + 0.5 * dt * k2[c];
                int m3 = rhs_masked(&g, ys, k3);
                for (int c = 0; c < 5; ++c) ys[c] = y[c] + dt * k3[c];
                int m4 = rhs_masked(&g, ys, k4);
                if (!m2 && !m3 && !m4)
                    for (int c = 0; c < 5; ++c) /* wr.py:92-93 */
                        yn[c] = y[c] + (dt / 6.0) * (k1[c] + 2 * k2[c] + 2 * k3[c] + k4[c]);
            }
            if (fabs(yn[1]) >= 0.5 * RWRT_PI)                  /* wr.py:721-726 */
                for (int c = 0; c < 5; ++c) yn[c] = NAN;
            double dd = haversine(yn[0], yn[1], plon, plat);   /* wr.py:728-734 */
            if (fabs(dd) >= cut)
                for (int c = 0; c < 5; ++c) yn[c] = NAN;
            double m[12], ug, vg;
            sample_mercator(&g, yn[0], yn[1], m);              /* wr.py:737-746 */
            ugvg_extent(m[0], m[1], m[6], m[7], yn[2], yn[3], &ug, &vg);
            size_t o = (size_t)(it + 1) * N + r;
            for (int c = 0; c < 5; ++c) { pln[c][o] = yn[c]; y[c] = yn[c]; }
            pln[5][o] = ug;
            pln[6][o] = vg;
            plon = yn[0];
            plat = yn[1];
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Known-answer ODEs of rkf45.py's __main__ demo (rkf45.py:772-882): the solver alone      */
/* ------------------------------------------------------------------------------------ */

static void kat_rhs(const void *ctx, const double *y, double *dy)
{
    const double *p = (const double *)ctx; /* p[0] = kind, p[1] = current t (set by caller; unused by autonomous) */
    int kind = (int)p[0];
    if (kind == 0) { /* Lorenz-63, sigma=10, rho=28, beta=8/3 */
        dy[0] = 10.0 * (y[1] - y[0]);
        dy[1] = y[0] * (28.0 - y[2]) - y[1];
        dy[2] = y[0] * y[1] - (8.0 / 3.0) * y[2];
    } else if (kind == 1) { /* augmented dx/dt = 2 t  as (x, t) */
        dy[0] = 2.0 * y[1];
        dy[1] = 1.0;
    } else { /* augmented dx/dt = exp(0.1 t) as (x, t) */
        dy[0] = exp(0.1 * y[1]);
        dy[1] = 1.0;
    }
}

/* Integrate a KAT system from 0 to t_end with the reference controller; returns accepted steps. */
int64_t orc_kat(int kind, int n, const double *y0, double t_end, double rtol, double atol, double gmin, double *y_out)
{
    double p[2] = {(double)kind, 0};
    rk_state s;
    s.n = n;
    s.t = 0;
    for (int c = 0; c < n; ++c) s.y[c] = y0[c];
    double f0[8];
    kat_rhs(p, s.y, f0);
    s.h_abs = initial_step(kat_rhs, p, s.y, f0, rtol, atol, n);
    int64_t att = 0;
    int64_t acc = advance_to(kat_rhs, p, &s, t_end, rtol, atol, gmin, &att);
    for (int c = 0; c < n; ++c) y_out[c] = s.y[c];
    return acc;
}

/* ------------------------------------------------------------------------------------ */
/* t = 0 state: dispersion cubic, slot order, initial group velocity                      */
/* ------------------------------------------------------------------------------------ */

/* bs.py:942-982 change_roots_order (mwn has 3 entries, NaN-padded; deg = number of real roots) */
static void change_roots_order(double *mwn, int deg)
{
    double t;
#define SWAP(a, b) (t = mwn[a], mwn[a] = mwn[b], mwn[b] = t)
    if (deg == 3) {
        int best = 1;
        for (int i = 1; i < 3; ++i)
            if (mwn[i] >= 0. && mwn[i] < mwn[best]) { SWAP(i, best); best = i; }
        if (mwn[0] < 0) SWAP(0, 1);
        if ((mwn[1] < 0 && mwn[2] < 0 && mwn[1] < mwn[2]) || (mwn[1] > 0 && mwn[2] < 0.)) SWAP(1, 2);
    } else if (deg == 2) {
        /* the loop body breaks on its first iteration in both branches (bs.py:962-968) */
        if (!isnan(mwn[0]) && mwn[0] > 0) { /* swap with itself */ }
        else SWAP(0, 1);
    } else if (deg == 1) {
        for (int i = 0; i < 3; ++i) {
            if (!isnan(mwn[i]) && mwn[i] >= 0 && i != 0) SWAP(i, 0);
            else if (!isnan(mwn[i]) && mwn[i] <= 0. && i != 2) SWAP(i, 1);
        }
    }
#undef SWAP
    for (int i = 0; i < 3; ++i)
        if (!isnan(mwn[i]) && fabs(mwn[i]) > 100.) mwn[i] = NAN;
    t = mwn[0]; mwn[0] = mwn[2]; mwn[2] = t; /* return mwn[::-1] */
}

/* companion-matrix eigenvalues in LAPACK zgeev order: see oracle/zgeev3.c (included below) */
#include "zgeev3.inc"

/*
 * bs.py:985-1040 cal_ky_numpy for one (source, k): fills l[3] (slot order of the reference).
 * The root ORDER that feeds change_roots_order is the order np.roots (numba -> LAPACK zgeev on
 * the companion matrix, bs.py:38-40) returns its eigenvalues in; zgeev3 reproduces it.
 */
static void cal_ky_point(double fu, double fv, double fqx, double fqy, double freq, double k, double *l)
{
    l[0] = l[1] = l[2] = NAN;
    if (k == 0) return;
    double ps = freq / k * RWRT_REARTH;                 /* bs.py:1005 */
    double coeff[4];
    coeff[0] = pow(k, 3.0) * (fu - ps - (fqy / (k * k)));  /* bs.py:1006-1011 (zwn**3, zwn**2) */
    coeff[1] = (k * k) * fv + fqx;
    coeff[2] = k * (fu - ps);
    coeff[3] = fv;
    int deg = 3;
    while (deg > 0 && fabs(coeff[deg]) == 0) --deg;     /* bs.py:1017-1019 */
    if (deg < 1) return;
    double p[4]; /* descending powers: coeff[::-1] */
    for (int i = 0; i <= deg; ++i) p[i] = coeff[deg - i];
    double wr[3], wi[3];
    int nr = orc_poly_roots(p, deg, wr, wi);            /* np.roots order */
    double mwn[3] = {NAN, NAN, NAN};
    int cnt = 0;
    for (int i = 0; i < nr; ++i)
        if (fabs(wi[i]) < RWRT_DELT) { if (cnt < 3) mwn[cnt] = wr[i]; ++cnt; }  /* bs.py:1030-1032 */
    change_roots_order(mwn, cnt);
    l[0] = mwn[0]; l[1] = mwn[1]; l[2] = mwn[2];
}


/*
 * wr.py:344-395 ray_initial_numpy: writes row 0 of the 7 planes, each plane (nt, 3, nsource, nzwn).
 */
int orc_ray_init(int nlon_f, int nlat, const double *fields, const double *lon, const double *lat, int nsource,
                 const double *src_lon, const double *src_lat, int nzwn, const double *zwn, double freq, int nt,
                 double *out)
{
    orc_grid g = {nlon_f, nlat, fields, lon[0], lon[1] - lon[0], lat[0], lat[1] - lat[0]};
    size_t N = (size_t)3 * nsource * nzwn, pl = (size_t)nt * N;
    double *rlon = out, *rlat = out + pl, *rzwn = out + 2 * pl, *rmwn = out + 3 * pl, *ramp = out + 4 * pl,
           *rug = out + 5 * pl, *rvg = out + 6 * pl;
    for (int s = 0; s < nsource; ++s) {
        double m[12];
        sample_mercator(&g, src_lon[s], src_lat[s], m);                  /* wr.py:360-365 */
        for (int iz = 0; iz < nzwn; ++iz) {
            double l[3];
            cal_ky_point(m[0], m[1], m[6], m[7], freq, zwn[iz], l);
            for (int ir = 0; ir < 3; ++ir) {
                size_t o = ((size_t)ir * nsource + s) * nzwn + iz;
                rlon[o] = 1.0 * src_lon[s];
                rlat[o] = 1.0 * src_lat[s];
                rzwn[o] = 1.0 * zwn[iz];
                rmwn[o] = l[ir];
                ramp[o] = isnan(l[ir]) ? NAN : 1.0;
                ugvg_numpy(m[0], m[1], m[6], m[7], zwn[iz], l[ir], &rug[o], &rvg[o]);
            }
        }
    }
    return 0;
}

/* roots of one polynomial (descending coefficients, degree deg <= 3) in np.roots order */
int orc_roots(const double *p, int deg, double *wr, double *wi) { return orc_poly_roots(p, deg, wr, wi); }

/* bs.py:985-1040 for arrays: out (n, 3) */
void orc_cal_ky(int n, const double *fu, const double *fv, const double *fqx, const double *fqy, double freq, double k,
                double *out)
{
    for (int i = 0; i < n; ++i) cal_ky_point(fu[i], fv[i], fqx[i], fqy[i], freq, k, out + 3 * i);
}
