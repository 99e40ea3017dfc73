// rwrt_bs.cu -- K1: basic-state derivative fields (BS.ready, bs.py:318-407).
//
// Compiled with -fmad=false: every expression below is the reference's NumPy expression with the
// same association, so the 18 stacked fields are bit-identical to bs.py's (given the same
// cos/sin(lat) tables).  HBM-bound stencil: 8 B read + (96 + 144) B written per grid point.
//
//   bs.py:121-131 gradient_x   centred, periodic in i          bs.py:133-142 gradient_y  one-sided at pole rows
//   bs.py:144-154 gradient_xx  3-point periodic                bs.py:156-166 gradient_yy pole rows copy neighbour
//   bs.py:168-195 gradient_xy  4-point, pole rows copy         bs.py:264-279 absolute vorticity q = a * zeta_abs
//   bs.py:291-305 smth9        Jacobi 9-point on [1:-2, 1:-2]  bs.py:349-372 stack order + wrap column
#include "rwrt_internal.h"

namespace rwrt {
namespace {

struct Geo {
    int nx, ny;
    double dx, dy;
};

__device__ __forceinline__ int wrapi(int i, int nx) { return i < 0 ? i + nx : (i >= nx ? i - nx : i); }

template <typename T>
__device__ __forceinline__ double gx(const T *f, const Geo &g, int i, int j)
{
    int ip = wrapi(i + 1, g.nx), im = wrapi(i - 1, g.nx);
    return ((double)f[(size_t)ip * g.ny + j] - (double)f[(size_t)im * g.ny + j]) / (2.0 * g.dx);
}

template <typename T>
__device__ __forceinline__ double gy(const T *f, const Geo &g, int i, int j)
{
    const T *r = f + (size_t)i * g.ny;
    if (j == 0) return ((double)r[1] - (double)r[0]) / g.dy;
    if (j == g.ny - 1) return ((double)r[g.ny - 1] - (double)r[g.ny - 2]) / g.dy;
    return ((double)r[j + 1] - (double)r[j - 1]) / (2.0 * g.dy);
}

template <typename T>
__device__ __forceinline__ double gxx(const T *f, const Geo &g, int i, int j)
{
    int ip = wrapi(i + 1, g.nx), im = wrapi(i - 1, g.nx);
    return ((double)f[(size_t)ip * g.ny + j] - 2.0 * (double)f[(size_t)i * g.ny + j] + (double)f[(size_t)im * g.ny + j]) /
           (g.dx * g.dx);
}

template <typename T>
__device__ __forceinline__ double gyy(const T *f, const Geo &g, int i, int j)
{
    int jj = min(max(j, 1), g.ny - 2);  // pole rows copy their neighbour
    const T *r = f + (size_t)i * g.ny;
    return ((double)r[jj + 1] - 2.0 * (double)r[jj] + (double)r[jj - 1]) / (g.dy * g.dy);
}

__device__ __forceinline__ double gxy(const double *f, const Geo &g, int i, int j)
{
    int jj = min(max(j, 1), g.ny - 2);
    int ip = wrapi(i + 1, g.nx), im = wrapi(i - 1, g.nx);
    const double *rp = f + (size_t)ip * g.ny, *rm = f + (size_t)im * g.ny;
    return (rp[jj + 1] - rp[jj - 1] - rm[jj + 1] + rm[jj - 1]) / (4.0 * g.dx * g.dy);
}

// float32 input: the differences are formed in float32 (bs.py:334-335 pass self.u/self.v uncast)
__device__ __forceinline__ double gxy_f32(const float *f, const Geo &g, int i, int j)
{
    int jj = min(max(j, 1), g.ny - 2);
    int ip = wrapi(i + 1, g.nx), im = wrapi(i - 1, g.nx);
    const float *rp = f + (size_t)ip * g.ny, *rm = f + (size_t)im * g.ny;
    float d = rp[jj + 1] - rp[jj - 1];
    d = d - rm[jj + 1];
    d = d + rm[jj - 1];
    return (double)d / (4.0 * g.dx * g.dy);
}

// bs.py:291-305; scipy.ndimage.convolve accumulates the 3x3 products in row-major kernel order
__device__ __forceinline__ double smth9(const double *f, const Geo &g, int i, int j)
{
    double c = f[(size_t)i * g.ny + j];
    if (i < 1 || i > g.nx - 3 || j < 1 || j > g.ny - 3) return c;
    const double kq = 0.25 / 4, kp = 0.5 / 4, kc = -(0.5 + 0.25);
    const double *r0 = f + (size_t)(i - 1) * g.ny + j, *r1 = f + (size_t)i * g.ny + j, *r2 = f + (size_t)(i + 1) * g.ny + j;
    double s = 0.0;
    s += kq * r0[-1];
    s += kp * r0[0];
    s += kq * r0[1];
    s += kp * r1[-1];
    s += kc * r1[0];
    s += kp * r1[1];
    s += kq * r2[-1];
    s += kp * r2[0];
    s += kq * r2[1];
    return c + s;
}

__global__ void __launch_bounds__(256) bs_q_kernel(BsArgs a, Geo g)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, m = blockIdx.z;
    if (j >= g.ny) return;
    size_t mo = (size_t)m * g.nx * g.ny;
    const float *u = a.u + mo, *v = a.v + mo;
    int jj = min(max(j, 1), g.ny - 2);  // bs.py:278-279: rows 0 / -1 copy rows 1 / -2
    double vx = gx(v, g, i, jj);
    const float *ur = u + (size_t)i * g.ny;
    double ucy = ((double)ur[jj + 1] * a.coslat[jj + 1] - (double)ur[jj - 1] * a.coslat[jj - 1]) / (2.0 * g.dy);
    a.q[mo + (size_t)i * g.ny + j] = (vx - ucy) / a.coslat[jj] + 2.0 * OMEGA * a.sinlat[jj] * REARTH;  // bs.py:276-277
}

__global__ void __launch_bounds__(256) bs_d2_kernel(BsArgs a, Geo g)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, m = blockIdx.z;
    if (j >= g.ny) return;
    size_t mo = (size_t)m * g.nx * g.ny, o = mo + (size_t)i * g.ny + j;
    const double *q = a.q + mo;
    a.qxx[o] = gxx(q, g, i, j);
    a.qyy[o] = gyy(q, g, i, j);
    a.qxy[o] = gxy(q, g, i, j);
}

__global__ void __launch_bounds__(256) bs_pack_kernel(BsArgs a, Geo g)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, m = blockIdx.z;
    if (j >= g.ny) return;
    size_t mo = (size_t)m * g.nx * g.ny;
    const float *u = a.u + mo, *v = a.v + mo;
    const double *q = a.q + mo, *qxx = a.qxx + mo, *qyy = a.qyy + mo, *qxy = a.qxy + mo;
    double f[18];
    f[0] = (double)u[(size_t)i * g.ny + j];
    f[1] = (double)v[(size_t)i * g.ny + j];
    f[2] = gx(u, g, i, j);
    f[3] = gy(u, g, i, j);
    f[4] = gx(v, g, i, j);
    f[5] = gy(v, g, i, j);
    f[6] = gx(q, g, i, j);
    f[7] = gy(q, g, i, j);
    f[8] = smth9(qxx, g, i, j);              // bs.py:345-347: smoothed AFTER the third derivatives were taken
    f[9] = smth9(qxy, g, i, j);
    f[10] = qxy[(size_t)i * g.ny + j];       // qyx = copy of the unsmoothed qxy (bs.py:337)
    f[11] = smth9(qyy, g, i, j);
    f[12] = gx(qxx, g, i, j);                // qxxx
    f[13] = gy(qxx, g, i, j);                // qxxy
    f[14] = gy(qxy, g, i, j);                // qxyy
    f[15] = gy(qyy, g, i, j);                // qyyy
    f[16] = gx(qxy, g, i, j);                // qyxx
    f[17] = gx(qyy, g, i, j);                // qyyx
    int nxf = g.nx + (a.xcyclic ? 1 : 0);
    for (int rep = 0; rep < 2; ++rep) {
        int io = (rep == 0) ? i : g.nx;  // wrap column: copy of column 0 (bs.py:370-372)
        if (rep == 1 && !(a.xcyclic && i == 0)) break;
        size_t po = ((size_t)m * nxf + io) * g.ny + j;
        double2 *pk = reinterpret_cast<double2 *>(a.fpk + po * NPK);
        pk[0] = make_double2(f[0], f[1]);
        pk[1] = make_double2(f[2], f[3]);
        pk[2] = make_double2(f[4], f[5]);
        pk[3] = make_double2(f[6], f[7]);
        pk[4] = make_double2(f[8], f[9]);
        pk[5] = make_double2(f[11], 0.0);
        if (a.fields18) {
            double2 *o = reinterpret_cast<double2 *>(a.fields18 + po * 18);
#pragma unroll
            for (int k = 0; k < 9; ++k) o[k] = make_double2(f[2 * k], f[2 * k + 1]);
        }
    }
    if (a.diag9) {
        size_t n = (size_t)g.nx * g.ny, o = (size_t)i * g.ny + j;
        double *d = a.diag9 + (size_t)m * 9 * n;
        double uyy = gyy(u, g, i, j);
        d[0 * n + o] = q[o];
        d[1 * n + o] = gxx(u, g, i, j);
        d[2 * n + o] = uyy;
        d[3 * n + o] = gxx(v, g, i, j);
        d[4 * n + o] = gyy(v, g, i, j);
        d[5 * n + o] = gxy_f32(u, g, i, j);
        d[6 * n + o] = gxy_f32(v, g, i, j);
        double bm = qnan(), ks = qnan();
        if (j > 0 && j < g.ny - 1) {
            double cl = a.coslat[j], sl = a.sinlat[j], uu = f[0];
            bm = (2 * OMEGA * (cl * cl) + (-cl * uyy + sl * f[3] + uu / cl) / REARTH) / REARTH;  // bs.py:379-383
            if (bm > 0 && uu > 0) ks = sqrt(bm * cl / uu) * REARTH;                                // bs.py:393-404
        }
        d[7 * n + o] = bm;
        d[8 * n + o] = ks;
    }
}

__global__ void trig_kernel(int n, const double *lat, double *c, double *s)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) {
        c[j] = cos(lat[j]);
        s[j] = sin(lat[j]);
    }
}

__global__ void pack18_kernel(size_t npts, const double *f18, double *fpk)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const double *s = f18 + i * 18;
    double *d = fpk + i * NPK;
#pragma unroll
    for (int k = 0; k < 10; ++k) d[k] = s[k];
    d[10] = s[11];
    d[11] = 0.0;
}

// The 12th double of a packed point: the 11th field (qyy) of its latitude neighbour (i, j + 1), so that the sixth
// 16-byte piece of a cell's (x0, y0) point holds that field for BOTH of its latitude corners and the cooperative
// integrators read it with one LDS.128 (rwrt_rk45.cuh, RWRT_PAD_F10).  The last latitude row repeats its own value
// (its neighbour index is clamped onto it, interpolation.py:116-119).  Nothing else reads this slot.
__global__ void pack_pad_kernel(size_t npts, int nlat, double *fpk)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const int j = (int)(i % (size_t)nlat);
    fpk[i * NPK + 11] = fpk[(j + 1 < nlat ? i + 1 : i) * NPK + 10];
}

}  // namespace

void launch_bs_ready(const BsArgs &a, cudaStream_t s)
{
    Geo g{a.nlon, a.nlat, 2.0 * PI / a.nlon, PI / (a.nlat - 1)};  // bs.py:77-78
    dim3 grid((a.nlat + 255) / 256, a.nlon, a.nmember);
    bs_q_kernel<<<grid, 256, 0, s>>>(a, g);
    bs_d2_kernel<<<grid, 256, 0, s>>>(a, g);
    bs_pack_kernel<<<grid, 256, 0, s>>>(a, g);
    const size_t npts = (size_t)a.nmember * (a.nlon + (a.xcyclic ? 1 : 0)) * a.nlat;
    pack_pad_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, s>>>(npts, a.nlat, a.fpk);
}

void launch_trig_tables(int nlat, const double *lat, double *coslat, double *sinlat, cudaStream_t s)
{
    trig_kernel<<<(nlat + 127) / 128, 128, 0, s>>>(nlat, lat, coslat, sinlat);
}

void launch_pack_from_fields18(int nmember, int nlon_f, int nlat, const double *fields18, double *fpk, cudaStream_t s)
{
    size_t npts = (size_t)nmember * nlon_f * nlat;
    pack18_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, s>>>(npts, fields18, fpk);
    pack_pad_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, s>>>(npts, nlat, fpk);
}

}  // namespace rwrt
