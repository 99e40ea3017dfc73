// rwrt_shsf.cu -- the reference's optional background pre-filter SHSF (SHSF.py:10-33):
//     SHSF(data, L) = MakeGridDH( pad( SHExpandDH(data, sampling), L ) )
// i.e. project a Driscoll-Healy grid onto the spherical harmonics of degree <= L and evaluate the result on
// the (2L+2)-row DH grid.  The reference delegates the arithmetic to the un-vendored package pyshtools
// (parity unpinned); what is implemented is the published algorithm behind those calls: DH (1994)
// quadrature weights, 4-pi normalised real harmonics without Condon-Shortley phase (pyshtools defaults).
//
// B200 mapping: the filter is linear, so for a batch of nmember grids (the 1000-member ensemble of
// BASELINE configs[3] is the case where it matters) it is four FP64 matrix products with plan-time tables:
//   1. longitude analysis      F[mem][j][c]   = sum_k  data[mem][j][k] * Cdft[k][c]        c = r*nl + m  (r: cos / sin)
//   2. Legendre analysis       Coef[c][l][mem] = sum_j  A[m][l][j]     * F[mem][j][c]       A = w_j Pbar_lm(cos theta_j)
//   3. Legendre synthesis      G[c][i][mem]    = sum_l  S[m][i][l]     * Coef[c][l][mem]    S = Pbar_lm(cos theta'_i)
//   4. longitude synthesis     out[mem][i][k'] = sum_c  G[c][i][mem]   * Cinv[c][k']
// all through one strided, batched, shared-memory-tiled DFMA kernel (FP64 has no tcgen05 path; the fields are
// fp64 by contract).  Tables are computed on the host in extended precision when the plan is created.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/rwrt.h"

namespace rwrt {
void set_error_string(const std::string &s);  // rwrt_api.cu

namespace {

// C[b][i][j] = sum_k A[b'][i][k] * B[b][k][j] with arbitrary element strides; b' = b % a_bmod when a_bmod > 0
struct GemmArgs {
    const double *A, *B;
    double *C;
    int M, N, K, batch, a_bmod;
    long long sa_i, sa_k, sa_b, sb_k, sb_j, sb_b, sc_i, sc_j, sc_b;
};

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) gemm_f64_strided(GemmArgs g)
{
    __shared__ double As[TK][TM + 1];
    __shared__ double Bs[TK][TN + 1];
    const int b = blockIdx.z;
    const int ab = g.a_bmod > 0 ? b % g.a_bmod : b;
    const double *A = g.A + (long long)ab * g.sa_b;
    const double *B = g.B + (long long)b * g.sb_b;
    double *C = g.C + (long long)b * g.sc_b;
    const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
    double acc[4][4] = {};
    for (int k0 = 0; k0 < g.K; k0 += TK) {
        // A tile: TM x TK = 1024 elements, 4 per thread; consecutive threads along whichever index is contiguous
        for (int e = threadIdx.x; e < TM * TK; e += 256) {
            int ii, kk;
            if (g.sa_k == 1) { kk = e % TK; ii = e / TK; } else { ii = e % TM; kk = e / TM; }
            int gi = i0 + ii, gk = k0 + kk;
            As[kk][ii] = (gi < g.M && gk < g.K) ? A[gi * g.sa_i + gk * g.sa_k] : 0.0;
        }
        for (int e = threadIdx.x; e < TK * TN; e += 256) {
            int jj, kk;
            if (g.sb_k == 1) { kk = e % TK; jj = e / TK; } else { jj = e % TN; kk = e / TN; }
            int gj = j0 + jj, gk = k0 + kk;
            Bs[kk][jj] = (gj < g.N && gk < g.K) ? B[gk * g.sb_k + gj * g.sb_j] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            double a[4], bb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = As[kk][ty + 16 * r];
#pragma unroll
            for (int c = 0; c < 4; ++c) bb[c] = Bs[kk][tx + 16 * c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], bb[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int gi = i0 + ty + 16 * r;
        if (gi >= g.M) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int gj = j0 + tx + 16 * c;
            if (gj < g.N) C[gi * g.sc_i + gj * g.sc_j] = acc[r][c];
        }
    }
}

// The same product on the FP64 tensor-core path: mma.sync m8n8k4 (DMMA).  The CUDA-core kernel above needs one LDS.64
// per two DFMA (4 x 4 outputs per thread) and reached 25-33 % of the DFMA peak, bound by the shared-memory pipe; here a
// warp owns a 32 x 32 tile = 4 x 4 fragments of 8 x 8, and every k-step of 4 costs each thread 8 LDS.64 for 16 DMMA
// (= 128 FMA per thread): 1/8 of the shared-memory traffic per flop, 1/8 of the issue slots.  Same arguments, strides and
// tile loads; fragments come straight from the k-major tiles, whose row stride (tile + 4 doubles) makes the
// (lane & 3, lane >> 2) access pattern of a half-warp hit 16 distinct 8-byte banks.  (tcgen05 has no FP64 kind: for
// double precision mma.sync IS the tensor-core path on sm_100a.)
constexpr int MPAD = 4;
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(128) gemm_f64_mma(GemmArgs g)
{
    // two tile buffers: the next k-tile is fetched into registers before the current one is multiplied and stored to the
    // other buffer afterwards -- global-load latency hides behind 64 DMMA per warp, one __syncthreads per k-tile
    __shared__ double As[2][TK][TM + MPAD];
    __shared__ double Bs[2][TK][TN + MPAD];
    const int b = blockIdx.z;
    const int ab = g.a_bmod > 0 ? b % g.a_bmod : b;
    const double *A = g.A + (long long)ab * g.sa_b;
    const double *B = g.B + (long long)b * g.sb_b;
    double *C = g.C + (long long)b * g.sc_b;
    const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;          // this warp's 32 x 32 corner of the 64 x 64 tile
    const int fk = lane & 3, fr = lane >> 2;                         // fragment coordinates of this lane
    // this thread's 8 + 8 tile elements e = tid + 128 r: consecutive threads run along whichever index is contiguous in
    // memory, so (row, k) of element r is (tid / 16 + 8 r, tid % 16) or (tid % 64, tid / 64 + 2 r) -- one base pointer and
    // one stride per operand, everything else is recomputed from r (keeps the kernel at 4 blocks per SM)
    constexpr int EPT = TM * TK / 128;
    const bool a_kfast = (g.sa_k == 1), b_kfast = (g.sb_k == 1);
    const int t16 = threadIdx.x & 15, t64 = threadIdx.x & 63, q16 = threadIdx.x >> 4, q64 = threadIdx.x >> 6;
    const double *pa = a_kfast ? A + (long long)(i0 + q16) * g.sa_i + t16 : A + (long long)(i0 + t64) * g.sa_i + (long long)q64 * g.sa_k;
    const double *pb = b_kfast ? B + (long long)(j0 + q16) * g.sb_j + t16 : B + (long long)(j0 + t64) * g.sb_j + (long long)q64 * g.sb_k;
    const long long sta = a_kfast ? 8 * g.sa_i : 2 * g.sa_k, stb = b_kfast ? 8 * g.sb_j : 2 * g.sb_k;
    double ra[EPT], rb[EPT];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int ai = a_kfast ? q16 + 8 * r : t64, ak = a_kfast ? t16 : q64 + 2 * r;
            const int bj = b_kfast ? q16 + 8 * r : t64, bk = b_kfast ? t16 : q64 + 2 * r;
            ra[r] = (i0 + ai < g.M && k0 + ak < g.K) ? pa[r * sta + (long long)k0 * g.sa_k] : 0.0;
            rb[r] = (j0 + bj < g.N && k0 + bk < g.K) ? pb[r * stb + (long long)k0 * g.sb_k] : 0.0;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int ai = a_kfast ? q16 + 8 * r : t64, ak = a_kfast ? t16 : q64 + 2 * r;
            const int bj = b_kfast ? q16 + 8 * r : t64, bk = b_kfast ? t16 : q64 + 2 * r;
            As[buf][ak][ai] = ra[r];
            Bs[buf][bk][bj] = rb[r];
        }
    };
    double acc[4][4][2] = {};
    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < g.K; k0 += TK) {
        const bool more = k0 + TK < g.K;
        if (more) fetch(k0 + TK);
#pragma unroll
        for (int ks = 0; ks < TK; ks += 4) {
            double a[4], bb[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = As[buf][ks + fk][wm + 8 * mi + fr];      // A fragment: row fr, column fk of the 8 x 4 block
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) bb[ni] = Bs[buf][ks + fk][wn + 8 * ni + fr];     // B fragment: row fk, column fr of the 4 x 8 block
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bb[ni]);
        }
        if (more) stash(buf ^ 1);        // the other buffer: its last readers passed the barrier of the previous iteration
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int gi = i0 + wm + 8 * mi + fr;                       // C fragment: row fr, columns 2 fk and 2 fk + 1
        if (gi >= g.M) continue;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int gj = j0 + wn + 8 * ni + 2 * fk + h;
                if (gj < g.N) C[gi * g.sc_i + gj * g.sc_j] = acc[mi][ni][h];
            }
    }
}

int launch_gemm(const GemmArgs &g, cudaStream_t s)
{
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return 0;
    dim3 grid((g.N + TN - 1) / TN, (g.M + TM - 1) / TM, g.batch);
    static const bool dfma = [] { const char *e = getenv("RWRT_SHSF_GEMM"); return e && !strcmp(e, "dfma"); }();   // cross-check path
    if (dfma) gemm_f64_strided<<<grid, 256, 0, s>>>(g);
    else gemm_f64_mma<<<grid, 128, 0, s>>>(g);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

typedef long double ld;

// Driscoll-Healy weights for int_0^pi g sin(theta) dtheta on theta_j = pi j / n (exact for degree < n; sum = 2)
std::vector<ld> dh_weights(int n)
{
    const ld pi = acosl(-1.0L);
    std::vector<ld> w(n);
    for (int j = 0; j < n; ++j) {
        ld th = pi * j / n, s = 0;
        for (int k = 0; k < n / 2; ++k) s += sinl((2 * k + 1) * th) / (2 * k + 1);
        w[j] = 4.0L / n * sinl(th) * s;
    }
    return w;
}

// 4-pi normalised Pbar_lm(cos theta_i), no Condon-Shortley phase, for one order m and l = m..lmax: out[l - m][i]
void plm_bar_order(int m, int lmax, const std::vector<ld> &x, std::vector<std::vector<ld>> &out)
{
    const size_t nx = x.size();
    out.assign(lmax - m + 1, std::vector<ld>(nx));
    for (size_t i = 0; i < nx; ++i) {
        ld s = sqrtl(fmaxl(1.0L - x[i] * x[i], 0.0L));
        ld pmm = 1.0L;
        for (int q = 1; q <= m; ++q) pmm *= (q == 1 ? sqrtl(3.0L) : sqrtl((ld)(2 * q + 1) / (ld)(2 * q))) * s;
        out[0][i] = pmm;
        if (m + 1 <= lmax) out[1][i] = sqrtl((ld)(2 * m + 3)) * x[i] * pmm;
        for (int l = m + 2; l <= lmax; ++l) {
            ld a = sqrtl((ld)(4.0L * l * l - 1) / (ld)((ld)l * l - (ld)m * m));
            ld b = sqrtl((ld)(((ld)(l - 1)) * (l - 1) - (ld)m * m) / (ld)(4.0L * (l - 1) * (l - 1) - 1));
            out[l - m][i] = a * (x[i] * out[l - m - 1][i] - b * out[l - m - 2][i]);
        }
    }
}

}  // namespace
}  // namespace rwrt

using namespace rwrt;

struct rwrt_shsf_plan_s {
    int device = 0, n_in = 0, nlon_in = 0, L = 0, Lc = 0, nl = 0, n_out = 0, nlon_out = 0;
    double *cdft = nullptr, *atab = nullptr, *stab = nullptr, *cinv = nullptr;  // device tables
    double *work = nullptr;                                                     // F, Coef, G for `work_members` grids
    int work_members = 0;
};

#define SH_CU(call)                                                                         \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            set_error_string(std::string(#call) + ": " + cudaGetErrorString(e_));           \
            return RWRT_ERR_CUDA;                                                           \
        }                                                                                   \
    } while (0)

extern "C" int rwrt_shsf_plan_create(int device, int nlat, int in_sampling, int truncation, int out_sampling,
                                     rwrt_shsf_plan *out)
{
    if (!out) return RWRT_ERR_ARG;
    *out = nullptr;
    if (nlat < 2 || (nlat & 1) || truncation < 0 || (in_sampling != 1 && in_sampling != 2) ||
        (out_sampling != 1 && out_sampling != 2)) {
        set_error_string("rwrt_shsf_plan_create: a Driscoll-Healy grid has an even number of latitudes n and n or 2n "
                         "longitudes; truncation >= 0");
        return RWRT_ERR_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        set_error_string("rwrt_shsf_plan_create: no CUDA device (there is no CPU fallback)");
        return RWRT_ERR_CUDA;
    }
    SH_CU(cudaSetDevice(device));
    auto *p = new rwrt_shsf_plan_s;
    p->device = device;
    p->n_in = nlat;
    p->nlon_in = in_sampling * nlat;
    p->L = truncation;
    p->Lc = truncation < nlat / 2 - 1 ? truncation : nlat / 2 - 1;  // SHExpandDH resolves degrees <= n/2 - 1; pad() zero-fills above
    p->nl = p->Lc + 1;
    p->n_out = 2 * truncation + 2;
    p->nlon_out = out_sampling * p->n_out;
    const int nl = p->nl, cw = 2 * nl, N = p->n_in, n = p->n_out;
    const ld pi = acosl(-1.0L);
    // longitude tables: c = r * nl + m
    std::vector<double> cdft((size_t)p->nlon_in * cw), cinv((size_t)cw * p->nlon_out);
    for (int k = 0; k < p->nlon_in; ++k)
        for (int m = 0; m < nl; ++m) {
            ld ph = 2 * pi * ((long long)m * k % p->nlon_in) / p->nlon_in;
            // (1 / 4pi) * (2pi / nlon): normalisation of c_lm = (1/4pi) int f Y_lm dOmega
            cdft[(size_t)k * cw + m] = (double)(cosl(ph) / (2.0L * p->nlon_in));
            cdft[(size_t)k * cw + nl + m] = (double)(sinl(ph) / (2.0L * p->nlon_in));
        }
    for (int m = 0; m < nl; ++m)
        for (int k = 0; k < p->nlon_out; ++k) {
            ld ph = 2 * pi * ((long long)m * k % p->nlon_out) / p->nlon_out;
            cinv[(size_t)m * p->nlon_out + k] = (double)cosl(ph);
            cinv[(size_t)(nl + m) * p->nlon_out + k] = (double)sinl(ph);
        }
    // Legendre tables
    std::vector<ld> xin(N), xout(n), w = dh_weights(N);
    for (int j = 0; j < N; ++j) xin[j] = cosl(pi * j / N);
    for (int i = 0; i < n; ++i) xout[i] = cosl(pi * i / n);
    std::vector<double> atab((size_t)nl * nl * N, 0.0), stab((size_t)nl * n * nl, 0.0);
    std::vector<std::vector<ld>> P;
    for (int m = 0; m < nl; ++m) {
        plm_bar_order(m, p->Lc, xin, P);
        for (int l = m; l <= p->Lc; ++l)
            for (int j = 0; j < N; ++j) atab[((size_t)m * nl + l) * N + j] = (double)(w[j] * P[l - m][j]);
        plm_bar_order(m, p->Lc, xout, P);
        for (int l = m; l <= p->Lc; ++l)
            for (int i = 0; i < n; ++i) stab[((size_t)m * n + i) * nl + l] = (double)P[l - m][i];
    }
    auto up = [&](double **d, const std::vector<double> &h) -> int {
        if (cudaMalloc((void **)d, h.size() * sizeof(double)) != cudaSuccess) return -1;
        return cudaMemcpy(*d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1;
    };
    if (up(&p->cdft, cdft) || up(&p->cinv, cinv) || up(&p->atab, atab) || up(&p->stab, stab)) {
        cudaGetLastError();
        set_error_string("rwrt_shsf_plan_create: out of device memory for the transform tables");
        rwrt_shsf_plan_destroy(p);
        return RWRT_ERR_NOMEM;
    }
    *out = p;
    return RWRT_OK;
}

extern "C" int rwrt_shsf_plan_destroy(rwrt_shsf_plan p)
{
    if (!p) return RWRT_OK;
    cudaSetDevice(p->device);
    cudaFree(p->cdft);
    cudaFree(p->cinv);
    cudaFree(p->atab);
    cudaFree(p->stab);
    cudaFree(p->work);
    delete p;
    return RWRT_OK;
}

extern "C" int rwrt_shsf_out_shape(rwrt_shsf_plan p, int *nlat_out, int *nlon_out)
{
    if (!p || !nlat_out || !nlon_out) return RWRT_ERR_ARG;
    *nlat_out = p->n_out;
    *nlon_out = p->nlon_out;
    return RWRT_OK;
}

extern "C" int rwrt_shsf_apply(rwrt_shsf_plan p, const double *grid, int nmember, double *out, void *stream)
{
    if (!p || !grid || !out || nmember < 0) {
        set_error_string("rwrt_shsf_apply: bad argument");
        return RWRT_ERR_ARG;
    }
    if (nmember == 0) return RWRT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    SH_CU(cudaSetDevice(p->device));
    const long long nl = p->nl, cw = 2 * nl, N = p->n_in, n = p->n_out, nm = nmember;
    const size_t need = (size_t)(nm * N * cw + cw * nl * nm + cw * n * nm);
    if (nmember > p->work_members) {
        SH_CU(cudaStreamSynchronize(s));
        cudaFree(p->work);
        p->work = nullptr;
        p->work_members = 0;
        if (cudaMalloc((void **)&p->work, need * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            set_error_string("rwrt_shsf_apply: out of device memory for the work arrays");
            return RWRT_ERR_NOMEM;
        }
        p->work_members = nmember;
    }
    double *F = p->work, *Coef = F + nm * N * cw, *G = Coef + cw * nl * nm;
    GemmArgs g;
    // 1. F[mem][j][c] = data[mem][j][:] . Cdft[:][c]
    g = GemmArgs{grid, p->cdft, F, (int)N, (int)cw, p->nlon_in, nmember, 0, p->nlon_in, 1, N * p->nlon_in, cw, 1, 0, cw, 1, N * cw};
    if (launch_gemm(g, s)) goto fail;
    // 2. Coef[c][l][mem] = A[m][l][:] . F[mem][:][c]
    g = GemmArgs{p->atab, F, Coef, (int)nl, nmember, (int)N, (int)cw, (int)nl, N, 1, nl * N, cw, N * cw, 1, nm, 1, nl * nm};
    if (launch_gemm(g, s)) goto fail;
    // 3. G[c][i][mem] = S[m][i][:] . Coef[c][:][mem]
    g = GemmArgs{p->stab, Coef, G, (int)n, nmember, (int)nl, (int)cw, (int)nl, nl, 1, n * nl, nm, 1, nl * nm, nm, 1, n * nm};
    if (launch_gemm(g, s)) goto fail;
    // 4. out[mem][i][k'] = G[:][i][mem] . Cinv[:][k']
    g = GemmArgs{G, p->cinv, out, (int)n, p->nlon_out, (int)cw, nmember, 0, nm, n * nm, 1, p->nlon_out, 1, 0, p->nlon_out, 1, n * p->nlon_out};
    if (launch_gemm(g, s)) goto fail;
    return RWRT_OK;
fail:
    set_error_string(std::string("rwrt_shsf_apply: kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    return RWRT_ERR_CUDA;
}
