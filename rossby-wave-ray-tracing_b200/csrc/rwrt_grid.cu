// rwrt_grid.cu -- K4: ray-density / wave-ray-flux gridding of trajectories.
//
// The reference ships no gridding code (the manual's section 4 describes tools
// that are not in the repository), so the definition below is this project's own and its checker
// is oracle/gridding.py -- parity unpinned against the authors' tool:
//   domain  : three longitude circles [-360, 720) deg x [-90, 90] deg (rlon is never wrapped, wr.py:947)
//   cells   : ix = floor((lon_deg + 360) / dlon_deg), iy = min(floor((lat_deg + 90) / dlat_deg), ncy-1)
//   sample  : every output row it >= 1 of every slot whose lon, lat and l are finite
//   count   : int64 += 1;   flux_x / flux_y : int64 += rint(ug|vg * 2^20)  (fixed point: the sums are
//             order independent, so results are bit-identical for any sharding of the rays over GPUs)
// Compiled with -fmad=false so lon * rad2deg + 360 is two IEEE operations, as in the NumPy checker.
#include "rwrt_internal.h"

namespace rwrt {
namespace {

__global__ void __launch_bounds__(256) grid_kernel(long long nslot, int nt, const double *out, double dlon, double dlat,
                                                   int ncx, int ncy, unsigned long long *count, unsigned long long *fx,
                                                   unsigned long long *fy)
{
    long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nslot) return;
    const size_t pl = (size_t)nt * nslot;
    for (int it = blockIdx.y + 1; it < nt; it += gridDim.y) {      // gridDim.y is capped at 65535 rows per pass
        const size_t o = (size_t)it * nslot + slot;
        const long long c = grid_cell(out[o], out[pl + o], out[3 * pl + o], dlon, dlat, ncx, ncy);
        if (c >= 0) grid_add(count, fx, fy, c, out[5 * pl + o], out[6 * pl + o]);
    }
}

// The same accumulation from the per-ray records the integrator writes (rec[ray][row][8] = lon, lat, k, l, A, ug, vg,
// pad; rows [it_begin, min(it_end, rows_done[ray])) are valid): no (7, nt, N) planes are needed, so a sweep that only
// wants the grids never expands its records.  One thread per (ray, row): a warp reads 2 KB of contiguous records.
__global__ void __launch_bounds__(256) grid_records_kernel(const double *rec, int rec_rows, const int32_t *rows_done,
                                                           long long nlive, int it_begin, int it_end, double dlon,
                                                           double dlat, int ncx, int ncy, unsigned long long *count,
                                                           unsigned long long *fx, unsigned long long *fy)
{
    const int len = it_end - it_begin;
    const long long total = nlive * (long long)len;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long ray = i / len;
        const int r = (int)(i - ray * len);
        if (it_begin + r >= rows_done[ray]) continue;              // the ray died before this row (wr.py:838-855)
        const double2 *q = reinterpret_cast<const double2 *>(rec + ((size_t)ray * rec_rows + r) * 8);
        const double2 a = __ldcs(q), b = __ldcs(q + 1), c2 = __ldcs(q + 2), d = __ldcs(q + 3);
        const long long c = grid_cell(a.x, a.y, b.y, dlon, dlat, ncx, ncy);
        if (c >= 0) grid_add(count, fx, fy, c, c2.y, d.x);
    }
}

}  // namespace

void launch_grid_accumulate(long long nslot, int nt, const double *out7, double dlon_deg, double dlat_deg, int ncx,
                            int ncy, long long *count, long long *fx, long long *fy, cudaStream_t s)
{
    if (nslot <= 0 || nt < 2) return;
    dim3 grid((unsigned)((nslot + 255) / 256), (unsigned)(nt - 1 < 65535 ? nt - 1 : 65535));
    grid_kernel<<<grid, 256, 0, s>>>(nslot, nt, out7, dlon_deg, dlat_deg, ncx, ncy, (unsigned long long *)count,
                                     (unsigned long long *)fx, (unsigned long long *)fy);
}

void launch_grid_records(const double *rec, int rec_rows, const int32_t *rows_done, long long nlive, int it_begin, int it_end,
                         double dlon_deg, double dlat_deg, int ncx, int ncy, long long *count, long long *fx, long long *fy,
                         int sm_count, cudaStream_t s)
{
    const long long total = nlive * (long long)(it_end - it_begin);
    if (total <= 0) return;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count * 64;                // grid-stride beyond a few waves
    if (blocks > cap) blocks = cap;
    grid_records_kernel<<<(unsigned)blocks, 256, 0, s>>>(rec, rec_rows, rows_done, nlive, it_begin, it_end, dlon_deg, dlat_deg,
                                                          ncx, ncy, (unsigned long long *)count, (unsigned long long *)fx,
                                                          (unsigned long long *)fy);
}

}  // namespace rwrt

// ---------------------------------------------------------------------------------------------
// FP64 roofline denominator: a register-resident DFMA chain microbenchmark (8 independent chains
// per thread, enough warps to saturate the FP64 pipe).  Used by bench.py / profiles only.
// ---------------------------------------------------------------------------------------------
namespace rwrt {
namespace {
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *sink, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) sink[0] = s;  // keeps the chain alive without a store in practice
}
}  // namespace

int measure_fp64_peak(int sm_count, double *tflops, cudaStream_t s)
{
    double *sink = nullptr;
    if (cudaMalloc(&sink, 8) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, blocks = sm_count * 8, threads = 256;
    dfma_peak_kernel<<<blocks, threads, 0, s>>>(sink, 64, 0.999999, 1e-9);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, s);
        dfma_peak_kernel<<<blocks, threads, 0, s>>>(sink, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
}  // namespace rwrt
