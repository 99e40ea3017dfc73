// rwrt_rk45_params.h -- kernel parameter block of the RK45 kernels (shared by both builds and the API).
#pragma once
#include "rwrt_device.cuh"

namespace rwrt {

constexpr int RWRT_STATE = 16;  // doubles per live ray in the state array (128 B)

struct RunParams {
    Grid g;
    const int32_t *src_member;  // (nsource) or nullptr
    int nsource, nzwn, nt;
    long long N;                // slots = 3 * nsource * nzwn
    double tstep, ttotal, rtol, atol, gmin, cut;
    double *out;                // 7 planes, each (nt, N)
    size_t out_pl;              // doubles between the planes of `out` (nt * N; N when only row 0 exists: host-side expansion)
    double *dst;                // expand_records writes row `it` of plane c to dst[c * dst_pl + (it - dst_it0) * N + slot]:
    size_t dst_pl;              // normally `out` itself (dst_pl = nt * N, dst_it0 = 0); a window-sized buffer when the rows
    int dst_it0;                // of one time window are expanded for a direct copy to the host (rwrt_real2d_host)
    const int32_t *live;        // compacted list of live slots
    long long nlive;
    double *state;              // (nlive, RWRT_STATE): y[5], f[5], h_abs, cos(prev lat), t, prev lon, prev lat, counters
    int coop;                   // 1: cooperative-gather integrator (default); 0: per-lane gather (cross-check)
    // optional record layout (RWRT_RECORDS=1): the integrator writes one 64-byte record per (ray, output row) and
    // expand_records turns them into the reference layout afterwards (no partial-sector stores)
    double *rec;                // (nlive, rec_rows, 8) or nullptr
    int rec_rows;               // rows per ray in `rec` (= rows of the largest time window)
    int rec_it0;                // record row r of a ray holds output row rec_it0 + r (= it_begin of the window that wrote it)
    int32_t *rows_done;         // (nlive): number of valid rows 0..rows_done-1 of a ray so far
    const int32_t *inv;         // (N): slot -> index in the live list, -1 for NaN-root slots
    int it_begin, it_end;       // this launch produces output rows [it_begin, it_end) (time window)
    int nquanta, qshift;        // the cooperative kernel integrates a ray in nquanta time quanta (work items); quantum q ends at row
                                // it_begin + len - (len >> qshift (q + 1)), the last one at it_end
    long long *accepted;        // (N) or nullptr
    long long *attempts;        // (N) or nullptr
    signed char *status;        // (N) or nullptr
    unsigned long long *queue;  // work-fetch counter
    unsigned long long *total_accepted;
    // Gridding (rwrt_ray_sweep_grid): the cooperative integrator adds every output sample to the ray-density / flux
    // grids as it produces it (three fire-and-forget integer RED per output row) -- or, with g_count == nullptr and
    // records on, grid_records_kernel consumes the window's records afterwards.
    unsigned long long *g_count, *g_fx, *g_fy;   // (g_ncx, g_ncy) each, or nullptr
    double g_dlon, g_dlat;      // cell size in degrees
    int g_ncx, g_ncy;
    int no_rows;                // 1: no output rows are stored at all (grid-only sweep: neither records nor planes)
};

// What one instantiation of the cooperative integrator does with an output row
constexpr int OUT_ROWS = 1;     // store it (64-byte record, or the seven planes directly)
constexpr int OUT_GRID = 2;     // add it to the grids

}  // namespace rwrt
