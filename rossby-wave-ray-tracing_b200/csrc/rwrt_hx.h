// rwrt_hx.h -- host half of rwrt_real2d_host: ray records (64 bytes per live ray and output row, the layout the
// integrator writes) -> the reference's seven planes (nt, 3, nsource, nzwn) (wr.py:767-781, 856-885) in the caller's
// host memory.  Plain C++ (no CUDA): the same arithmetic-free transpose expand_records does on the device
// (rwrt_rk45.cuh), so that only the records (13 GB on C2) cross the link instead of the expanded rows (23 GB).
#pragma once
#include <cstddef>
#include <cstdint>
#include <immintrin.h>

namespace rwrt {

struct HxView {
    double *out;                // (7, nt, N) host array; row 0 (the t = 0 state) is already there
    size_t pl;                  // doubles between planes (nt * N)
    long long N;                // slots
    const int32_t *inv;         // slot -> live ray index, < 0 for NaN-root slots
    const double *stage;        // records of this window: [live ray][rec_rows][8]
    const int32_t *rows_done;   // per live ray: rows 0 .. rows_done-1 exist, later rows are NaN (dead ray)
    int rec_rows;               // rows per ray in `stage`
};

#ifndef RWRT_HX_GROUP
#define RWRT_HX_GROUP 32
#endif
#ifndef RWRT_HX_TROWS
#define RWRT_HX_TROWS 16
#endif
constexpr int HX_GROUP = RWRT_HX_GROUP;    // slots per tile: 7 planes x 16 rows x 32 slots x 8 B = 28 KB of L1
constexpr int HX_TROWS = RWRT_HX_TROWS;    // rows per tile pass

static inline void hx_stream_line(double *dst, const double *src, int n)
{
    // n doubles to dst with non-temporal stores where 16-byte alignment allows
    int j = 0;
    if (((uintptr_t)dst & 8) && n > 0) dst[j] = src[j], ++j;
#if defined(HX_TRY_AVX512)
    for (; j + 2 <= n && ((uintptr_t)(dst + j) & 63); j += 2) _mm_stream_pd(dst + j, _mm_loadu_pd(src + j));
    for (; j + 8 <= n; j += 8) _mm512_stream_pd(dst + j, _mm512_loadu_pd(src + j));
#elif defined(HX_TRY_AVX2)
    for (; j + 2 <= n && ((uintptr_t)(dst + j) & 31); j += 2) _mm_stream_pd(dst + j, _mm_loadu_pd(src + j));
    for (; j + 4 <= n; j += 4) _mm256_stream_pd(dst + j, _mm256_loadu_pd(src + j));
#endif
    for (; j + 2 <= n; j += 2) _mm_stream_pd(dst + j, _mm_loadu_pd(src + j));
    if (j < n) dst[j] = src[j];
}

// Rows [b, e) of slots [s0, s1).  Row `it` of a ray sits at stage[(li * rec_rows + it - b) * 8 .. + 7].
static inline void hx_expand_rows(const HxView &v, int b, int e, long long s0, long long s1)
{
    const double nan = __builtin_nan("");
    alignas(64) double tile[HX_TROWS][7][HX_GROUP];
    for (long long g0 = s0; g0 < s1; g0 += HX_GROUP) {
        const int n = (int)((s1 - g0) < HX_GROUP ? (s1 - g0) : HX_GROUP);
        for (int r0 = b; r0 < e; r0 += HX_TROWS) {
            const int nr = (e - r0) < HX_TROWS ? (e - r0) : HX_TROWS;
            // gather: a ray's rows are contiguous, read them in order
            for (int j = 0; j < n; ++j) {
                const int li = v.inv[g0 + j];
                if (li < 0) {       // NaN-root slot: planes 0..4 keep their t = 0 value, u_g / v_g are NaN (wr.py:768-781)
                    double c0[5];
                    for (int c = 0; c < 5; ++c) c0[c] = v.out[(size_t)c * v.pl + g0 + j];
                    for (int r = 0; r < nr; ++r) {
                        for (int c = 0; c < 5; ++c) tile[r][c][j] = c0[c];
                        tile[r][5][j] = nan;
                        tile[r][6][j] = nan;
                    }
                } else {
                    const int rd = v.rows_done[li];
                    const double *rec = v.stage + ((size_t)li * v.rec_rows + (r0 - b)) * 8;
                    int r = 0;
                    for (; r < nr && r0 + r < rd; ++r)
                        for (int c = 0; c < 7; ++c) tile[r][c][j] = rec[r * 8 + c];
                    for (; r < nr; ++r)
                        for (int c = 0; c < 7; ++c) tile[r][c][j] = nan;
                }
            }
            // scatter: n consecutive doubles per (plane, row)
            for (int r = 0; r < nr; ++r)
                for (int c = 0; c < 7; ++c)
                    hx_stream_line(v.out + (size_t)c * v.pl + (size_t)(r0 + r) * v.N + g0, tile[r][c], n);
        }
    }
    _mm_sfence();
}

}  // namespace rwrt
