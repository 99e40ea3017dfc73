// rwrt_hx.h -- host half of rwrt_real2d_host: the packed rows of one time window (seven planes of LIVE rays only,
// what pack_records writes on the device) -> the reference's seven planes (nt, 3, nsource, nzwn) (wr.py:767-781,
// 856-885) in the caller's host memory.  Plain C++ (rwrt_hx.cpp, no CUDA), no arithmetic: slots without a real root
// repeat their t = 0 values (u_g, v_g NaN), live slots take the value of their ray.  Only 56 bytes per live ray and
// row cross the link (11.6 GB on C2) instead of the expanded rows (23.2 GB).
#pragma once
#include <cstddef>
#include <cstdint>

namespace rwrt {

struct HxView {
    double *out;                // (7, nt, N) host array; row 0 (the t = 0 state) is already there
    size_t pl;                  // doubles between planes (nt * N)
    long long N;                // slots
    const int32_t *inv;         // slot -> live ray index, < 0 for slots without a root
    const double *packed;       // [7][rows][nray]: plane, row of this window, live ray (NaN once a ray is dead)
    long long nray;             // live rays (>= 1)
    int rows;                   // rows of this window in `packed`
};

#ifndef RWRT_HX_CHUNK
#define RWRT_HX_CHUNK 4096
#endif
constexpr int HX_CHUNK = RWRT_HX_CHUNK;  // slots per task: slot map 16 KB + constants 192 KB stay in L2 across the 7 x rows passes

// Rows [b, b + v.rows) of slots [s0, s1), s1 - s0 <= HX_CHUNK.  Non-temporal stores, as wide as the CPU has them
// (RWRT_HX_ISA=sse2 | avx2 | avx512 forces a narrower path).
void hx_expand_rows(const HxView &v, int b, long long s0, long long s1);
const char *hx_isa_name();

}  // namespace rwrt
