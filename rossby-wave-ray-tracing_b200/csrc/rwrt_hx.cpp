// rwrt_hx.cpp -- see rwrt_hx.h.  Compiled by the host compiler alone; three store widths, chosen once at run time.
#include "rwrt_hx.h"

#include <cstdlib>
#include <cstring>
#include <immintrin.h>

namespace rwrt {
namespace {

// value of slot j of the chunk in one (plane, row): its ray's value, or the slot's constant when it has no root
#define HX_SCALAR(j) (inv[j] < 0 ? k[j] : src[inv[j]])

void line_sse2(double *dst, const double *src, const int32_t *inv, const double *k, int n)
{
    int j = 0;
    for (; j < n && ((uintptr_t)(dst + j) & 15); ++j) dst[j] = HX_SCALAR(j);
    for (; j + 2 <= n; j += 2) {
        const int l0 = inv[j], l1 = inv[j + 1];
        const double x0 = src[l0 < 0 ? 0 : l0], x1 = src[l1 < 0 ? 0 : l1];     // branch-free: index 0 always exists
        _mm_stream_pd(dst + j, _mm_set_pd(l1 < 0 ? k[j + 1] : x1, l0 < 0 ? k[j] : x0));
    }
    for (; j < n; ++j) dst[j] = HX_SCALAR(j);
}

__attribute__((target("avx2"))) void line_avx2(double *dst, const double *src, const int32_t *inv, const double *k, int n)
{
    int j = 0;
    for (; j < n && ((uintptr_t)(dst + j) & 31); ++j) dst[j] = HX_SCALAR(j);
    const __m128i zero = _mm_setzero_si128();
    for (; j + 4 <= n; j += 4) {
        const __m128i l = _mm_loadu_si128((const __m128i *)(inv + j));
        const __m256d x = _mm256_i32gather_pd(src, _mm_max_epi32(l, zero), 8);
        const __m256d no_root = _mm256_castsi256_pd(_mm256_cvtepi32_epi64(l));   // sign bit set <=> no root
        _mm256_stream_pd(dst + j, _mm256_blendv_pd(x, _mm256_loadu_pd(k + j), no_root));
    }
    for (; j < n; ++j) dst[j] = HX_SCALAR(j);
}

__attribute__((target("avx512f,avx2"))) void line_avx512(double *dst, const double *src, const int32_t *inv, const double *k,
                                                         int n)
{
    int j = 0;
    for (; j < n && ((uintptr_t)(dst + j) & 63); ++j) dst[j] = HX_SCALAR(j);
    const __m256i zero = _mm256_setzero_si256();
    for (; j + 8 <= n; j += 8) {
        const __m256i l = _mm256_loadu_si256((const __m256i *)(inv + j));
        const __mmask8 no_root = (__mmask8)_mm256_movemask_ps(_mm256_castsi256_ps(l));
        const __m512d x = _mm512_i32gather_pd(_mm256_max_epi32(l, zero), src, 8);
        _mm512_stream_pd(dst + j, _mm512_mask_blend_pd(no_root, x, _mm512_loadu_pd(k + j)));
    }
    for (; j < n; ++j) dst[j] = HX_SCALAR(j);
}

typedef void (*line_fn)(double *, const double *, const int32_t *, const double *, int);

int pick_isa()
{
    int isa = 0;
    if (__builtin_cpu_supports("avx2")) isa = 1;
    if (__builtin_cpu_supports("avx512f")) isa = 2;
    if (const char *e = getenv("RWRT_HX_ISA")) {
        int want = !strcmp(e, "sse2") ? 0 : !strcmp(e, "avx2") ? 1 : !strcmp(e, "avx512") ? 2 : isa;
        if (want < isa) isa = want;
    }
    return isa;
}

}  // namespace

const char *hx_isa_name()
{
    static const char *names[] = {"sse2", "avx2", "avx512"};
    return names[pick_isa()];
}

void hx_expand_rows(const HxView &v, int b, long long s0, long long s1)
{
    static const line_fn fns[] = {line_sse2, line_avx2, line_avx512};
    const line_fn line = fns[pick_isa()];
    const double nan = __builtin_nan("");
    const int n = (int)(s1 - s0);
    if (n <= 0 || n > HX_CHUNK) return;
    // per-slot constants of this chunk: planes 0..4 of a slot without a root keep their t = 0 value (wr.py:768-781)
    static thread_local double cst[6][HX_CHUNK];
    for (int c = 0; c < 5; ++c) memcpy(cst[c], v.out + (size_t)c * v.pl + s0, (size_t)n * sizeof(double));
    for (int j = 0; j < n; ++j) cst[5][j] = nan;
    for (int r = 0; r < v.rows; ++r)
        for (int c = 0; c < 7; ++c)
            line(v.out + (size_t)c * v.pl + (size_t)(b + r) * v.N + s0, v.packed + ((size_t)c * v.rows + r) * (size_t)v.nray,
                 v.inv + s0, cst[c < 5 ? c : 5], n);
    _mm_sfence();
}

}  // namespace rwrt
