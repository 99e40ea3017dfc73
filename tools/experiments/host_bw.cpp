// host memory bandwidth probe: T threads, each streams a private 256 MB source into a private 256 MB destination
// (a) memcpy, (b) 64-byte non-temporal stores of gathered 8-byte values (the access pattern of a host-side transpose)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>
int main(int argc, char **argv)
{
    int T = argc > 1 ? atoi(argv[1]) : (int)std::thread::hardware_concurrency();
    const size_t n = 32u << 20;  // doubles per thread (256 MB)
    std::vector<double *> src(T), dst(T);
    for (int t = 0; t < T; ++t) {
        src[t] = (double *)aligned_alloc(64, n * 8); dst[t] = (double *)aligned_alloc(64, n * 8);
        memset(src[t], 1, n * 8); memset(dst[t], 0, n * 8);
    }
    for (int mode = 0; mode < 2; ++mode) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                if (mode == 0) memcpy(dst[t], src[t], n * 8);
                else
                    for (size_t i = 0; i < n; i += 8) {
                        __m128d a = _mm_load_pd(src[t] + i), b = _mm_load_pd(src[t] + i + 2), c = _mm_load_pd(src[t] + i + 4), d = _mm_load_pd(src[t] + i + 6);
                        _mm_stream_pd(dst[t] + i, a); _mm_stream_pd(dst[t] + i + 2, b); _mm_stream_pd(dst[t] + i + 4, c); _mm_stream_pd(dst[t] + i + 6, d);
                    }
            });
        for (auto &x : th) x.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("%s: %d threads, %.1f GB/s read + %.1f GB/s write\n", mode ? "nt-store copy" : "memcpy", T, T * n * 8 / s / 1e9, T * n * 8 / s / 1e9);
    }
    return 0;
}
