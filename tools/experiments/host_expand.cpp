// host-side expansion probe: T threads expand synthetic time windows of packed rows (rwrt_hx.cpp) into the reference planes
//   g++ -O2 -std=c++17 -pthread [-DRWRT_HX_CHUNK=16384] tools/experiments/host_expand.cpp -o /tmp/hx && RWRT_HX_ISA=avx2 /tmp/hx 16
// (this container, 8 threads: sse2 38, avx2 45, avx512 48 GB/s written; a plain non-temporal copy: 33 / 41 / 49 GB/s)
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "../../rossby-wave-ray-tracing_b200/csrc/rwrt_hx.cpp"
using namespace rwrt;
int main(int argc, char **argv)
{
    int T = argc > 1 ? atoi(argv[1]) : (int)std::thread::hardware_concurrency();
    const long long N = 1146960;
    const int rows = 15, nt = 1 + 4 * rows;     // 4 windows' worth of output rows
    std::vector<int32_t> inv(N);
    long long nlive = 0;
    for (long long s = 0; s < N; ++s) inv[s] = ((s / 6) % 2) ? (int32_t)nlive++ : -1;   // half the slots are live, in runs
    double *stage = (double *)aligned_alloc(64, (size_t)7 * rows * nlive * 8);      // packed [7][rows][nlive]
    for (size_t i = 0; i < (size_t)7 * rows * nlive; ++i) stage[i] = (double)i;
    const size_t pl = (size_t)nt * N;
    double *out = (double *)aligned_alloc(64, 7 * pl * 8);
    memset(out, 0, 7 * pl * 8);
    HxView v{out, pl, N, inv.data(), stage, nlive, rows};
    const long long chunk = HX_CHUNK;
    const int nchunk = (int)((N + chunk - 1) / chunk);
    double best = 1e9; for (int rep = 0; rep < 12; ++rep) {
        std::atomic<int> next{0};
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&] {
                for (;;) {
                    int task = next.fetch_add(1);
                    if (task >= 4 * nchunk) return;
                    int w = task / nchunk, c = task % nchunk;
                    long long s0 = c * chunk, s1 = s0 + chunk < N ? s0 + chunk : N;
                    hx_expand_rows(v, 1 + w * rows, s0, s1);
                }
            });
        for (auto &x : th) x.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        double wr = 4.0 * rows * N * 7 * 8, rdb = 4.0 * rows * nlive * 56;
        if (s < best) best = s; if (rep == 11) printf("%d threads: best %.1f ms for 4 windows, %.1f GB/s written + %.1f GB/s read\n", T, best * 1e3, wr / best / 1e9, rdb / best / 1e9);
    }
    return 0;
}
