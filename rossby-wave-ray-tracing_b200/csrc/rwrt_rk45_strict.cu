// Strict build of the RK45 kernels: compiled with -fmad=false (see build.py).
#define RWRT_VARIANT strict
#define RWRT_FAST false
#include "rwrt_rk45.cuh"
#include "rwrt_internal.h"

namespace rwrt {
void launch_rk45_setup_strict(const RunParams &p, int prepare_blocks, int fill_blocks, cudaStream_t s, cudaStream_t aux)
{
    if (p.nlive > 0) strict::rk45_prepare<<<prepare_blocks, 128, 0, s>>>(p);
    // The constant rows of NaN-root slots touch no slot the integrator writes: low-priority side stream.
    if (p.N > 0 && p.nt > 1 && !p.rec && !p.no_rows)   // record layout: expand_records writes every row instead; sweep: no rows at all
        strict::fill_static_slots<<<dim3(fill_blocks, p.nt - 1 < 24 ? p.nt - 1 : 24), 256, 0, aux>>>(p);
}
void launch_expand_records_strict(const RunParams &p, cudaStream_t s)
{
    int rows = p.it_end - p.it_begin;
    if (p.N > 0 && rows > 0)
        strict::expand_records<<<dim3((unsigned)((p.N + strict::EX_SLOTS - 1) / strict::EX_SLOTS), (rows + strict::EX_ROWS_PER_BLOCK - 1) / strict::EX_ROWS_PER_BLOCK), 256, 0, s>>>(p);
}

// One time window [p.it_begin, p.it_end) of output rows for every live ray (the per-lane kernel has no
// windows: it is launched once with the whole range).
void launch_rk45_window_strict(const RunParams &p, int integrate_blocks, cudaStream_t s)
{
    if (p.nlive <= 0) return;
    if (p.coop) {
        const bool grid = p.g_count != nullptr;
        if (p.no_rows && grid)
            strict::rk45_integrate_coop<OUT_GRID><<<integrate_blocks, strict::CO_THREADS, strict::CO_SMEM_BYTES, s>>>(p);
        else if (grid)
            strict::rk45_integrate_coop<OUT_ROWS | OUT_GRID><<<integrate_blocks, strict::CO_THREADS, strict::CO_SMEM_BYTES, s>>>(p);
        else
            strict::rk45_integrate_coop<OUT_ROWS><<<integrate_blocks, strict::CO_THREADS, strict::CO_SMEM_BYTES, s>>>(p);
    } else
        strict::rk45_integrate<<<integrate_blocks, strict::RK_THREADS, strict::RK_SMEM_BYTES, s>>>(p);
}
void launch_rk4_strict(const RunParams &p, cudaStream_t s)
{
    if (p.N <= 0 || p.nt <= 1) return;
    if (p.coop) {
        cudaFuncSetAttribute(strict::rk4_integrate_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, strict::RK4_SMEM_BYTES);
        strict::rk4_integrate_coop<<<(unsigned)((p.N + 127) / 128), strict::RK_THREADS, strict::RK4_SMEM_BYTES, s>>>(p);
    } else {
        strict::rk4_integrate<<<(unsigned)((p.N + 127) / 128), 128, 0, s>>>(p);
    }
}

int threads_rk45_strict(bool coop) { return coop ? strict::CO_THREADS : strict::RK_THREADS; }
int occupancy_rk45_strict(bool coop)
{
    int nb = 0;
    if (coop) {
        cudaFuncSetAttribute(strict::rk45_integrate_coop<OUT_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, strict::CO_SMEM_BYTES);
        cudaFuncSetAttribute(strict::rk45_integrate_coop<OUT_GRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, strict::CO_SMEM_BYTES);
        cudaFuncSetAttribute(strict::rk45_integrate_coop<OUT_ROWS | OUT_GRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, strict::CO_SMEM_BYTES);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, strict::rk45_integrate_coop<OUT_ROWS>, strict::CO_THREADS, strict::CO_SMEM_BYTES);
#if RWRT_TMEM_KS
        // the occupancy calculator answers 1 for a kernel that allocates tensor memory (it cannot know how many
        // columns tcgen05.alloc will ask for); this build takes 64 of the 512 columns per block
        if (nb < strict::CO_BLOCKS_PER_SM) nb = strict::CO_BLOCKS_PER_SM;
#endif
        return nb;
    }
    cudaFuncSetAttribute(strict::rk45_integrate, cudaFuncAttributeMaxDynamicSharedMemorySize, strict::RK_SMEM_BYTES);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, strict::rk45_integrate, strict::RK_THREADS, strict::RK_SMEM_BYTES);
    return nb;
}
__global__ void rhs_kernel_strict(Grid g, long long n, const double *y, const int32_t *member, double *dy)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double yy[5], d[5];
    for (int c = 0; c < 5; ++c) yy[c] = y[c * n + i];
    const double *fb = g.fpk + (member ? (size_t)member[i] * g.member_stride : 0);
    ray_rhs<RWRT_FAST>(g, fb, yy, d);
    for (int c = 0; c < 5; ++c) dy[c * n + i] = d[c];
}
void launch_rhs_strict(const Grid &g, long long n, const double *y, const int32_t *member, double *dy, cudaStream_t s)
{
    if (n > 0) rhs_kernel_strict<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g, n, y, member, dy);
}
}  // namespace rwrt
