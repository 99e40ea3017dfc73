// Fast build of the RK45 kernels: default nvcc floating-point contraction (FMA) on.
#define RWRT_VARIANT fast
#define RWRT_FAST true
#include "rwrt_rk45.cuh"
#include "rwrt_internal.h"

namespace rwrt {
void launch_rk45_setup_fast(const RunParams &p, int prepare_blocks, int fill_blocks, cudaStream_t s, cudaStream_t aux)
{
    if (p.nlive > 0) fast::rk45_prepare<<<prepare_blocks, 128, 0, s>>>(p);
    // The constant rows of NaN-root slots touch no slot the integrator writes: low-priority side stream.
    if (p.N > 0 && p.nt > 1 && !p.rec && !p.no_rows)   // record layout: expand_records writes every row instead; sweep: no rows at all
        fast::fill_static_slots<<<dim3(fill_blocks, p.nt - 1 < 24 ? p.nt - 1 : 24), 256, 0, aux>>>(p);
}
void launch_expand_records_fast(const RunParams &p, cudaStream_t s)
{
    int rows = p.it_end - p.it_begin;
    if (p.N > 0 && rows > 0)
        fast::expand_records<<<dim3((unsigned)((p.N + fast::EX_SLOTS - 1) / fast::EX_SLOTS), (rows + fast::EX_ROWS_PER_BLOCK - 1) / fast::EX_ROWS_PER_BLOCK), 256, 0, s>>>(p);
}

// One time window [p.it_begin, p.it_end) of output rows for every live ray (the per-lane kernel has no
// windows: it is launched once with the whole range).
void launch_rk45_window_fast(const RunParams &p, int integrate_blocks, cudaStream_t s)
{
    if (p.nlive <= 0) return;
    if (p.coop) {
        const bool grid = p.g_count != nullptr;
        if (p.no_rows && grid)
            fast::rk45_integrate_coop<OUT_GRID><<<integrate_blocks, fast::CO_THREADS, fast::CO_SMEM_BYTES, s>>>(p);
        else if (grid)
            fast::rk45_integrate_coop<OUT_ROWS | OUT_GRID><<<integrate_blocks, fast::CO_THREADS, fast::CO_SMEM_BYTES, s>>>(p);
        else
            fast::rk45_integrate_coop<OUT_ROWS><<<integrate_blocks, fast::CO_THREADS, fast::CO_SMEM_BYTES, s>>>(p);
    } else
        fast::rk45_integrate<<<integrate_blocks, fast::RK_THREADS, fast::RK_SMEM_BYTES, s>>>(p);
}
void launch_rk4_fast(const RunParams &p, cudaStream_t s)
{
    if (p.N <= 0 || p.nt <= 1) return;
    if (p.coop) {
        cudaFuncSetAttribute(fast::rk4_integrate_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, fast::RK4_SMEM_BYTES);
        fast::rk4_integrate_coop<<<(unsigned)((p.N + 127) / 128), fast::RK_THREADS, fast::RK4_SMEM_BYTES, s>>>(p);
    } else {
        fast::rk4_integrate<<<(unsigned)((p.N + 127) / 128), 128, 0, s>>>(p);
    }
}

int threads_rk45_fast(bool coop) { return coop ? fast::CO_THREADS : fast::RK_THREADS; }
int occupancy_rk45_fast(bool coop)
{
    int nb = 0;
    if (coop) {
        cudaFuncSetAttribute(fast::rk45_integrate_coop<OUT_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast::CO_SMEM_BYTES);
        cudaFuncSetAttribute(fast::rk45_integrate_coop<OUT_GRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast::CO_SMEM_BYTES);
        cudaFuncSetAttribute(fast::rk45_integrate_coop<OUT_ROWS | OUT_GRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast::CO_SMEM_BYTES);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fast::rk45_integrate_coop<OUT_ROWS>, fast::CO_THREADS, fast::CO_SMEM_BYTES);
#if RWRT_TMEM_KS
        // the occupancy calculator answers 1 for a kernel that allocates tensor memory (it cannot know how many
        // columns tcgen05.alloc will ask for); this build takes 64 of the 512 columns per block
        if (nb < fast::CO_BLOCKS_PER_SM) nb = fast::CO_BLOCKS_PER_SM;
#endif
        return nb;
    }
    cudaFuncSetAttribute(fast::rk45_integrate, cudaFuncAttributeMaxDynamicSharedMemorySize, fast::RK_SMEM_BYTES);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fast::rk45_integrate, fast::RK_THREADS, fast::RK_SMEM_BYTES);
    return nb;
}
__global__ void rhs_kernel_fast(Grid g, long long n, const double *y, const int32_t *member, double *dy)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double yy[5], d[5];
    for (int c = 0; c < 5; ++c) yy[c] = y[c * n + i];
    const double *fb = g.fpk + (member ? (size_t)member[i] * g.member_stride : 0);
    ray_rhs<RWRT_FAST>(g, fb, yy, d);
    for (int c = 0; c < 5; ++c) dy[c * n + i] = d[c];
}
void launch_rhs_fast(const Grid &g, long long n, const double *y, const int32_t *member, double *dy, cudaStream_t s)
{
    if (n > 0) rhs_kernel_fast<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g, n, y, member, dy);
}
}  // namespace rwrt
