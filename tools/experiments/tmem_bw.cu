// micro-benchmark: per-thread TMEM scratch traffic (tcgen05.ld/st 32x32b) at 4 blocks x 128 threads per SM
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(128, 4) k(int iters, unsigned long long *cyc, double *sink)
{
    __shared__ unsigned slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((unsigned)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned base = slot + ((unsigned)(warp * 32) << 16);
    unsigned v[8];
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(base), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    unsigned acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // ld x8 + wait
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(base));
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]));
            acc += v[0] ^ v[7];
        } else if (MODE == 1) {  // ld x2 + wait
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(base + (it & 7)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]));
            acc += v[0] ^ v[1];
        } else if (MODE == 2) {  // st x8 (no wait)
            v[0] = acc + it;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(base), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        } else if (MODE == 3) {  // st x8 + wait::st
            v[0] = acc + it;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(base), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        } else if (MODE == 4) {  // 4 x ld x8 then one wait
            unsigned w[4][8];
            for (int j = 0; j < 4; ++j)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(w[j][0]), "=r"(w[j][1]), "=r"(w[j][2]), "=r"(w[j][3]), "=r"(w[j][4]), "=r"(w[j][5]), "=r"(w[j][6]), "=r"(w[j][7]) : "r"(base + 8 * j));
            for (int j = 0; j < 4; ++j) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(w[j][0]), "+r"(w[j][1]), "+r"(w[j][2]), "+r"(w[j][3]), "+r"(w[j][4]), "+r"(w[j][5]), "+r"(w[j][6]), "+r"(w[j][7]));
                acc += w[j][0] ^ w[j][7];
            }
        } else if (MODE == 5) {  // ld 16x... x16 single
            unsigned w[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]) : "r"(base));
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(w[0]), "+r"(w[15]), "+r"(w[7]), "+r"(w[8]));
            acc += w[0] ^ w[15] ^ w[7] ^ w[8];
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    sink[blockIdx.x * 128 + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(slot) : "memory");
}
template <int MODE>
void run(const char *name, int blocks, int iters, int instr_per_iter)
{
    unsigned long long *cyc; double *sink;
    cudaMalloc(&cyc, blocks * 8); cudaMalloc(&sink, blocks * 128 * 8);
    k<MODE><<<blocks, 128>>>(iters, cyc, sink);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<blocks, 128>>>(iters, cyc, sink); cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    unsigned long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("%-28s blocks=%4d: %8.1f cycles/iter per warp, %6.2f cyc per instr per SM  (%s) %.2f ms\n", name, blocks, (double)h[0] / iters,
           (double)h[0] / iters / (instr_per_iter * (blocks >= 148 ? 4.0 * (blocks / 148) : 4.0)), cudaGetErrorString(e), ms);
}
int main()
{
    for (int blocks : {1, 148, 592}) {
        run<0>("ld.x8+wait", blocks, 20000, 1);
        run<1>("ld.x2+wait", blocks, 20000, 1);
        run<2>("st.x8", blocks, 20000, 1);
        run<3>("st.x8+wait::st", blocks, 20000, 1);
        run<4>("4x ld.x8 + waits", blocks, 20000, 4);
        run<5>("ld.x16+wait", blocks, 20000, 1);
    }
    return 0;
}
