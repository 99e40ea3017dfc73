// FP64 tensor-core (mma.sync m8n8k4, SASS DMMA.8x8x4) throughput on one B200: register-resident chains, 8 independent
// accumulators per warp.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 dmma_peak.cu -o dmma_peak && ./dmma_peak
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(double *sink, int iters, double a, double b)
{
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}
int main()
{
    double *sink;
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int warps = 1; warps <= 8; warps *= 2) {
        const int iters = 20000, blocks = sms * 4, threads = 32 * warps;
        k<<<blocks, threads>>>(sink, 100, 1e-3, 1e-3);
        cudaEventRecord(e0);
        k<<<blocks, threads>>>(sink, iters, 1e-3, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double flop = 2.0 * 256 * 8.0 * iters * blocks * warps;
        printf("%d warps per block x 4 blocks per SM: %.2f TFLOP/s fp64 via DMMA.8x8x4\n", warps, flop / ms / 1e9);
    }
    return 0;
}
