// micro-benchmark: how the lanes of a quarter-warp are mapped onto the 24 16-byte pieces of a bilinear cell
// (2 runs of 192 contiguous bytes: points (x0,y0),(x0,y1) and (x1,y0),(x1,y1) of the packed point table) when
// they are copied into the warp's staging rows with cp.async -- the gather of rk45_integrate_coop, isolated.
//   MAP 0 (the kernel up to v20): lane q8 -> corner q8>>1, half q8&1 of sector i = 0..2 (4 x 32 B per instruction)
//   MAP 1: instruction j copies pieces 8j .. 8j+7 of the 24 (128 contiguous bytes per quarter-warp and instruction,
//          except j = 1: 64 B of each run)
//   MAP 2: lanes 0-3 / 4-7 copy 64 contiguous bytes of run 0 / run 1 per instruction
// and what the owner lane's reads cost: LD 0 = 20 LDS.128 + 4 LDS.64 (the four 11th fields), LD 1 = 22 LDS.128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldgsts_map ldgsts_map.cu && ./ldgsts_map
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int REC = 400, NLAT = 181, NLON = 361, PT = 96;

__device__ __forceinline__ void cp16(unsigned dst, const void *src, unsigned long long pol, bool cg)
{
    if (cg) asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "l"(pol));
    else asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "l"(pol));
}

template <int MAP, int LD, bool CG>
__global__ void __launch_bounds__(128, 3) k(const char *tab, int iters, int nfetch, double *sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *stage = smem + warp * 32 * REC;
    const unsigned stage_u32 = (unsigned)__cvta_generic_to_shared(stage);
    const int q8 = lane & 7, grp = lane >> 3;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
    const unsigned dx = NLAT * PT;
    unsigned so[3], d_o[3];
    for (int j = 0; j < 3; ++j) {
        if (MAP == 0) {
            so[j] = (q8 >> 1 & 1) * PT + (q8 >> 2) * dx + (q8 & 1) * 16 + 32 * j;
            d_o[j] = (q8 >> 1) * 96 + (q8 & 1) * 16 + 32 * j;
        } else if (MAP == 1) {
            const int p = 8 * j + q8, r = p / 12, o = (p % 12) * 16;
            so[j] = r * dx + o;
            d_o[j] = r * 192 + o;
        } else {
            const int r = q8 >> 2, o = (q8 & 3) * 16 + 64 * j;
            so[j] = r * dx + o;
            d_o[j] = r * 192 + o;
        }
    }
    unsigned rng = (blockIdx.x * 4 + warp) * 2654435761u + 12345u;
    double acc = 0;
    const unsigned char *mine = stage + lane * REC;
    for (int it = 0; it < iters; ++it) {
        __syncwarp();
        for (int e = grp; e < nfetch; e += 4) {
            unsigned h = (rng + e * 0x9e3779b9u) * 747796405u;
            h ^= h >> 15;
            const unsigned x = h % (NLON - 1), y = (h >> 12) % (NLAT - 1);
            const char *sp = tab + ((size_t)x * NLAT + y) * PT;
            const unsigned dst = stage_u32 + ((e * 5) & 31) * REC;
            cp16(dst + d_o[0], sp + so[0], pol, CG);
            cp16(dst + d_o[1], sp + so[1], pol, CG);
            cp16(dst + d_o[2], sp + so[2], pol, CG);
        }
        rng = rng * 1664525u + 1013904223u;
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncwarp();
        if (LD == 0) {
#pragma unroll
            for (int kx = 0; kx < 6; ++kx) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (kx < 5) {
                        const double2 v = *reinterpret_cast<const double2 *>(mine + c * 96 + kx * 16);
                        acc += v.x * v.y;
                    } else {
                        acc += *reinterpret_cast<const double *>(mine + c * 96 + kx * 16);
                    }
                }
            }
        } else {
#pragma unroll
            for (int kx = 0; kx < 22; ++kx) {
                const double2 v = *reinterpret_cast<const double2 *>(mine + kx * 16);
                acc += v.x * v.y;
            }
        }
    }
    sink[blockIdx.x * 128 + threadIdx.x] = acc;
}

template <int MAP, int LD, bool CG>
static void run(const char *name, const char *tab, double *sink, int nfetch)
{
    const int smem = 4 * 32 * REC, iters = 4000, grid = 148 * 3;
    cudaFuncSetAttribute(k<MAP, LD, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<MAP, LD, CG><<<grid, 128, smem>>>(tab, 200, nfetch, sink);
    cudaEventRecord(a);
    k<MAP, LD, CG><<<grid, 128, smem>>>(tab, iters, nfetch, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    // 12 warps per SM, each `iters` iterations
    const double cyc = ms * 1e-3 * 1.965e9 / (12.0 * iters);
    printf("%-34s nfetch %2d: %8.3f ms  %7.1f SM-cycles per warp-iteration  (%s)\n", name, nfetch, ms, cyc, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    char *tab;
    double *sink;
    const size_t bytes = (size_t)NLON * NLAT * PT + 4096;
    cudaMalloc(&tab, bytes);
    cudaMemset(tab, 0, bytes);
    cudaMalloc(&sink, 148 * 3 * 128 * sizeof(double));
    for (int nf : {11, 32}) {
        run<0, 0, false>("map0 corner/half   ld 20x128+4x64", tab, sink, nf);
        run<1, 0, false>("map1 contiguous    ld 20x128+4x64", tab, sink, nf);
        run<2, 0, false>("map2 2 x 64 B      ld 20x128+4x64", tab, sink, nf);
        run<0, 1, false>("map0 corner/half   ld 22x128", tab, sink, nf);
        run<1, 1, false>("map1 contiguous    ld 22x128", tab, sink, nf);
        run<0, 0, true>("map0 corner/half cg ld 20x128+4x64", tab, sink, nf);
        run<1, 0, true>("map1 contiguous  cg ld 20x128+4x64", tab, sink, nf);
    }
    run<0, 0, false>("map0, no copies    ld 20x128+4x64", tab, sink, 0);
    run<0, 1, false>("map0, no copies    ld 22x128", tab, sink, 0);
    return 0;
}
