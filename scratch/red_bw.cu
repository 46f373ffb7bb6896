// microbenchmark: chip-wide fp32 reduce-add throughput into global memory (L2 atomic units), the budget a fused
// five-product attention backward would need for its dQ partial sums: one 128x128 fp32 tile (64 KB) per (query
// tile, key block) pair, i.e. 16-25 B/clk/SM at the tensor rates the split kernels reach.
//   tma   : cp.reduce.async.bulk.global.shared::cta.add.f32, 4 x 16 KB per tile, issued by one thread
//   redv4 : red.global.add.v4.f32 from registers, 256 threads, coalesced 16 B per lane
//   red   : red.global.add.f32, coalesced 4 B per lane
// Address patterns: "staggered" = CTA c starts at tile c and walks forward, so concurrent CTAs hit distinct tiles
// (what a staggered query-tile order gives); "same" = every CTA hits the same tile at the same step (lock-step order).
// The destination is one head's dQ: 293 tiles x 64 KB = 19 MB (L2-resident), or 32 heads = 614 MB (streams through HBM).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
constexpr int TILE_F = 128 * 128;            // floats per tile
constexpr int TILE_B = TILE_F * 4;           // 64 KB
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) k_tma(float* dst, int ntiles, int iters, int same) {
    extern __shared__ __align__(128) float tile[];
    for (int i = threadIdx.x; i < TILE_F; i += blockDim.x) tile[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int it = 0; it < iters; ++it) {
            const int t = (same ? it : it + (int)blockIdx.x) % ntiles;
            float* d = dst + (size_t)t * TILE_F;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                             ::"l"(d + c * (TILE_F / 4)), "r"(smem_u32(tile + c * (TILE_F / 4))), "n"(TILE_B / 4) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // at most two tiles in flight
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
template <int VEC>
__global__ void __launch_bounds__(256) k_red(float* dst, int ntiles, int iters, int same) {
    for (int it = 0; it < iters; ++it) {
        const int t = (same ? it : it + (int)blockIdx.x) % ntiles;
        float* d = dst + (size_t)t * TILE_F;
        if (VEC == 4) {
#pragma unroll
            for (int j = 0; j < TILE_F / (256 * 4); ++j) {
                float* p = d + (j * 256 + threadIdx.x) * 4;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(1.0f) : "memory");
            }
        } else {
#pragma unroll 16
            for (int j = 0; j < TILE_F / 256; ++j) {
                float* p = d + j * 256 + threadIdx.x;
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
            }
        }
    }
}
static void run(const char* name, int kind, float* dst, int ntiles, int grid, int iters, int same, double mhz) {
    CK(cudaMemset(dst, 0, (size_t)ntiles * TILE_B));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {          // first repetition warms up
        CK(cudaEventRecord(e0));
        if (kind == 0) k_tma<<<grid, 256, TILE_B>>>(dst, ntiles, iters, same);
        else if (kind == 1) k_red<4><<<grid, 256>>>(dst, ntiles, iters, same);
        else k_red<1><<<grid, 256>>>(dst, ntiles, iters, same);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = (double)grid * iters * TILE_B;
    // correctness: every element of tile 0 received the same number of adds; total over the buffer = 2 * grid * iters * TILE_F
    float* h = (float*)malloc((size_t)ntiles * TILE_B);
    CK(cudaMemcpy(h, dst, (size_t)ntiles * TILE_B, cudaMemcpyDeviceToHost));
    double sum = 0; for (size_t i = 0; i < (size_t)ntiles * TILE_F; ++i) sum += h[i];
    free(h);
    const double expect = 2.0 * grid * iters * TILE_F;
    printf("%-6s %-9s tiles %5d grid %4d: %8.1f GB/s  %6.2f B/clk/SM (at %.0f MHz, 148 SMs)  %s\n", name, same ? "same" : "staggered", ntiles, grid,
           bytes / ms * 1e-6, bytes / (ms * 1e-3) / (mhz * 1e6) / 148.0, mhz, sum == expect ? "sum ok" : "SUM MISMATCH");
}
int main() {
    int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, max clock %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_B));
    const int big = 293 * 32, small = 293;
    float* dst; CK(cudaMalloc(&dst, (size_t)big * TILE_B));
    const int iters = 400;
    for (int kind = 0; kind < 3; ++kind) {
        const char* nm = kind == 0 ? "tma" : kind == 1 ? "redv4" : "red";
        const int it = kind == 2 ? iters / 4 : iters;
        run(nm, kind, dst, small, 148, it, 0, mhz);
        run(nm, kind, dst, small, 296, it, 0, mhz);
        run(nm, kind, dst, big, 296, it, 0, mhz);
        run(nm, kind, dst, small, 148, it / 4, 1, mhz);
    }
    CK(cudaFree(dst));
    return 0;
}
