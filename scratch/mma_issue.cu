// microbenchmark: issue style for the dq kernel's per-sub-block MMA mix.
//   STYLE 0: every tcgen05.mma individually predicated on elect.sync (umma_*_e)
//   STYLE 1: one elect.sync per batch, the batch issued inside the elected branch
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../longcat_video_tta_b200/csrc/ptx.cuh"
using namespace b200;
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
template <int STYLE, int NS, int ND, int WAITS, int WK = 1, int COMMITS = 1, int DUAL = 0>
__global__ void __launch_bounds__(576, 1) k(long long* cycles, int reps) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t tptr;
    __shared__ volatile uint32_t flag;
    for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { flag = 7; mbar_init(&bar, DUAL ? 2 : 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tptr);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const int warp = threadIdx.x >> 5;
    if (warp == 1 || (DUAL && warp == 2)) {
        const uint32_t b = smem_u32(smem);
        constexpr uint32_t i64 = umma_idesc_bf16(128, NS, 0, 0), i128 = umma_idesc_bf16(128, ND, 0, 1);
        long long t0 = clock64();
        for (int r = DUAL ? warp - 1 : 0; r < reps; r += DUAL ? 2 : 1) {
            const uint32_t buf = (r & 1) * 128;
            if (WAITS) { for (int w = 0; w < WAITS; ++w) {
                if (WK == 1) { mbar_wait(&bar3, 1); tc_fence_after(); }
                if (WK == 2) { mbar_wait(&bar3, 1); }
                if (WK == 3) { while (!test_wait(&bar3, 1)) {} }
                if (WK == 4) { tc_fence_after(); }
                if (WK == 6) { while (flag != 7) {} }
                if (WK == 5) { while (!test_wait(&bar3, 1)) {} tc_fence_after(); }
            } }  // parity 1 of a fresh barrier: already complete
            const int stg = r % 6;
            const uint32_t kb = b + stg * 16384, vb = b + 98304 + stg * 16384;
            const uint64_t kd = umma_desc_kmajor(kb), vd = umma_desc_kmajor(vb), kmn = umma_desc_mnmajor(kb, 8192);
            if (STYLE == 0) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) umma_ts_e(buf, 384 + ks * 8, umma_desc_advance(kd, (ks >> 2) * 8192 + (ks & 3) * 32), i64, ks ? 1 : 0);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) umma_ts_e(buf + 64, 448 + ks * 8, umma_desc_advance(vd, (ks >> 2) * 8192 + (ks & 3) * 32), i64, ks ? 1 : 0);
                umma_commit_e(&bar2);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts_e(256, buf + 64 + (ks >> 1) * 32 + (ks & 1) * 8, umma_desc_advance(kmn, ks * 2048), i128, 1);
                umma_commit_e(&bar2);
            } else {
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) umma_ts(buf, 384 + ks * 8, umma_desc_advance(kd, (ks >> 2) * 8192 + (ks & 3) * 32), i64, ks ? 1 : 0);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) umma_ts(buf + 64, 448 + ks * 8, umma_desc_advance(vd, (ks >> 2) * 8192 + (ks & 3) * 32), i64, ks ? 1 : 0);
                    if (COMMITS) umma_commit(&bar2);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_ts(256, buf + 64 + (ks >> 1) * 32 + (ks & 1) * 8, umma_desc_advance(kmn, ks * 2048), i128, 1);
                    if (COMMITS) umma_commit(&bar2);
                }
                __syncwarp();
            }
        }
        umma_commit_e(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 2 + warp - 1] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tptr); }
}
int main(int argc, char** argv) {
    int style = argc > 1 ? atoi(argv[1]) : 0;
    long long* cyc; cudaMalloc(&cyc, 148 * 16);
    const int reps = 2000;
#define RUN(S, NS, ND, W, ...) { cudaFuncSetAttribute(k<S, NS, ND, W, ##__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k<S, NS, ND, W, ##__VA_ARGS__><<<148, 576, 200 * 1024>>>(cyc, reps); }
    switch (style) {
        case 0: RUN(0, 64, 128, 0); break;
        case 1: RUN(1, 64, 128, 0); break;
        case 2: RUN(0, 16, 16, 0); break;   // tiny MMAs: the tensor pipe needs ~8 clk each, the rest is issue cost
        case 3: RUN(1, 16, 16, 0); break;
        case 4: RUN(0, 16, 16, 3); break;   // + three waits on an already-complete mbarrier per sub-block
        case 5: RUN(1, 16, 16, 3); break;
        case 6: RUN(0, 64, 128, 3); break;
        case 7: RUN(1, 64, 128, 3); break;
        case 12: RUN(1, 16, 16, 3, 2); break;
        case 13: RUN(1, 16, 16, 3, 3); break;
        case 14: RUN(1, 16, 16, 3, 4); break;
        case 15: RUN(1, 16, 16, 3, 5); break;
        case 25: RUN(1, 64, 128, 3, 5); break;
        case 30: RUN(1, 64, 128, 3, 5, 1, 1); break;
        case 31: RUN(1, 64, 128, 3, 1, 1, 1); break;
        case 32: RUN(0, 64, 128, 3, 1, 1, 1); break;
        case 26: RUN(1, 64, 128, 3, 5, 0); break;
        case 27: RUN(1, 64, 128, 3, 6, 1); break;
        case 28: RUN(1, 16, 16, 3, 6, 1); break;
        case 29: RUN(1, 16, 16, 3, 5, 0); break;
    }
    cudaError_t e = cudaDeviceSynchronize();
    long long hh[2] = {0, 0}; cudaMemcpy(hh, cyc, 16, cudaMemcpyDeviceToHost); long long h = hh[0] > hh[1] ? hh[0] : hh[1];
    printf("style %d: %7.1f cycles per sub-block (tensor-pipe time 768)  [%s]\n", style, (double)h / reps, cudaGetErrorString(e));
    return 0;
}
