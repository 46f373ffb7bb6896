// microbenchmark: TS-form tcgen05.mma (A in TMEM) rate as a function of where A and D sit in TMEM and of the accumulate pattern
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../longcat_video_tta_b200/csrc/ptx.cuh"
using namespace b200;
template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* cycles, int reps, uint32_t a_col, uint32_t d0, uint32_t d1, int acc_first, int a_step) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tptr);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (threadIdx.x < 32) {
        const uint32_t b = smem_u32(smem);
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
        const uint64_t bd = umma_desc_kmajor(b);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = (r & 1) ? d1 : d0;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
                umma_ts_e(d, a_col + ks * a_step, umma_desc_advance(bd, (ks >> 2) * 8192 + (ks & 3) * 32), idesc, (ks || acc_first) ? 1 : 0);
        }
        umma_commit_e(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tptr); }
}
int main(int argc, char** argv) {
    const int n = atoi(argv[1]); const uint32_t a_col = atoi(argv[2]), d0 = atoi(argv[3]), d1 = atoi(argv[4]); const int acc_first = atoi(argv[5]);
    const int a_step = argc > 6 ? atoi(argv[6]) : 8;
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    const int reps = 2000;
    if (n == 64) { cudaFuncSetAttribute(k<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k<64><<<148, 128, 200 * 1024>>>(cyc, reps, a_col, d0, d1, acc_first, a_step); }
    else { cudaFuncSetAttribute(k<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k<128><<<148, 128, 200 * 1024>>>(cyc, reps, a_col, d0, d1, acc_first, a_step); }
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("TS N=%d A@%u(step %d) D@%u/%u acc_first=%d : %6.1f clk per MMA [%s]\n", n, a_col, a_step, d0, d1, acc_first, (double)h / (reps * 8), cudaGetErrorString(e));
    return 0;
}
