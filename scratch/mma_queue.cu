// microbenchmark: how deep is the tcgen05.mma issue queue?  Issue 12 MMAs (N=128: 64 clk, or N=64: 32 clk of pipe time
// each) from an idle pipe and record when each issue instruction retires from the issuing thread's point of view.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../longcat_video_tta_b200/csrc/ptx.cuh"
using namespace b200;
template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tptr);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (threadIdx.x < 32) {
        const uint32_t b = smem_u32(smem);
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
        const uint64_t bd = umma_desc_kmajor(b);
        long long t[14];
        for (int rep = 0; rep < 2; ++rep) {   // second repetition = warm instruction cache
            t[0] = clock64();
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                umma_ts_e(256, (i & 7) * 8, umma_desc_advance(bd, (i & 3) * 32), idesc, 1);
                t[i + 1] = clock64();
            }
            umma_commit_e(&bar);
            mbar_wait(&bar, rep);
            t[13] = clock64();
        }
        if (threadIdx.x == 0 && blockIdx.x == 0) for (int i = 0; i < 14; ++i) out[i] = t[i] - t[0];
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tptr); }
}
int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 128;
    long long* d; cudaMalloc(&d, 14 * 8);
    if (n == 128) { cudaFuncSetAttribute(k<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024); k<128><<<1, 128, 140 * 1024>>>(d); }
    else { cudaFuncSetAttribute(k<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024); k<64><<<1, 128, 140 * 1024>>>(d); }
    cudaError_t e = cudaDeviceSynchronize();
    long long h[14]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%d issue-return times (clk since first issue):", n);
    for (int i = 1; i <= 12; ++i) printf(" %lld", h[i]);
    printf(" | all complete %lld [%s]\n", h[13], cudaGetErrorString(e));
    return 0;
}
