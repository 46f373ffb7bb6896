// microbenchmark: cycles per tcgen05.mma (M=128, bf16, K=16) for the operand forms used by the attention kernels
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../longcat_video_tta_b200/csrc/ptx.cuh"
using namespace b200;
// mode 0: SS, A K-major, B K-major ; 1: SS, B MN-major ; 2: TS (A in TMEM), B K-major ; 3: TS, B MN-major
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) k(long long* cycles, int reps) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tptr);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t t = tptr;
    if (threadIdx.x < 32) {
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 65536);
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, (MODE & 1));
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {   // one 128 x N x 128 product = 8 MMAs of K=16
                const uint64_t bd = (MODE & 1) ? umma_desc_mnmajor(b + (ks & 7) * 2048, 16384) : umma_desc_kmajor(b + (ks >> 2) * 32768 + (ks & 3) * 32);
                if (MODE & 2) umma_ts_e(256, ks * 8, bd, idesc, 1);
                else umma_ss_e(256, umma_desc_kmajor(a + (ks >> 2) * 16384 + (ks & 3) * 32), bd, idesc, 1);
            }
        }
        umma_commit_e(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(t); }
}
template <int MODE, int N> void run(const char* name) {
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(k<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 2000;
    k<MODE, N><<<148, 128, 200 * 1024>>>(cyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s N=%3d : %7.1f cycles per K=16 MMA, %7.1f per 128xNx128 product  (ideal %d)  [%s]\n", name, N,
           (double)h / (reps * 8), (double)h / reps, N / 2 * 8 / 1 * 1, cudaGetErrorString(e));
    cudaFree(cyc);
}
int main(int argc, char** argv) {
    int w = argc > 1 ? atoi(argv[1]) : 0;
    switch (w) {
        case 0: run<0, 128>("SS  A K-major, B K-major"); break;
        case 1: run<0, 64>("SS  A K-major, B K-major"); break;
        case 2: run<0, 256>("SS  A K-major, B K-major"); break;
        case 3: run<1, 128>("SS  A K-major, B MN-major"); break;
        case 4: run<1, 256>("SS  A K-major, B MN-major"); break;
        case 5: run<2, 128>("TS  A TMEM,    B K-major"); break;
        case 6: run<3, 128>("TS  A TMEM,    B MN-major"); break;
        case 7: run<3, 64>("TS  A TMEM,    B MN-major"); break;
        case 8: run<2, 64>("TS  A TMEM,    B K-major"); break;
        case 9: run<0, 32>("SS  A K-major, B K-major"); break;
    }
    return 0;
}
