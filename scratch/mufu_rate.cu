// microbenchmark: MUFU.EX2 issue rate per SM sub-partition, alone and inside the softmax instruction mix
// (FFMA + EX2 + FADD + half an F2FP per element), for 1 / 2 / 4 warps per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
    asm volatile("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %4}; mov.b64 rc, {%5, %5};\n"
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1) {
    asm volatile("{ .reg .b64 ra, rd; mov.b64 ra, {%2, %3}; mov.b64 rd, {%0, %1};\n"
        "add.rn.f32x2 rd, rd, ra; mov.b64 {%0, %1}, rd; }" : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float scale, float negm) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = -0.01f * (threadIdx.x + i);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    uint32_t pk = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // exponentials only, 16 independent chains
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = ex2(x[i]);
        } else if (MODE == 1) {     // softmax mix
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const float p0 = ex2(fmaf(x[i], scale, negm)), p1 = ex2(fmaf(x[i + 1], scale, negm));
                if (i & 2) s0 += p0 + p1; else s1 += p0 + p1;
                __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
                pk ^= *reinterpret_cast<uint32_t*>(&h);
                x[i] -= 0.001f; x[i + 1] -= 0.002f;     // keeps the inputs changing (two more FADDs per pair)
            }
        } else if (MODE == 2) {     // FFMA only (same count as the mix: 2 per element)
#pragma unroll
            for (int i = 0; i < 16; ++i) { x[i] = fmaf(x[i], scale, negm); s2 = fmaf(x[i], scale, s2); }
        } else if (MODE == 4) {     // softmax mix with packed f32x2 FMA / ADD
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float p0, p1;
                ffma2(p0, p1, x[i], x[i + 1], scale, negm);
                p0 = ex2(p0); p1 = ex2(p1);
                if (i & 2) fadd2(s0, s1, p0, p1); else fadd2(s2, s3, p0, p1);
                __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
                pk ^= *reinterpret_cast<uint32_t*>(&h);
                fadd2(x[i], x[i + 1], -0.001f, -0.002f);
            }
        } else {                    // mix without the exponentials
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const float p0 = fmaf(x[i], scale, negm), p1 = fmaf(x[i + 1], scale, negm);
                if (i & 2) s0 += p0 + p1; else s1 += p0 + p1;
                __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
                pk ^= *reinterpret_cast<uint32_t*>(&h);
                x[i] -= 0.001f; x[i + 1] -= 0.002f;
            }
        }
    }
    const long long t1 = clock64();
    float acc = s0 + s1 + s2 + s3 + __uint_as_float(pk);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int threads) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    k<MODE><<<148, threads>>>(out, cyc, iters, 0.1275f, -0.3f);
    k<MODE><<<148, threads>>>(out, cyc, iters, 0.1275f, -0.3f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const int wps = threads / 128;   // warps per sub-partition
    printf("%-28s %d warp(s)/sub-partition: %7.2f clk per element per warp, %6.2f clk per element per sub-partition\n", name, wps,
           avg / (iters * 16.0), avg / (iters * 16.0 * wps));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int th : {128, 256, 512}) run<0>("EX2 only", th);
    for (int th : {128, 256, 512}) run<1>("FFMA+EX2+FADDx2+F2FP/2", th);
    for (int th : {128, 256, 512}) run<4>("FFMA2+EX2x2+FADD2x2+F2FP /2", th);
    for (int th : {128, 256, 512}) run<2>("FFMA x2 only", th);
    for (int th : {128, 256, 512}) run<3>("mix without EX2", th);
    return 0;
}
