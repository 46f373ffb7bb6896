// b200tta_selfcheck: runs one small tcgen05 GEMM (both B layouts, ragged M) against a CUDA-core reference on
// the device.  Returns EARCH on anything that is not sm_100 -- the library has no fallback path.
#include "host_common.h"
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>
#include <vector>

namespace {
__global__ void fill_kernel(__nv_bfloat16* p, int n, unsigned seed) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        unsigned x = (i + 1) * 2654435761u ^ seed;
        x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
        p[i] = __float2bfloat16(((x & 0xffff) / 65535.0f - 0.5f) * 0.5f);
    }
}
__global__ void ref_gemm_kernel(float* D, const __nv_bfloat16* A, const __nv_bfloat16* B, int M, int N, int K, int b_mn) {
    int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (n >= N || m >= M) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k)
        acc += __bfloat162float(A[(long long)m * K + k]) * __bfloat162float(b_mn ? B[(long long)k * N + n] : B[(long long)n * K + k]);
    D[(long long)m * N + n] = acc;
}
__global__ void cmp_kernel(const float* ref, const float* got, int n, float* max_err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicMax(reinterpret_cast<int*>(max_err), __float_as_int(fabsf(ref[i] - got[i])));
}
}  // namespace

extern "C" int b200tta_selfcheck(void) {
    if (int rc = b200::require_sm100()) return rc;
    const int M = 200, N = 256, K = 192;
    __nv_bfloat16 *A = nullptr, *B = nullptr;
    float *ref = nullptr, *got = nullptr, *err = nullptr;
    B200_CUDA(cudaMalloc(&A, sizeof(__nv_bfloat16) * M * K));
    B200_CUDA(cudaMalloc(&B, sizeof(__nv_bfloat16) * N * K));
    B200_CUDA(cudaMalloc(&ref, sizeof(float) * M * N));
    B200_CUDA(cudaMalloc(&got, sizeof(float) * M * N));
    B200_CUDA(cudaMalloc(&err, sizeof(float)));
    fill_kernel<<<(M * K + 255) / 256, 256>>>(A, M * K, 1u);
    fill_kernel<<<(N * K + 255) / 256, 256>>>(B, N * K, 2u);
    int rc = B200TTA_OK;
    for (int b_mn = 0; b_mn < 2 && rc == B200TTA_OK; ++b_mn) {
        ref_gemm_kernel<<<dim3((N + 127) / 128, M), 128>>>(ref, A, B, M, N, K, b_mn);
        b200tta_gemm_seg seg;
        memset(&seg, 0, sizeof(seg));
        seg.a = A; seg.lda = K; seg.b = B; seg.ldb = b_mn ? N : K; seg.k = K; seg.b_mn_major = b_mn;
        b200tta_gemm_epi epi;
        memset(&epi, 0, sizeof(epi));
        epi.mode = B200TTA_EPI_STORE_F32; epi.d = got; epi.ldd = N;
        cudaMemset(got, 0, sizeof(float) * M * N);
        cudaMemset(err, 0, sizeof(float));
        rc = b200tta_gemm(M, N, &seg, 1, &epi, nullptr);
        if (rc != B200TTA_OK) break;
        cmp_kernel<<<(M * N + 255) / 256, 256>>>(ref, got, M * N, err);
        float h = 0.f;
        if (cudaMemcpy(&h, err, sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
            b200::set_last_error("selfcheck: device error: %s", cudaGetErrorString(cudaGetLastError()));
            rc = B200TTA_ECUDA;
        } else if (!(h < 1e-2f)) {
            b200::set_last_error("selfcheck: tcgen05 GEMM (b_mn_major=%d) differs from the reference, max |err| = %g", b_mn, h);
            rc = B200TTA_ECUDA;
        }
    }
    cudaFree(A); cudaFree(B); cudaFree(ref); cudaFree(got); cudaFree(err);
    return rc;
}
