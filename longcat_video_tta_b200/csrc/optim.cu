// Multi-tensor gradient-norm, clip coefficient and AdamW over the adapter parameters: one launch each for the
// whole tensor list (480 LoRA tensors; 1..49 for the delta / norm / FiLM methods).  HBM-bound: 128-bit accesses
// where the layout allows, grid sized over (chunks, tensors).
#include "host_common.h"
#include "ptx.cuh"
#include <math.h>

namespace b200 {
namespace {

constexpr int OPT_THREADS = 256;
constexpr int OPT_CHUNK = OPT_THREADS * 8;  // elements per block iteration

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

// gradient element for parameter element i (grads of LoRA "down" matrices are accumulated transposed)
__device__ __forceinline__ float grad_at(const b200tta_tensor_desc& d, long long i) {
    if (d.t_rows > 0) {
        const long long row = i / d.t_cols, col = i - row * d.t_cols;  // param [t_rows, t_cols]; grad [t_cols, t_rows]
        return d.grad[col * d.t_rows + row];
    }
    return d.grad[i];
}

__global__ void __launch_bounds__(OPT_THREADS) sumsq_kernel(const b200tta_tensor_desc* __restrict__ descs,
                                                            float* __restrict__ sumsq) {
    __shared__ float red[OPT_THREADS / 32];
    const b200tta_tensor_desc d = descs[blockIdx.y];
    float s = 0.f;
    if ((d.numel & 3) == 0 && (reinterpret_cast<uintptr_t>(d.grad) & 15) == 0) {
        const float4* G4 = reinterpret_cast<const float4*>(d.grad);
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < (d.numel >> 2); i += (long long)gridDim.x * OPT_THREADS) {
            const float4 g = G4[i];
            s4[0] = fmaf(g.x, g.x, s4[0]); s4[1] = fmaf(g.y, g.y, s4[1]); s4[2] = fmaf(g.z, g.z, s4[2]); s4[3] = fmaf(g.w, g.w, s4[3]);
        }
        s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    } else {
        for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < d.numel; i += (long long)gridDim.x * OPT_THREADS) {
            const float g = d.grad[i];  // layout does not matter for a sum of squares
            s += g * g;
        }
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < OPT_THREADS / 32) {
        s = red[threadIdx.x];
        for (int o = OPT_THREADS / 64; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
        if (threadIdx.x == 0 && s != 0.f) atomicAdd(sumsq + blockIdx.y, s);
    }
}

__global__ void clip_coef_kernel(float* __restrict__ coef, float* __restrict__ total_out, const float* __restrict__ sumsq,
                                 int n, float max_norm, int per_tensor, float grad_scale) {
    // single block; grad_scale multiplies every gradient before the norm (e.g. 1/world after an all-reduce sum)
    __shared__ float red[32];
    __shared__ float total_s;
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += sumsq[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x + 31) / 32 ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) total_s = sqrtf(s) * grad_scale;
    }
    __syncthreads();
    const float total = total_s;
    if (threadIdx.x == 0 && total_out) *total_out = total;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float norm = per_tensor ? sqrtf(sumsq[i]) * grad_scale : total;
        coef[i] = fminf(1.0f, max_norm / (norm + 1e-6f));
    }
}

struct AdamArgs {
    float grad_scale, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt;
    int faithful_bf16;
};

__global__ void __launch_bounds__(OPT_THREADS) adamw_kernel(const b200tta_tensor_desc* __restrict__ descs,
                                                            const float* __restrict__ coef, AdamArgs a) {
    const b200tta_tensor_desc d = descs[blockIdx.y];
    const float c = (coef ? coef[blockIdx.y] : 1.0f);
    const float step_size = a.lr / a.bc1;
    for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < d.numel; i += (long long)gridDim.x * OPT_THREADS) {
        float g = grad_at(d, i) * a.grad_scale;
        if (d.is_bf16 && d.master == nullptr && a.faithful_bf16) {
            // torch foreach AdamW on bf16 tensors: every op computes in fp32 and rounds its result to bf16
            __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(d.param);
            __nv_bfloat16* M = reinterpret_cast<__nv_bfloat16*>(d.exp_avg);
            __nv_bfloat16* V = reinterpret_cast<__nv_bfloat16*>(d.exp_avg_sq);
            g = bf16r(bf16r(g) * c);
            float p = __bfloat162float(P[i]), m = __bfloat162float(M[i]), v = __bfloat162float(V[i]);
            p = bf16r(p * (1.0f - a.lr * a.wd));
            m = bf16r(m + (1.0f - a.beta1) * (g - m));
            v = bf16r(v * a.beta2);
            v = bf16r(v + (1.0f - a.beta2) * g * g);
            float den = bf16r(sqrtf(v));
            den = bf16r(den / a.bc2_sqrt);
            den = bf16r(den + a.eps);
            p = bf16r(p + (-step_size) * (m / den));
            P[i] = __float2bfloat16(p); M[i] = __float2bfloat16(m); V[i] = __float2bfloat16(v);
        } else {
            // fp32 math; states fp32 when a master copy exists or the parameter itself is fp32
            g *= c;
            const bool f32_state = d.master != nullptr || !d.is_bf16;
            float p = d.master ? d.master[i]
                               : (d.is_bf16 ? __bfloat162float(reinterpret_cast<__nv_bfloat16*>(d.param)[i])
                                            : reinterpret_cast<float*>(d.param)[i]);
            float m = f32_state ? reinterpret_cast<float*>(d.exp_avg)[i]
                                : __bfloat162float(reinterpret_cast<__nv_bfloat16*>(d.exp_avg)[i]);
            float v = f32_state ? reinterpret_cast<float*>(d.exp_avg_sq)[i]
                                : __bfloat162float(reinterpret_cast<__nv_bfloat16*>(d.exp_avg_sq)[i]);
            p *= (1.0f - a.lr * a.wd);
            m += (1.0f - a.beta1) * (g - m);
            v = v * a.beta2 + (1.0f - a.beta2) * g * g;
            const float den = sqrtf(v) / a.bc2_sqrt + a.eps;
            p -= step_size * (m / den);
            if (f32_state) {
                reinterpret_cast<float*>(d.exp_avg)[i] = m;
                reinterpret_cast<float*>(d.exp_avg_sq)[i] = v;
            } else {
                reinterpret_cast<__nv_bfloat16*>(d.exp_avg)[i] = __float2bfloat16(m);
                reinterpret_cast<__nv_bfloat16*>(d.exp_avg_sq)[i] = __float2bfloat16(v);
            }
            if (d.master) d.master[i] = p;
            if (d.is_bf16) reinterpret_cast<__nv_bfloat16*>(d.param)[i] = __float2bfloat16(p);
            else reinterpret_cast<float*>(d.param)[i] = p;
        }
    }
}

// plain SGD as torch.optim.SGD(momentum=0) computes it (run_full_tta.py:132-138): g += wd * p ; p -= lr * g
__global__ void __launch_bounds__(OPT_THREADS) sgd_kernel(const b200tta_tensor_desc* __restrict__ descs,
                                                          const float* __restrict__ coef, float grad_scale, float lr, float wd) {
    const b200tta_tensor_desc d = descs[blockIdx.y];
    const float c = (coef ? coef[blockIdx.y] : 1.0f) * grad_scale;
    // the 13.6 B-parameter case: bf16 parameter without master, plain gradient layout -> 8 elements per thread, 128-bit accesses
    if (d.is_bf16 && d.master == nullptr && d.t_rows == 0 && (d.numel & 7) == 0 &&
        ((reinterpret_cast<uintptr_t>(d.param) | reinterpret_cast<uintptr_t>(d.grad)) & 15) == 0) {
        uint4* P8 = reinterpret_cast<uint4*>(d.param);
        const float4* G4 = reinterpret_cast<const float4*>(d.grad);
        const long long n8 = d.numel >> 3;
        for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n8; i += (long long)gridDim.x * OPT_THREADS) {
            uint4 u = P8[i];
            const float4 g0 = G4[2 * i], g1 = G4[2 * i + 1];
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float2 pv = unpack_bf16x2(w[k]);
                pv.x -= lr * (g[2 * k] * c + wd * pv.x);
                pv.y -= lr * (g[2 * k + 1] * c + wd * pv.y);
                w[k] = pack_bf16x2(pv.x, pv.y);
            }
            P8[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return;
    }
    for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < d.numel; i += (long long)gridDim.x * OPT_THREADS) {
        float p = d.master ? d.master[i]
                           : (d.is_bf16 ? __bfloat162float(reinterpret_cast<__nv_bfloat16*>(d.param)[i])
                                        : reinterpret_cast<float*>(d.param)[i]);
        const float g = grad_at(d, i) * c + wd * p;
        p -= lr * g;
        if (d.master) d.master[i] = p;
        if (d.is_bf16) reinterpret_cast<__nv_bfloat16*>(d.param)[i] = __float2bfloat16(p);
        else reinterpret_cast<float*>(d.param)[i] = p;
    }
}

inline dim3 opt_grid(int n, long long max_numel) {
    long long gx = (max_numel + OPT_CHUNK - 1) / OPT_CHUNK;
    if (gx < 1) gx = 1;
    if (gx > 512) gx = 512;      // (full-model TTA: tensors of up to 135 M elements)
    return dim3((unsigned)gx, (unsigned)n);
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200tta_mt_sumsq(const b200tta_tensor_desc* descs_dev, int32_t n, int64_t max_numel, float* sumsq,
                                b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(descs_dev && sumsq && n > 0 && n <= 65535 && max_numel > 0, "mt_sumsq: bad arguments (n=%d)", n);
    sumsq_kernel<<<opt_grid(n, max_numel), OPT_THREADS, 0, (cudaStream_t)stream>>>(descs_dev, sumsq);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_clip_coef(float* coef, float* total_norm_out, const float* sumsq, int32_t n, float max_norm,
                                 int32_t per_tensor, float grad_scale, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(coef && sumsq && n > 0, "clip_coef: bad arguments");
    clip_coef_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(coef, total_norm_out, sumsq, n, max_norm, per_tensor, grad_scale);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_mt_adamw(const b200tta_tensor_desc* descs_dev, int32_t n, int64_t max_numel, const float* coef,
                                float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay,
                                int32_t step, int32_t faithful_bf16, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(descs_dev && n > 0 && n <= 65535 && max_numel > 0 && step >= 1, "mt_adamw: bad arguments (n=%d step=%d)", n, step);
    AdamArgs a;
    a.grad_scale = grad_scale; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
    a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a.faithful_bf16 = faithful_bf16;
    adamw_kernel<<<opt_grid(n, max_numel), OPT_THREADS, 0, (cudaStream_t)stream>>>(descs_dev, coef, a);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_mt_sgd(const b200tta_tensor_desc* descs_dev, int32_t n, int64_t max_numel, const float* coef,
                              float grad_scale, float lr, float weight_decay, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(descs_dev && n > 0 && n <= 65535 && max_numel > 0, "mt_sgd: bad arguments (n=%d)", n);
    sgd_kernel<<<opt_grid(n, max_numel), OPT_THREADS, 0, (cudaStream_t)stream>>>(descs_dev, coef, grad_scale, lr, weight_decay);
    B200_LAUNCHED();
    return B200TTA_OK;
}
