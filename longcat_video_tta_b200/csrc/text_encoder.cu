// UMT5 text-encoder pieces (text half: encode_prompt, delta_experiment/scripts/common.py:228-255, calls
// transformers' UMT5EncoderModel; its arithmetic is restated in oracle/umt5_oracle.py).  The projections and the gated
// feed-forward run on the tcgen05 GEMM (gemm.cu, GEGLU / residual epilogues); what is left is small and latency- or
// HBM-bound at 512 tokens:
//   * t5_rmsnorm_kernel : scale-only RMS norm over d_model (no mean, no bias), one warp per row, row kept in registers
//   * t5_attn_kernel    : softmax(Q K^T + relative-position bias + key mask) V with head_dim 64 and NO 1/sqrt(d) scaling;
//                         4.3 GFLOP per layer at 512 tokens x 64 heads -- 512 CTAs of four warps, FlashAttention-2 style
//                         on warp-level mma.sync (a tcgen05 tile of 128 rows would leave 3/4 of the chip idle here and
//                         the whole kernel is ~20 us either way)
#include "host_common.h"
#include "ptx.cuh"

namespace b200 {
namespace {

constexpr float T5_LOG2E = 1.4426950408889634f;
__device__ __forceinline__ float t5_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float T5_MASKED = -3.4028234663852886e38f;   // what transformers adds for a padded key: finfo(float32).min

// Y = bf16(w * bf16(x * rsqrt(mean(x^2) + eps))): the two roundings transformers' UMT5LayerNorm makes on a bf16 model
__global__ void __launch_bounds__(256) t5_rmsnorm_kernel(__nv_bfloat16* __restrict__ Y, long long ldy,
                                                         const __nv_bfloat16* __restrict__ X, long long ldx,
                                                         const __nv_bfloat16* __restrict__ w, long long rows, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    constexpr int MAXC = 16;                  // 16 x 32 lanes x 8 elements = 4096 columns in registers
    const int chunks = C >> 3;
    const uint4* x4 = reinterpret_cast<const uint4*>(X + row * ldx);
    uint4 v[MAXC];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + i * 32;
        if (c < chunks) {
            v[i] = __ldg(x4 + c);
            const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(u[j]);
                ss = fmaf(f.x, f.x, ss);
                ss = fmaf(f.y, f.y, ss);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float r = rsqrtf(ss / (float)C + eps);
    const uint4* w4 = reinterpret_cast<const uint4*>(w);
    uint4* y4 = reinterpret_cast<uint4*>(Y + row * ldy);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + i * 32;
        if (c < chunks) {
            const uint4 wv = __ldg(w4 + c);
            const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
            const uint32_t wu[4] = {wv.x, wv.y, wv.z, wv.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(u[j]);
                const float2 g = unpack_bf16x2(wu[j]);
                const float2 h = unpack_bf16x2(pack_bf16x2(f.x * r, f.y * r));      // first rounding: .to(bf16)
                o[j] = pack_bf16x2(g.x * h.x, g.y * h.y);
            }
            y4[c] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ attention
constexpr int TA_D = 64, TA_Q = 64, TA_K = 64, TA_LD = TA_D + 8 /* 144-byte rows: conflict-free ldmatrix */, TA_THREADS = 128;

__device__ __forceinline__ void ta_load_tile(__nv_bfloat16 (*dst)[TA_LD], const __nv_bfloat16* src, long long ld, int row0,
                                             int n_rows) {
    // 64 rows x 8 chunks of 16 bytes; rows past the end are zero-filled
    for (int c = threadIdx.x; c < TA_K * (TA_D / 8); c += TA_THREADS) {
        const int r = c >> 3, k = c & 7;
        const bool ok = row0 + r < n_rows;
        cp_async16(&dst[r][k * 8], src + (long long)(ok ? row0 + r : 0) * ld + k * 8, ok);
    }
}

__global__ void __launch_bounds__(TA_THREADS, 4) t5_attn_kernel(__nv_bfloat16* __restrict__ O, long long ldo,
                                                             const __nv_bfloat16* __restrict__ Q, long long ldq,
                                                             const __nv_bfloat16* __restrict__ K, long long ldk,
                                                             const __nv_bfloat16* __restrict__ V, long long ldv,
                                                             const float* __restrict__ rel_bias,
                                                             const int* __restrict__ key_valid, int n_tok, int heads) {
    extern __shared__ __align__(16) uint8_t ta_smem[];
    auto q_s = reinterpret_cast<__nv_bfloat16(*)[TA_LD]>(ta_smem);
    auto k_s = q_s + TA_Q;
    auto v_s = k_s + TA_K;
    const int n_blk = (n_tok + TA_K - 1) / TA_K;
    float* rel_s = reinterpret_cast<float*>(v_s + TA_K);       // [n_blk * 64 + 64]: bias of (key - local query row + 63)
    float* mask_s = rel_s + n_blk * TA_K + 64;                 // [n_blk * 64]: 0 / finfo.min / -inf past the end

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * TA_Q, head = blockIdx.y, b = blockIdx.z;
    const long long tok0 = (long long)b * n_tok;
    Q += tok0 * ldq + head * TA_D;
    K += tok0 * ldk + head * TA_D;
    V += tok0 * ldv + head * TA_D;
    O += tok0 * ldo + head * TA_D;

    ta_load_tile(q_s, Q, ldq, q0, n_tok);
    ta_load_tile(k_s, K, ldk, 0, n_tok);
    ta_load_tile(v_s, V, ldv, 0, n_tok);
    cp_async_commit();
    {   // relative position = key - query; the table holds positions -(n_tok-1) .. n_tok-1
        const float* rel = rel_bias + (long long)head * (2 * n_tok - 1);
        for (int x = threadIdx.x; x < n_blk * TA_K + 64; x += TA_THREADS) {
            const int pos = x - 63 - q0 + n_tok - 1;
            rel_s[x] = (pos >= 0 && pos < 2 * n_tok - 1) ? __ldg(rel + pos) : 0.f;
        }
        for (int j = threadIdx.x; j < n_blk * TA_K; j += TA_THREADS)
            mask_s[j] = j >= n_tok ? -INFINITY : ((key_valid == nullptr || __ldg(key_valid + tok0 + j) != 0) ? 0.f : T5_MASKED);
    }
    cp_async_wait<0>();
    __syncthreads();

    uint32_t qf[4][4];      // A fragments of this warp's 16 query rows, four 16-wide slices of the head dimension
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qf[ks], &q_s[warp * 16 + (lane & 15)][ks * 16 + (lane >> 4) * 8]);

    const int g = lane >> 2, t2 = (lane & 3) * 2;
    const int il0 = warp * 16 + g, il1 = il0 + 8;     // local query rows of this thread's accumulator halves
    float o_acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kb = 0; kb < n_blk; ++kb) {
        if (kb > 0) {
            __syncthreads();                          // every warp is done with the previous K / V block
            ta_load_tile(k_s, K, ldk, kb * TA_K, n_tok);
            ta_load_tile(v_s, V, ldv, kb * TA_K, n_tok);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
        }
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {          // B fragments of two 16-wide head-dim slices per ldmatrix
                uint32_t kf[4];
                ldmatrix_x4(kf, &k_s[nt * 8 + (lane & 7)][kp * 32 + (lane >> 3) * 8]);
                mma_bf16(s[nt], qf[kp * 2], kf[0], kf[1]);
                mma_bf16(s[nt], qf[kp * 2 + 1], kf[2], kf[3]);
            }
        }
        // scores + position bias + key mask (no 1/sqrt(d): T5 folds it into the initialisation), block row maximum
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int jl = kb * TA_K + nt * 8 + t2;   // key index of element 0 / 2; +1 for elements 1 / 3
            const float k0 = mask_s[jl], k1 = mask_s[jl + 1];
            s[nt][0] += rel_s[jl - il0 + 63] + k0;
            s[nt][1] += rel_s[jl + 1 - il0 + 63] + k1;
            s[nt][2] += rel_s[jl - il1 + 63] + k0;
            s[nt][3] += rel_s[jl + 1 - il1 + 63] + k1;
            bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
            bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
        }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
        const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);     // finite: key 0 of block 0 is never past the end
        const float a0 = t5_exp2((m0 - mn0) * T5_LOG2E), a1 = t5_exp2((m1 - mn1) * T5_LOG2E);
        m0 = mn0; m1 = mn1;
        l0 *= a0; l1 *= a1;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) { o_acc[dt][0] *= a0; o_acc[dt][1] *= a0; o_acc[dt][2] *= a1; o_acc[dt][3] *= a1; }
        uint32_t pf[4][4];      // P as A fragments: accumulator tiles 2kk, 2kk+1 form one 16-key slice
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = t5_exp2((s[nt][0] - mn0) * T5_LOG2E), p1 = t5_exp2((s[nt][1] - mn0) * T5_LOG2E);
            const float p2 = t5_exp2((s[nt][2] - mn1) * T5_LOG2E), p3 = t5_exp2((s[nt][3] - mn1) * T5_LOG2E);
            l0 += p0 + p1;
            l1 += p2 + p3;
            pf[nt >> 1][(nt & 1) * 2] = pack_bf16x2(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {          // two 8-wide output slices per ldmatrix
                uint32_t vf[4];
                ldmatrix_x4_trans(vf, &v_s[kk * 16 + (lane & 15)][dp * 16 + (lane >> 4) * 8]);
                mma_bf16(o_acc[dp * 2], pf[kk], vf[0], vf[1]);
                mma_bf16(o_acc[dp * 2 + 1], pf[kk], vf[2], vf[3]);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float r0 = 1.f / l0, r1 = 1.f / l1;
    const int i0 = q0 + il0, i1 = q0 + il1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        if (i0 < n_tok)
            *reinterpret_cast<uint32_t*>(O + (long long)i0 * ldo + dt * 8 + t2) = pack_bf16x2(o_acc[dt][0] * r0, o_acc[dt][1] * r0);
        if (i1 < n_tok)
            *reinterpret_cast<uint32_t*>(O + (long long)i1 * ldo + dt * 8 + t2) = pack_bf16x2(o_acc[dt][2] * r1, o_acc[dt][3] * r1);
    }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200tta_t5_rmsnorm(void* Y, int64_t ldy, const void* X, int64_t ldx, const void* w, int64_t rows, int32_t C,
                                  float eps, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(Y && X && w && rows > 0, "t5_rmsnorm: null argument or no rows");
    B200_REQUIRE(C > 0 && C % 8 == 0 && C <= 4096, "t5_rmsnorm: C=%d must be a multiple of 8, at most 4096", C);
    B200_REQUIRE(aligned16(Y) && aligned16(X) && aligned16(w) && ldy % 8 == 0 && ldx % 8 == 0,
                 "t5_rmsnorm: rows must be 16-byte aligned");
    const int warps = rows >= 148 * 16 ? 8 : 2;     // a 512-token prompt has 512 rows: spread them over every SM
    t5_rmsnorm_kernel<<<(unsigned)((rows + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)Y, ldy, (const __nv_bfloat16*)X, ldx, (const __nv_bfloat16*)w, rows, C, eps);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_t5_attn(void* O, int64_t ldo, const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V,
                               int64_t ldv, const float* rel_bias, const int32_t* key_valid, int32_t n_tok, int32_t heads,
                               int32_t batch, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(O && Q && K && V && rel_bias, "t5_attn: null argument");
    B200_REQUIRE(n_tok > 0 && n_tok <= 8192 && heads > 0 && heads <= 65535 && batch > 0 && batch <= 65535,
                 "t5_attn: n_tok=%d heads=%d batch=%d out of range", n_tok, heads, batch);
    B200_REQUIRE(aligned16(O) && aligned16(Q) && aligned16(K) && aligned16(V) && ldo % 8 == 0 && ldq % 8 == 0 &&
                     ldk % 8 == 0 && ldv % 8 == 0,
                 "t5_attn: head rows (64 bf16) must be 16-byte aligned");
    const int n_blk = (n_tok + TA_K - 1) / TA_K;
    const size_t smem = (size_t)(TA_Q + 2 * TA_K) * TA_LD * 2 + (size_t)(2 * n_blk * TA_K + 64) * 4;
    if (smem > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(t5_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((n_tok + TA_Q - 1) / TA_Q), (unsigned)heads, (unsigned)batch);
    t5_attn_kernel<<<grid, TA_THREADS, smem, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)O, ldo, (const __nv_bfloat16*)Q, ldq, (const __nv_bfloat16*)K, ldk, (const __nv_bfloat16*)V, ldv,
        rel_bias, key_valid, n_tok, heads);
    B200_LAUNCHED();
    return B200TTA_OK;
}
