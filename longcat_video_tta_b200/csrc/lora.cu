// LoRA side products: the rank-r "down" projections and the adapter-gradient reductions over tokens.
// They are HBM-bound (each reads an [n_tok, k] activation once and produces r <= 64 columns), so they use
// warp-level mma.sync.m16n8k16 fed through cp.async-staged shared memory; the frozen-weight GEMMs and the
// rank-r "up" product run on tcgen05 (gemm.cu).  Also hosts the fused entry points lora_linear_fwd/bwd.
#include "host_common.h"
#include "ptx.cuh"

namespace b200 {
namespace {

// ------------------------------------------------------------------------------------ down projection
// T[n_tok, r] = bf16(scale * X[n_tok, k] * Wd^T)   (Wd [r,k]; or Wd [k,r] when wd_t)
// block: 4 warps x 16 rows = 64 rows; K chunks of 64; X tile and Wd chunk double-buffered via cp.async.
constexpr int DN_ROWS = 64, DN_KC = 64, DN_XS = DN_KC + 8 /*pad*/, DN_THREADS = 128;
template <int NT>  // NT = r_pad / 8 n-tiles
__global__ void __launch_bounds__(DN_THREADS) lora_down_kernel(__nv_bfloat16* __restrict__ T, long long ldt,
                                                               const __nv_bfloat16* __restrict__ X, long long ldx,
                                                               const __nv_bfloat16* __restrict__ Wd, int wd_t,
                                                               long long n_tok, long long k, int r, float scale) {
    constexpr int RP = NT * 8;
    __shared__ __align__(16) __nv_bfloat16 xs[2][DN_ROWS][DN_XS];
    __shared__ __align__(16) __nv_bfloat16 ws[2][RP][DN_XS];  // [n][k] (k contiguous)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long row0 = (long long)blockIdx.x * DN_ROWS;
    const int nchunks = (int)((k + DN_KC - 1) / DN_KC);

    auto load_chunk = [&](int buf, int ch) {
        const long long k0 = (long long)ch * DN_KC;
        // X tile: 64 rows x 16 vectors of 8 elements
        for (int i = tid; i < DN_ROWS * (DN_KC / 8); i += DN_THREADS) {
            const int rr = i / (DN_KC / 8), v = i % (DN_KC / 8);
            const long long gr = row0 + rr, gk = k0 + v * 8;
            const bool ok = gr < n_tok && gk < k;
            cp_async16(&xs[buf][rr][v * 8], X + (ok ? gr * ldx + gk : 0), ok);
        }
        if (!wd_t) {
            for (int i = tid; i < RP * (DN_KC / 8); i += DN_THREADS) {
                const int n = i / (DN_KC / 8), v = i % (DN_KC / 8);
                const long long gk = k0 + v * 8;
                const bool ok = n < r && gk < k;
                cp_async16(&ws[buf][n][v * 8], Wd + (ok ? (long long)n * k + gk : 0), ok);
            }
        } else {
            // Wd [k, r]: transpose while staging (plain loads; the chunk is small and L2 resident)
            for (int i = tid; i < RP * DN_KC; i += DN_THREADS) {
                const int kk = i / RP, n = i % RP;
                const long long gk = k0 + kk;
                ws[buf][n][kk] = (n < r && gk < k) ? Wd[gk * r + n] : __float2bfloat16(0.f);
            }
        }
        cp_async_commit();
    };

    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

    load_chunk(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) { load_chunk(buf ^ 1, ch + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < DN_KC / 16; ++ks) {
            uint32_t a[4];
            // A 16x16 tile: rows warp*16 + (lane % 16), k offset ks*16 + (lane / 16) * 8
            ldmatrix_x4(a, &xs[buf][warp * 16 + (lane & 15)][ks * 16 + (lane >> 4) * 8]);
            const int g = lane >> 2, t = lane & 3;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&ws[buf][j * 8 + g][ks * 16 + 2 * t]);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&ws[buf][j * 8 + g][ks * 16 + 2 * t + 8]);
                mma_bf16(acc[j], a, b0, b1);
            }
        }
        __syncthreads();
    }
    const int g = lane >> 2, t = lane & 3;
    const long long r_lo = row0 + warp * 16 + g, r_hi = r_lo + 8;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int col = j * 8 + 2 * t;
        if (col < r) {  // r is even (r % 8 == 0 enforced by the host unless r < 8, handled by padding ldt)
            if (r_lo < n_tok) *reinterpret_cast<uint32_t*>(T + r_lo * ldt + col) = pack_bf16x2(acc[j][0] * scale, acc[j][1] * scale);
            if (r_hi < n_tok) *reinterpret_cast<uint32_t*>(T + r_hi * ldt + col) = pack_bf16x2(acc[j][2] * scale, acc[j][3] * scale);
        }
    }
}

// ------------------------------------------------------------------------------------ adapter gradient
// G[m, r] (f32) += P[n_tok, m]^T * Q[n_tok, r].  grid (ceil(m/64), splits); block 4 warps, warp w owns
// m rows [w*16, w*16+16) of the block's 64; token chunks of 64 staged as [tok][m] / [tok][r] and read with
// ldmatrix.trans (both operands are token-major).
constexpr int GR_M = 64, GR_TC = 64, GR_THREADS = 128;
template <int NT>
__global__ void __launch_bounds__(GR_THREADS) lora_grad_kernel(float* __restrict__ G, const __nv_bfloat16* __restrict__ P,
                                                               long long ldp, const __nv_bfloat16* __restrict__ Q,
                                                               long long ldq, long long n_tok, long long m, int r,
                                                               long long tok_per_split) {
    constexpr int RP = NT * 8;
    __shared__ __align__(16) __nv_bfloat16 ps[2][GR_TC][GR_M + 8];
    __shared__ __align__(16) __nv_bfloat16 qs[2][GR_TC][RP + 8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m0 = (long long)blockIdx.x * GR_M;
    const long long t_begin = (long long)blockIdx.y * tok_per_split;
    const long long t_end = min(n_tok, t_begin + tok_per_split);
    if (t_begin >= t_end) return;
    const int nchunks = (int)((t_end - t_begin + GR_TC - 1) / GR_TC);

    auto load_chunk = [&](int buf, int ch) {
        const long long t0 = t_begin + (long long)ch * GR_TC;
        for (int i = tid; i < GR_TC * (GR_M / 8); i += GR_THREADS) {
            const int tt = i / (GR_M / 8), v = i % (GR_M / 8);
            const long long gt = t0 + tt, gm = m0 + v * 8;
            const bool ok = gt < t_end && gm < m;
            cp_async16(&ps[buf][tt][v * 8], P + (ok ? gt * ldp + gm : 0), ok);
        }
        for (int i = tid; i < GR_TC * (RP / 8); i += GR_THREADS) {
            const int tt = i / (RP / 8), v = i % (RP / 8);
            const long long gt = t0 + tt;
            const bool ok = gt < t_end && v * 8 < r;
            cp_async16(&qs[buf][tt][v * 8], Q + (ok ? gt * ldq + v * 8 : 0), ok);
        }
        cp_async_commit();
    };

    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

    load_chunk(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) { load_chunk(buf ^ 1, ch + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < GR_TC / 16; ++ks) {
            // A = P^T tile (16 m x 16 tok), stored [tok][m]: ldmatrix.trans of four 8x8 blocks
            //   a0: (m 0-7, k 0-7)  a1: (m 8-15, k 0-7)  a2: (m 0-7, k 8-15)  a3: (m 8-15, k 8-15)
            uint32_t a[4];
            {
                const int blk = lane >> 3, rr = lane & 7;
                const int tok = ks * 16 + (blk >> 1) * 8 + rr;
                const int mm = warp * 16 + (blk & 1) * 8;
                ldmatrix_x4_trans(a, &ps[buf][tok][mm]);
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                uint32_t b0, b1;
                ldmatrix_x2_trans(b0, b1, &qs[buf][ks * 16 + (lane & 15)][j * 8]);
                mma_bf16(acc[j], a, b0, b1);
            }
        }
        __syncthreads();
    }
    const int g = lane >> 2, t = lane & 3;
    const long long m_lo = m0 + warp * 16 + g, m_hi = m_lo + 8;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int col = j * 8 + 2 * t;
        if (col < r) {
            if (m_lo < m) { atomicAdd(G + m_lo * r + col, acc[j][0]); if (col + 1 < r) atomicAdd(G + m_lo * r + col + 1, acc[j][1]); }
            if (m_hi < m) { atomicAdd(G + m_hi * r + col, acc[j][2]); if (col + 1 < r) atomicAdd(G + m_hi * r + col + 1, acc[j][3]); }
        }
    }
}

}  // namespace
}  // namespace b200

using namespace b200;

#define NT_DISPATCH(RP, CALL)                                    \
    switch ((RP) / 8) {                                          \
        case 1: { constexpr int NT = 1; CALL; } break;           \
        case 2: { constexpr int NT = 2; CALL; } break;           \
        case 3: { constexpr int NT = 3; CALL; } break;           \
        case 4: { constexpr int NT = 4; CALL; } break;           \
        case 6: { constexpr int NT = 6; CALL; } break;           \
        case 8: { constexpr int NT = 8; CALL; } break;           \
        default: b200::set_last_error("lora: rank %d unsupported (pad to 8,16,24,32,48,64)", r); return B200TTA_EINVAL; \
    }

extern "C" int b200tta_lora_down(void* T, int64_t ldt, const void* X, int64_t ldx, const void* Wd, int32_t wd_t,
                                 int64_t n_tok, int64_t k, int32_t r, float scale, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(T && X && Wd && n_tok > 0 && k > 0 && r > 0 && r <= 64 && r % 2 == 0, "lora_down: bad arguments (r=%d)", r);
    B200_REQUIRE(aligned16(X) && ldx % 8 == 0 && k % 8 == 0 && ldt % 2 == 0 && (wd_t || aligned16(Wd)),
                 "lora_down: alignment (ldx=%lld k=%lld ldt=%lld)", (long long)ldx, (long long)k, (long long)ldt);
    const int rp = (r + 7) / 8 * 8;
    const int grid = (int)((n_tok + DN_ROWS - 1) / DN_ROWS);
    NT_DISPATCH(rp, (lora_down_kernel<NT><<<grid, DN_THREADS, 0, (cudaStream_t)stream>>>(
                        (__nv_bfloat16*)T, ldt, (const __nv_bfloat16*)X, ldx, (const __nv_bfloat16*)Wd, wd_t, n_tok, k, r,
                        scale)));
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_lora_grad(float* G, const void* P, int64_t ldp, const void* Q, int64_t ldq, int64_t n_tok,
                                 int64_t m, int32_t r, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(G && P && Q && n_tok > 0 && m > 0 && r > 0 && r <= 64, "lora_grad: bad arguments (r=%d)", r);
    B200_REQUIRE(aligned16(P) && aligned16(Q) && ldp % 8 == 0 && ldq % 8 == 0 && m % 8 == 0 && r % 8 == 0,
                 "lora_grad: alignment (ldp=%lld ldq=%lld m=%lld r=%d)", (long long)ldp, (long long)ldq, (long long)m, r);
    const int mt = (int)((m + GR_M - 1) / GR_M);
    int splits = (148 * 6 + mt - 1) / mt;
    const long long max_splits = (n_tok + GR_TC - 1) / GR_TC;
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    long long tps = (n_tok + splits - 1) / splits;
    tps = (tps + GR_TC - 1) / GR_TC * GR_TC;
    dim3 grid(mt, (unsigned)((n_tok + tps - 1) / tps));
    NT_DISPATCH(r, (lora_grad_kernel<NT><<<grid, GR_THREADS, 0, (cudaStream_t)stream>>>(
                       G, (const __nv_bfloat16*)P, ldp, (const __nv_bfloat16*)Q, ldq, n_tok, m, r, tps)));
    B200_LAUNCHED();
    return B200TTA_OK;
}

// ------------------------------------------------------------------------------------ fused LoRA linear
extern "C" int b200tta_lora_linear_fwd(const void* X, int64_t ldx, const void* W, const void* W_hi, const void* A,
                                       const void* B, void* XA, int64_t n_tok, int64_t in_features,
                                       int64_t out_features, int32_t r, float scale, const b200tta_gemm_epi* epi,
                                       b200tta_stream_t stream) {
    B200_REQUIRE(X && W && epi, "lora_linear_fwd: null argument");
    b200tta_gemm_seg segs[2];
    memset(segs, 0, sizeof(segs));
    segs[0].a = X; segs[0].lda = ldx; segs[0].b = W; segs[0].ldb = in_features; segs[0].b_hi = W_hi; segs[0].k = in_features;
    int nseg = 1;
    const int64_t n_total = W_hi ? 2 * out_features : out_features;
    if (r > 0) {
        B200_REQUIRE(A && B && XA, "lora_linear_fwd: rank %d given but A/B/XA null", r);
        B200_REQUIRE(r % 8 == 0 && r <= 64, "lora_linear_fwd: rank %d must be a multiple of 8, <= 64 (pad on the host)", r);
        B200_REQUIRE(!W_hi, "lora_linear_fwd: LoRA on a co-tiled (w1|w3) linear is not supported; run them unfused");
        if (int rc = b200tta_lora_down(XA, r, X, ldx, A, 0, n_tok, in_features, r, scale, stream)) return rc;
        segs[1].a = XA; segs[1].lda = r; segs[1].b = B; segs[1].ldb = r; segs[1].k = r;
        nseg = 2;
    }
    return b200tta_gemm(n_tok, n_total, segs, nseg, epi, stream);
}

extern "C" int b200tta_lora_linear_bwd(const void* dY, int64_t lddy, const void* X, int64_t ldx, const void* W,
                                       const void* A, const void* B, const void* XA, void* U, float* dA_acc,
                                       float* dB_acc, int64_t n_tok, int64_t in_features, int64_t out_features,
                                       int32_t r, float scale, const b200tta_gemm_epi* epi, b200tta_stream_t stream) {
    B200_REQUIRE(dY && W, "lora_linear_bwd: null argument");
    if (r > 0) {
        B200_REQUIRE(A && B && XA && U && dA_acc && dB_acc && X, "lora_linear_bwd: rank %d given but an adapter buffer is null", r);
        B200_REQUIRE(r % 8 == 0 && r <= 64, "lora_linear_bwd: rank %d must be a multiple of 8, <= 64", r);
        // U = bf16(scale * dY B)   [n_tok, r]
        if (int rc = b200tta_lora_down(U, r, dY, lddy, B, 1, n_tok, out_features, r, scale, stream)) return rc;
        // dB += dY^T XA ; dA += U^T X
        if (int rc = b200tta_lora_grad(dB_acc, dY, lddy, XA, r, n_tok, out_features, r, stream)) return rc;
        // dA is [r, in]: computed as (X^T U)^T -> accumulate into a [in, r] layout?  No: G[m=in, r] is the
        // transpose of dA.  The host keeps dA_acc as [in, r] ("A^T gradient") and transposes once per step.
        if (int rc = b200tta_lora_grad(dA_acc, X, ldx, U, r, n_tok, in_features, r, stream)) return rc;
    }
    if (epi == nullptr || epi->d == nullptr) return B200TTA_OK;  // adapter gradients only
    b200tta_gemm_seg segs[2];
    memset(segs, 0, sizeof(segs));
    segs[0].a = dY; segs[0].lda = lddy; segs[0].b = W; segs[0].ldb = in_features; segs[0].k = out_features; segs[0].b_mn_major = 1;
    int nseg = 1;
    if (r > 0) {
        segs[1].a = U; segs[1].lda = r; segs[1].b = A; segs[1].ldb = in_features; segs[1].k = r; segs[1].b_mn_major = 1;
        nseg = 2;
    }
    return b200tta_gemm(n_tok, in_features, segs, nseg, epi, stream);
}
