// Flash-style attention backward for sm_100a, head_dim 128, bf16 -- recomputes S = Q K^T from Q, K and the saved
// log-sum-exp (no S/P tensors are ever stored).  Three kernels, all deterministic (no atomics):
//   delta_kernel : delta[h, i] = sum_d dO[i,h,d] * O[i,h,d]
//   dq_kernel    : one CTA per (128-row query tile, head), loop over K/V blocks:
//                    S = Q K^T, dP = dO V^T (SS) -> dS = P o (dP - delta) (bf16 over dP in TMEM) -> dQ += dS K (TS)
//   dkv_kernel   : one CTA per (128-row K/V block, head), loop over the query tiles that see it:
//                    S^T = K Q^T, dP^T = V dO^T (SS) -> P^T, dS^T (bf16 in TMEM) -> dV += P^T dO, dK += dS^T Q (TS)
// The same shared-memory tile serves as a K-major operand (contraction over d) and as an MN-major operand
// (contraction over tokens) -- only the UMMA descriptor differs, nothing is transposed in memory.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 one thread per TMEM lane (row).
#include "host_common.h"
#include "ptx.cuh"

namespace b200 {
namespace {

constexpr int D = 128, BT = 128;
constexpr int TILE_BYTES = 128 * 128 * 2, SUB_BYTES = TILE_BYTES / 2;
constexpr int STAGES = 2;
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BYTES = (2 + 2 * STAGES) * TILE_BYTES + 1024 + 256 + 2 * 2 * BT * 4;
constexpr int MAX_SEGS = 4;
constexpr float LOG2E = 1.4426950408889634f;

struct BwdParams {
    CUtensorMap tma_q, tma_k, tma_v, tma_do;
    __nv_bfloat16 *dQ, *dK, *dV;
    long long lddq, lddk, lddv;
    const float* LSE;
    const float* delta;
    int n_q, n_kv, heads;
    float scale, scale_log2;
    int n_seg;
    int seg_q_begin[MAX_SEGS], seg_q_end[MAX_SEGS], seg_kv_len[MAX_SEGS], seg_item0[MAX_SEGS + 1];
};

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, uint32_t tmem_addr, float mul, bool do_store) {
#pragma unroll 1
    for (int c = 0; c < D / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_addr + c * 32, r);
        tmem_ld_wait();
        if (do_store) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(r[i * 8 + 0]) * mul, __uint_as_float(r[i * 8 + 1]) * mul);
                o.y = pack_bf16x2(__uint_as_float(r[i * 8 + 2]) * mul, __uint_as_float(r[i * 8 + 3]) * mul);
                o.z = pack_bf16x2(__uint_as_float(r[i * 8 + 4]) * mul, __uint_as_float(r[i * 8 + 5]) * mul);
                o.w = pack_bf16x2(__uint_as_float(r[i * 8 + 6]) * mul, __uint_as_float(r[i * 8 + 7]) * mul);
                *reinterpret_cast<uint4*>(dst + c * 32 + i * 8) = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------------ delta = rowsum(dO * O)
__global__ void __launch_bounds__(256) delta_kernel(float* __restrict__ delta, const __nv_bfloat16* __restrict__ dO,
                                                    long long lddo, const __nv_bfloat16* __restrict__ O, long long ldo,
                                                    int n_q, int heads) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per (row, head)
    const int lane = threadIdx.x & 31;
    if (w >= (long long)n_q * heads) return;
    const int row = (int)(w / heads), h = (int)(w % heads);
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(dO + (long long)row * lddo + h * D + lane * 4));
    const uint2 b = __ldg(reinterpret_cast<const uint2*>(O + (long long)row * ldo + h * D + lane * 4));
    const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y);
    float s = a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) delta[(long long)h * n_q + row] = s;
}

// ------------------------------------------------------------------------------------ dQ
__global__ void __launch_bounds__(NUM_THREADS, 1) dq_kernel(const __grid_constant__ BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_s = smem;
    uint8_t* do_s = q_s + TILE_BYTES;
    uint8_t* k_s = do_s + TILE_BYTES;              // [STAGES]
    uint8_t* v_s = k_s + STAGES * TILE_BYTES;      // [STAGES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + STAGES * TILE_BYTES);
    uint64_t* qdo_full = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = k_full + STAGES;
    uint64_t* v_full = k_empty + STAGES;
    uint64_t* v_empty = v_full + STAGES;
    uint64_t* sdp_full = v_empty + STAGES;
    uint64_t* ds_full = sdp_full + 1;
    uint64_t* dq_done = ds_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(dq_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x, head = blockIdx.y;
    int seg = 0;
    while (seg + 1 < p.n_seg && item >= p.seg_item0[seg + 1]) ++seg;
    const int q0 = p.seg_q_begin[seg] + (item - p.seg_item0[seg]) * BT;
    const int q_end = p.seg_q_end[seg];
    const int kv_len = p.seg_kv_len[seg];
    const int n_blocks = (kv_len + BT - 1) / BT;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tma_q); tma_prefetch_desc(&p.tma_k); tma_prefetch_desc(&p.tma_v); tma_prefetch_desc(&p.tma_do);
        mbar_init(qdo_full, 1);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        }
        mbar_init(sdp_full, 1); mbar_init(ds_full, BT); mbar_init(dq_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t t_s = tmem_base, t_dp = tmem_base + 128, t_dq = tmem_base + 256;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(qdo_full, 2 * TILE_BYTES);
            for (int c = 0; c < 2; ++c) {
                tma_load_2d(q_s + c * SUB_BYTES, &p.tma_q, qdo_full, head * D + c * 64, q0);
                tma_load_2d(do_s + c * SUB_BYTES, &p.tma_do, qdo_full, head * D + c * 64, q0);
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_blocks; ++j) {
                mbar_wait(&k_empty[stage], phase ^ 1u);
                mbar_arrive_expect_tx(&k_full[stage], TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(k_s + stage * TILE_BYTES + c * SUB_BYTES, &p.tma_k, &k_full[stage], head * D + c * 64, j * BT);
                mbar_wait(&v_empty[stage], phase ^ 1u);
                mbar_arrive_expect_tx(&v_full[stage], TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(v_s + stage * TILE_BYTES + c * SUB_BYTES, &p.tma_v, &v_full[stage], head * D + c * 64, j * BT);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_kk = umma_idesc_bf16(BT, BT, 0, 0);
            constexpr uint32_t idesc_mn = umma_idesc_bf16(BT, D, 0, 1);
            mbar_wait(qdo_full, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_blocks; ++j) {
                const uint32_t qa = smem_u32(q_s), da = smem_u32(do_s);
                const uint32_t ka = smem_u32(k_s + stage * TILE_BYTES), va = smem_u32(v_s + stage * TILE_BYTES);
                mbar_wait(&k_full[stage], phase);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_ss(t_s, umma_desc_kmajor(qa + c * SUB_BYTES + ks * 32),
                                umma_desc_kmajor(ka + c * SUB_BYTES + ks * 32), idesc_kk, (c | ks) ? 1u : 0u);
                mbar_wait(&v_full[stage], phase);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_ss(t_dp, umma_desc_kmajor(da + c * SUB_BYTES + ks * 32),
                                umma_desc_kmajor(va + c * SUB_BYTES + ks * 32), idesc_kk, (c | ks) ? 1u : 0u);
                umma_commit(sdp_full);
                umma_commit(&v_empty[stage]);
                mbar_wait(ds_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < BT / 16; ++ks)
                    umma_ts(t_dq, t_dp + ks * 8, umma_desc_mnmajor(ka + ks * 2048, SUB_BYTES), idesc_mn, (j | ks) ? 1u : 0u);
                umma_commit(&k_empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            umma_commit(dq_done);
        }
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int q_row = q0 + row;
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const bool row_ok = q_row < q_end;
        const float lse2 = row_ok ? p.LSE[(long long)head * p.n_q + q_row] * LOG2E : 0.f;
        const float dlt = row_ok ? p.delta[(long long)head * p.n_q + q_row] : 0.f;
        for (int j = 0; j < n_blocks; ++j) {
            mbar_wait(sdp_full, j & 1);
            tc_fence_after();
            const int valid = kv_len - j * BT;
#pragma unroll 1
            for (int c = 0; c < BT / 32; ++c) {
                uint32_t s[32], dp[32], pk[16];
                tmem_ld_32x32b_x32(t_s + lane_addr + c * 32, s);
                tmem_ld_32x32b_x32(t_dp + lane_addr + c * 32, dp);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = fast_exp2(fmaf(__uint_as_float(s[i]), p.scale_log2, -lse2));
                    float p1 = fast_exp2(fmaf(__uint_as_float(s[i + 1]), p.scale_log2, -lse2));
                    if (!row_ok || c * 32 + i >= valid) p0 = 0.f;
                    if (!row_ok || c * 32 + i + 1 >= valid) p1 = 0.f;
                    pk[i >> 1] = pack_bf16x2(p0 * (__uint_as_float(dp[i]) - dlt), p1 * (__uint_as_float(dp[i + 1]) - dlt));
                }
                tmem_st_32x32b_x16(t_dp + lane_addr + c * 16, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(ds_full);
        }
        mbar_wait(dq_done, 0);
        tc_fence_after();
        store_row_bf16(p.dQ + (long long)q_row * p.lddq + head * D, t_dq + lane_addr, p.scale, row_ok);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// ------------------------------------------------------------------------------------ dK, dV
__global__ void __launch_bounds__(NUM_THREADS, 1) dkv_kernel(const __grid_constant__ BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* k_s = smem;
    uint8_t* v_s = k_s + TILE_BYTES;
    uint8_t* q_s = v_s + TILE_BYTES;               // [STAGES]
    uint8_t* do_s = q_s + STAGES * TILE_BYTES;     // [STAGES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(do_s + STAGES * TILE_BYTES);
    uint64_t* kv_full = bars;
    uint64_t* q_full = bars + 1;
    uint64_t* q_empty = q_full + STAGES;
    uint64_t* do_full = q_empty + STAGES;
    uint64_t* do_empty = do_full + STAGES;
    uint64_t* sdp_full = do_empty + STAGES;
    uint64_t* pds_full = sdp_full + 1;
    uint64_t* dkv_done = pds_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(dkv_done + 1);
    float* stat_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2 buffers][2][BT]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kv0 = blockIdx.x * BT, head = blockIdx.y;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tma_q); tma_prefetch_desc(&p.tma_k); tma_prefetch_desc(&p.tma_v); tma_prefetch_desc(&p.tma_do);
        mbar_init(kv_full, 1);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&do_full[i], 1); mbar_init(&do_empty[i], 1);
        }
        mbar_init(sdp_full, 1); mbar_init(pds_full, BT); mbar_init(dkv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t t_s = tmem_base, t_dp = tmem_base + 128, t_dv = tmem_base + 256, t_dk = tmem_base + 384;

    // every role walks the same list of query tiles: for each segment that can see this K/V block, its 128-row tiles
    auto for_each_tile = [&](auto&& fn) {
        int it = 0;
        for (int s = 0; s < p.n_seg; ++s) {
            if (kv0 >= p.seg_kv_len[s]) continue;
            for (int q0 = p.seg_q_begin[s]; q0 < p.seg_q_end[s]; q0 += BT) fn(it++, s, q0);
        }
        return it;
    };

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * TILE_BYTES);
            for (int c = 0; c < 2; ++c) {
                tma_load_2d(k_s + c * SUB_BYTES, &p.tma_k, kv_full, head * D + c * 64, kv0);
                tma_load_2d(v_s + c * SUB_BYTES, &p.tma_v, kv_full, head * D + c * 64, kv0);
            }
            for_each_tile([&](int it, int, int q0) {
                const int stage = it % STAGES;
                const uint32_t phase = (it / STAGES) & 1;
                mbar_wait(&q_empty[stage], phase ^ 1u);
                mbar_arrive_expect_tx(&q_full[stage], TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(q_s + stage * TILE_BYTES + c * SUB_BYTES, &p.tma_q, &q_full[stage], head * D + c * 64, q0);
                mbar_wait(&do_empty[stage], phase ^ 1u);
                mbar_arrive_expect_tx(&do_full[stage], TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(do_s + stage * TILE_BYTES + c * SUB_BYTES, &p.tma_do, &do_full[stage], head * D + c * 64, q0);
            });
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_kk = umma_idesc_bf16(BT, BT, 0, 0);
            constexpr uint32_t idesc_mn = umma_idesc_bf16(BT, D, 0, 1);
            mbar_wait(kv_full, 0);
            const uint32_t ka = smem_u32(k_s), va = smem_u32(v_s);
            const int n_it = for_each_tile([&](int it, int, int) {
                const int stage = it % STAGES;
                const uint32_t phase = (it / STAGES) & 1;
                const uint32_t qa = smem_u32(q_s + stage * TILE_BYTES), da = smem_u32(do_s + stage * TILE_BYTES);
                mbar_wait(&q_full[stage], phase);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_ss(t_s, umma_desc_kmajor(ka + c * SUB_BYTES + ks * 32),
                                umma_desc_kmajor(qa + c * SUB_BYTES + ks * 32), idesc_kk, (c | ks) ? 1u : 0u);
                mbar_wait(&do_full[stage], phase);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_ss(t_dp, umma_desc_kmajor(va + c * SUB_BYTES + ks * 32),
                                umma_desc_kmajor(da + c * SUB_BYTES + ks * 32), idesc_kk, (c | ks) ? 1u : 0u);
                umma_commit(sdp_full);
                mbar_wait(pds_full, it & 1);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < BT / 16; ++ks)
                    umma_ts(t_dv, t_s + ks * 8, umma_desc_mnmajor(da + ks * 2048, SUB_BYTES), idesc_mn, (it | ks) ? 1u : 0u);
                umma_commit(&do_empty[stage]);
#pragma unroll
                for (int ks = 0; ks < BT / 16; ++ks)
                    umma_ts(t_dk, t_dp + ks * 8, umma_desc_mnmajor(qa + ks * 2048, SUB_BYTES), idesc_mn, (it | ks) ? 1u : 0u);
                umma_commit(&q_empty[stage]);
            });
            (void)n_it;
            umma_commit(dkv_done);
        }
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;  // K/V row inside the block
        const int kv_row = kv0 + row;
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const int tid = threadIdx.x - 64;  // 0..127
        const int n_it = for_each_tile([&](int it, int s, int q0) {
            // stage LSE (log2 units) and delta of this query tile: thread i loads column i
            float* lse_s = stat_s + (it & 1) * 2 * BT;
            float* dl_s = lse_s + BT;
            {
                const int qc = q0 + tid;
                const bool ok = qc < p.seg_q_end[s];
                lse_s[tid] = ok ? p.LSE[(long long)head * p.n_q + qc] * LOG2E : INFINITY;  // exp2(-inf) = 0 masks the column
                dl_s[tid] = ok ? p.delta[(long long)head * p.n_q + qc] : 0.f;
            }
            named_bar_sync(1, BT);
            const bool row_ok = kv_row < p.seg_kv_len[s];
            mbar_wait(sdp_full, it & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BT / 32; ++c) {
                uint32_t st[32], dp[32], pk[16], dk[16];
                tmem_ld_32x32b_x32(t_s + lane_addr + c * 32, st);
                tmem_ld_32x32b_x32(t_dp + lane_addr + c * 32, dp);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float2 l2 = *reinterpret_cast<const float2*>(lse_s + c * 32 + i);
                    const float2 dl = *reinterpret_cast<const float2*>(dl_s + c * 32 + i);
                    float p0 = fast_exp2(fmaf(__uint_as_float(st[i]), p.scale_log2, -l2.x));
                    float p1 = fast_exp2(fmaf(__uint_as_float(st[i + 1]), p.scale_log2, -l2.y));
                    if (!row_ok) { p0 = 0.f; p1 = 0.f; }
                    pk[i >> 1] = pack_bf16x2(p0, p1);
                    dk[i >> 1] = pack_bf16x2(p0 * (__uint_as_float(dp[i]) - dl.x), p1 * (__uint_as_float(dp[i + 1]) - dl.y));
                }
                tmem_st_32x32b_x16(t_s + lane_addr + c * 16, pk);
                tmem_st_32x32b_x16(t_dp + lane_addr + c * 16, dk);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(pds_full);
        });
        mbar_wait(dkv_done, 0);
        tc_fence_after();
        const bool ok = kv_row < p.n_kv;
        if (n_it > 0) {
            store_row_bf16(p.dV + (long long)kv_row * p.lddv + head * D, t_dv + lane_addr, 1.0f, ok);
            store_row_bf16(p.dK + (long long)kv_row * p.lddk + head * D, t_dk + lane_addr, p.scale, ok);
        } else if (ok) {  // a K/V block no query sees: gradients are zero
            for (int c = 0; c < D / 8; ++c) {
                *reinterpret_cast<uint4*>(p.dV + (long long)kv_row * p.lddv + head * D + c * 8) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(p.dK + (long long)kv_row * p.lddk + head * D + c * 8) = make_uint4(0, 0, 0, 0);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200tta_attn_bwd(void* dQ, int64_t lddq, void* dK, int64_t lddk, void* dV, int64_t lddv, const void* dO,
                                int64_t lddo, const void* O, int64_t ldo, const float* LSE, float* delta, const void* Q,
                                int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, int32_t n_q,
                                int32_t n_kv, int32_t heads, float softmax_scale, const b200tta_attn_seg* segs,
                                int32_t n_seg, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(dQ && dK && dV && dO && O && LSE && delta && Q && K && V && n_q > 0 && n_kv > 0 && heads > 0,
                 "attn_bwd: null/empty argument");
    B200_REQUIRE(n_seg >= 1 && n_seg <= MAX_SEGS && segs, "attn_bwd: n_seg=%d not in [1,%d]", n_seg, MAX_SEGS);
    B200_REQUIRE(aligned16(dQ) && aligned16(dK) && aligned16(dV) && aligned16(dO) && aligned16(O) && aligned16(Q) &&
                     aligned16(K) && aligned16(V) && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 && lddo % 8 == 0 &&
                     ldo % 8 == 0 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0,
                 "attn_bwd: tensors must be 16-byte aligned [tokens, heads, 128] views");
    cudaStream_t st = (cudaStream_t)stream;
    BwdParams p;
    memset(&p, 0, sizeof(p));
    if (int rc = make_tmap_2d_bf16(&p.tma_q, Q, (uint64_t)heads * D, (uint64_t)n_q, (uint64_t)ldq * 2, 64, BT)) return rc;
    if (int rc = make_tmap_2d_bf16(&p.tma_do, dO, (uint64_t)heads * D, (uint64_t)n_q, (uint64_t)lddo * 2, 64, BT)) return rc;
    if (int rc = make_tmap_2d_bf16(&p.tma_k, K, (uint64_t)heads * D, (uint64_t)n_kv, (uint64_t)ldk * 2, 64, BT)) return rc;
    if (int rc = make_tmap_2d_bf16(&p.tma_v, V, (uint64_t)heads * D, (uint64_t)n_kv, (uint64_t)ldv * 2, 64, BT)) return rc;
    p.dQ = (__nv_bfloat16*)dQ; p.dK = (__nv_bfloat16*)dK; p.dV = (__nv_bfloat16*)dV;
    p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
    p.LSE = LSE; p.delta = delta; p.n_q = n_q; p.n_kv = n_kv; p.heads = heads;
    p.scale = softmax_scale; p.scale_log2 = softmax_scale * LOG2E;
    p.n_seg = n_seg;
    int items = 0;
    for (int s = 0; s < n_seg; ++s) {
        B200_REQUIRE(segs[s].q_begin >= 0 && segs[s].q_end > segs[s].q_begin && segs[s].q_end <= n_q &&
                         segs[s].kv_len > 0 && segs[s].kv_len <= n_kv,
                     "attn_bwd: bad segment %d", s);
        p.seg_q_begin[s] = segs[s].q_begin; p.seg_q_end[s] = segs[s].q_end; p.seg_kv_len[s] = segs[s].kv_len;
        p.seg_item0[s] = items;
        items += (segs[s].q_end - segs[s].q_begin + BT - 1) / BT;
    }
    p.seg_item0[n_seg] = items;
    static bool attr = false;
    if (!attr) {
        B200_CUDA(cudaFuncSetAttribute(dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        B200_CUDA(cudaFuncSetAttribute(dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr = true;
    }
    {
        const long long warps = (long long)n_q * heads;
        delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(delta, (const __nv_bfloat16*)dO, lddo,
                                                                          (const __nv_bfloat16*)O, ldo, n_q, heads);
        B200_LAUNCHED();
    }
    dq_kernel<<<dim3(items, heads), NUM_THREADS, SMEM_BYTES, st>>>(p);
    B200_LAUNCHED();
    dkv_kernel<<<dim3((n_kv + BT - 1) / BT, heads), NUM_THREADS, SMEM_BYTES, st>>>(p);
    B200_LAUNCHED();
    return B200TTA_OK;
}
