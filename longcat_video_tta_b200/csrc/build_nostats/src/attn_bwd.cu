// Flash-style attention backward for sm_100a, head_dim 128, bf16 -- recomputes S = Q K^T from Q, K and the saved
// log-sum-exp (no S/P tensors are ever stored).  Three kernels, no atomics:
//   delta_kernel : delta[h, i] = sum_d dO[i,h,d] * O[i,h,d]
//   dq_kernel    : one CTA per (128-row query tile, head); Q and dO live in TENSOR MEMORY as bf16 A operands; K/V stream
//                  through shared memory in 128-row blocks; every MMA is TS-form with N = 128:
//                    S = Q K^T, dP = dO V^T -> dS = P o (dP - delta) (bf16 over dP in TMEM) -> dQ += dS K
//   dkv_kernel   : one CTA per (128-row K/V block, head); K, V resident in shared memory; Q/dO stream through in 64-row
//                  sub-tiles together with their LSE / delta values (staged by the producer warp):
//                    S^T = K Q^T, dP^T = V dO^T (SS, N=64) -> P^T, dS^T (bf16 in TMEM) -> dV += P^T dO, dK += dS^T Q (TS);
//                  S^T and dP^T are signalled separately (the exponentials overlap the dP^T MMAs) and dV is issued
//                  as soon as P^T is stored, before dS^T exists
// Why the A operands sit in TMEM where they fit: an SS-form tcgen05.mma re-reads its 128 x 16 A slice (4 KB) and its
// N x 16 B slice from shared memory for every instruction, and shared memory delivers 128 B/clk -- measured
// (scratch/mma_rate.cu): SS N=64 costs 48 clk, N=32 40 clk, while the TS form (A in TMEM) runs at the tensor rate N/2.
// Pipelining: dkv double-buffers S^T/dP^T in TMEM (2 x 128 columns) and two compute warpgroups alternate sub-tiles, so
// the exponentials of sub-tile u overlap the tensor-core work of sub-tiles u-1 / u+1; dq keeps single 128-wide S and dP
// buffers and releases S as soon as it is loaded (scratch/variants/README.md has the measured alternatives).
// The bf16 P / dS tiles overwrite the fp32 S / dP columns they were computed from; each thread's stores land inside
// the column range it loaded itself ([32h, 32h+16) of [32h, 32h+32)), so no cross-thread barrier is needed.
// The same shared-memory tile serves as a K-major operand (contraction over d) and as an MN-major operand
// (contraction over tokens) -- only the UMMA descriptor differs, nothing is transposed in memory.
// Warp roles (dkv; dq uses one issuer, whose polls sit behind >= 512 clk of queued MMAs, and one compute group of 16
// warps): warp 0 TMA producer; warps 1 and 2 MMA issuers, one per S/dP buffer (= per compute group): any
// shared-memory access of an issuing thread (an mbarrier poll included) queues behind its own outstanding
// tcgen05.mma and costs ~100 clk during which a single issuer leaves the tensor pipe idle (measured,
// scratch/mma_issue.cu: 1025 clk per sub-block with one issuer and three waits, 768 = pipe time with two issuers
// alternating).  Each issuer runs warp-converged with one elected lane issuing -- a divergent `if (lane == 0)` region
// costs ~100 clk per MMA in uniform-register waterfall loops.  Warps 3-10 compute group 0, warps 11-18 compute group 1
// (two threads per TMEM lane = row, 32 columns each).  The dQ / dK / dV accumulators are zeroed up front and every MMA
// accumulates, so the two issuers need no ordering between each other; the fp32 summation order of the sub-block
// contributions is therefore not fixed run to run (differences at rounding level).
#include "host_common.h"
#include "ptx.cuh"
#include <stdlib.h>
#include <type_traits>

#ifndef B200TTA_ATTN_DEBUG
#define B200TTA_ATTN_DEBUG 0
#endif
#ifndef B200TTA_EXPERIMENT_NO_STATS_LDS
#define B200TTA_EXPERIMENT_NO_STATS_LDS 0
#endif
#ifndef B200TTA_DQ_DIRECT_RED
#define B200TTA_DQ_DIRECT_RED 0
#endif

namespace b200 {
namespace {

constexpr int D = 128, BT = 128, SUB = 64;
constexpr int TILE_BYTES = BT * D * 2;        // resident [128 x 128] operand: two [128 x 64] swizzled halves
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr int SUBT_BYTES = SUB * D * 2;       // streamed [64 x 128] operand: two [64 x 64] swizzled halves
constexpr int SUBH_BYTES = SUBT_BYTES / 2;
constexpr int DKV_STAGES = 4;                 // dkv: K, V resident (64 KB) + Q/dO sub-tiles
constexpr int GROUP_THREADS = 256;            // per TMEM buffer: two threads per row, 32 columns each
constexpr int NUM_THREADS = 96 + 2 * GROUP_THREADS;   // producer warp, two MMA issuer warps, two compute groups
constexpr int DKV_SMEM_BYTES = 2 * TILE_BYTES + 2 * DKV_STAGES * SUBT_BYTES + 1024 + 512 + DKV_STAGES * 2 * SUB * 4;
constexpr int MAX_SEGS = 4;
constexpr float LOG2E = 1.4426950408889634f;

struct BwdParams {
    CUtensorMap tma_k128, tma_v128;                        // box {64, 128}
    CUtensorMap tma_q64, tma_do64;                         // box {64, 64}
    const __nv_bfloat16 *Q, *dO;
    long long ldq, lddo;
    __nv_bfloat16 *dQ, *dK, *dV;
    long long lddq, lddk, lddv;
    const float* LSE;
    const float* delta;
    int n_q, n_kv, heads;
    float scale, scale_log2;
    int n_seg;
    int seg_q_begin[MAX_SEGS], seg_q_end[MAX_SEGS], seg_kv_len[MAX_SEGS], seg_item0[MAX_SEGS + 1];
    // block-sparse variant (NULL = dense): CSR lists over 128-token blocks, tokens in block-major order, one segment.
    //   q_off [heads * n_qblk + 1] / q_idx : key blocks attended by (head, query block), ascending
    //   k_off [heads * n_kblk + 1] / k_idx : query blocks that attend (head, key block), ascending (the transpose)
    const int* q_off; const int* q_idx; const int* k_off; const int* k_idx;
    // fused five-product kernel: Q / dO as 128-row boxes, fp32 dQ accumulator [n_q, heads * 128] (reduce-add target)
    CUtensorMap tma_q128, tma_do128;
    CUtensorMap tma_dqacc;                                 // fp32 [heads, n_q, 128], box {128, 16, 1}, no swizzle
    float* dq_acc;
};

#if B200TTA_ATTN_DEBUG
__device__ long long g_dbg[32];
#define DBG_ON (blockIdx.x == 0 && blockIdx.y == 0)
#define DBG_CLK() clock64()
#define DBG_SET(i, v) do { g_dbg[i] = (v); } while (0)
// developer timeline (scratch/bwd_timeline.py): SM clock at the hand-over points of one CTA of each kernel, TL_N
// consecutive K/V blocks (dq) / query sub-tiles (dkv) starting at TL_FIRST.  Slots per row: compute warp 0-4, issuer 8-11
// (see the TL_MARK sites).
constexpr int TL_CTA = 100, TL_FIRST = 100, TL_N = 24, TL_SLOTS = 16;
__device__ long long g_tl_dq[TL_N * TL_SLOTS], g_tl_dkv[2 * TL_N * TL_SLOTS];
#define TL_MARK(buf, first, n, cond, idx, slot)                                                             \
    do {                                                                                                    \
        if ((cond) && blockIdx.x == TL_CTA && blockIdx.y == 0 && (idx) >= (first) && (idx) < (first) + (n)) \
            buf[((idx) - (first)) * TL_SLOTS + (slot)] = clock64();                                          \
    } while (0)
#define TL_DQ(cond, t, slot) TL_MARK(g_tl_dq, TL_FIRST, TL_N, cond, t, slot)
#define TL_DKV(cond, u, slot) TL_MARK(g_tl_dkv, 2 * TL_FIRST, 2 * TL_N, cond, u, slot)
#define TL_F(cond, t, slot) TL_MARK(g_tl_dkv, TL_FIRST, 2 * TL_N, cond, t, slot)
#else
#define DBG_ON false
#define DBG_CLK() 0ll
#define DBG_SET(i, v) do { } while (0)
#define TL_DQ(cond, t, slot) do { } while (0)
#define TL_DKV(cond, u, slot) do { } while (0)
#define TL_F(cond, t, slot) do { } while (0)
#endif

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, uint32_t tmem_addr, float mul, bool do_store,
                                               int c_begin, int c_end) {
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
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

// S-like MMA, SS form: D[128 x 64] = A[128 x 128(d)] (resident in smem, K-major) * B[64 x 128(d)]^T (streamed, K-major)
__device__ __forceinline__ void mma_ss_n64(uint32_t d_tmem, uint32_t a_smem, uint32_t b_smem) {
    constexpr uint32_t idesc = umma_idesc_bf16(BT, SUB, 0, 0);
    const uint64_t ad = umma_desc_kmajor(a_smem), bd = umma_desc_kmajor(b_smem);
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            umma_ss_e(d_tmem, umma_desc_advance(ad, c * HALF_BYTES + ks * 32), umma_desc_advance(bd, c * SUBH_BYTES + ks * 32),
                      idesc, (c | ks) ? 1u : 0u);
}
// accumulate MMA: D[128 x 128(d)] += A[128 x 64] (bf16 in TMEM) * B[64 x 128(d)] (streamed tile read MN-major).
// A's 64 contraction columns sit where the compute threads left them: 16-column runs at +0 and +32.
__device__ __forceinline__ void mma_ts_k64(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_smem) {
    constexpr uint32_t idesc = umma_idesc_bf16(BT, D, 0, 1);
    const uint64_t bd = umma_desc_mnmajor(b_smem, SUBH_BYTES);
#pragma unroll
    for (int ks = 0; ks < SUB / 16; ++ks)
        umma_ts_e(d_tmem, a_tmem + (ks >> 1) * 32 + (ks & 1) * 8, umma_desc_advance(bd, ks * 2048), idesc, 1u);
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
// One CTA per 128-row query tile; K/V stream through in 128-row blocks; every MMA is TS-form with N = 128 (tensor
// rate, B operand only from shared memory: 160 KB per block = 1 250 clk against 1 536 clk of tensor work).
// TMEM: S [0,128) | dP [128,256) | dQ [256,384) | Q bf16 [384,448) | dO bf16 [448,512).  Single S and dP buffers:
//   S(t+1) is issued as soon as the compute threads have LOADED S(t) (s_free), so it runs under the exponentials;
//   dS(t) (bf16) overwrites the dP columns it was computed from and dP(t+1) is issued right behind dQ(t) in the issuer's
//   in-order stream.  All 16 compute warps work on the same tile (thread = row x 32-column quarter).
constexpr int DQ_STAGES = 3;
constexpr int DQ_SMEM_BYTES = 2 * DQ_STAGES * TILE_BYTES + 1024 + 512;

// D[128 x 128] = A[128 x 128(d)] (bf16 in TMEM, 64 columns) * B[128 x 128(d)]^T (streamed tile, K-major halves)
__device__ __forceinline__ void mma_ts_n128(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_smem) {
    constexpr uint32_t idesc = umma_idesc_bf16(BT, BT, 0, 0);
    const uint64_t bd = umma_desc_kmajor(b_smem);
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks)
        umma_ts_e(d_tmem, a_tmem + ks * 8, umma_desc_advance(bd, (ks >> 2) * HALF_BYTES + (ks & 3) * 32), idesc, ks ? 1u : 0u);
}
// D[128 x 128(d)] += A[128 x 128(tokens)] (bf16 in TMEM: 16-column runs every 32 columns) * B[128 x 128(d)] (MN-major)
__device__ __forceinline__ void mma_ts_k128(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_smem) {
    constexpr uint32_t idesc = umma_idesc_bf16(BT, D, 0, 1);
    const uint64_t bd = umma_desc_mnmajor(b_smem, HALF_BYTES);
#pragma unroll
    for (int ks = 0; ks < BT / 16; ++ks)
        umma_ts_e(d_tmem, a_tmem + (ks >> 1) * 32 + (ks & 1) * 8, umma_desc_advance(bd, ks * 2048), idesc, 1u);
}

__global__ void __launch_bounds__(NUM_THREADS, 1) dq_kernel(const __grid_constant__ BwdParams p) {
    constexpr int STAGES = DQ_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* k_s = smem;                           // [STAGES][TILE_BYTES]
    uint8_t* v_s = k_s + STAGES * TILE_BYTES;      // [STAGES][TILE_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + STAGES * TILE_BYTES);
    uint64_t* qdo_full = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = k_full + STAGES;
    uint64_t* v_full = k_empty + STAGES;
    uint64_t* v_empty = v_full + STAGES;
    uint64_t* s_full = v_empty + STAGES;     // tcgen05.commit: S(t) complete
    uint64_t* s_free = s_full + 1;           // compute threads: S(t) is in registers
    uint64_t* dp_full = s_free + 1;          // tcgen05.commit: dP(t) complete
    uint64_t* ds_full = dp_full + 1;         // compute threads: dS(t) written over dP
    uint64_t* dq_done = ds_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(dq_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x, head = blockIdx.y;
    int seg = 0;
    while (seg + 1 < p.n_seg && item >= p.seg_item0[seg + 1]) ++seg;
    const int q0 = p.seg_q_begin[seg] + (item - p.seg_item0[seg]) * BT;
    const int q_end = p.seg_q_end[seg];
    const int kv_len = p.seg_kv_len[seg];
    int n_blk = (kv_len + BT - 1) / BT;
    const int* blk_list = nullptr;       // block-sparse: the K/V blocks this (head, query block) attends
    if (p.q_off != nullptr) {
        const int o = p.q_off[head * gridDim.x + item];
        n_blk = p.q_off[head * gridDim.x + item + 1] - o;
        blk_list = p.q_idx + o;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tma_k128); tma_prefetch_desc(&p.tma_v128);
        mbar_init(qdo_full, 2 * GROUP_THREADS / 32);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        }
        mbar_init(s_full, 1); mbar_init(s_free, 2 * GROUP_THREADS / 32);
        mbar_init(dp_full, 1); mbar_init(ds_full, 2 * GROUP_THREADS / 32);
        mbar_init(dq_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_ptr != 0) __trap();   // all 512 columns, one CTA per SM: the allocation starts at 0
    constexpr uint32_t t_s = 0, t_dp = 128, t_dq = 256, t_q = 384, t_do = 448;

    if (warp == 0) {
        if (lane == 0) {
            for (int t = 0; t < n_blk; ++t) {
                const int st = t % STAGES;
                const uint32_t ph = (t / STAGES) & 1;
                mbar_wait(&k_empty[st], ph ^ 1u);
                mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
                const int kb = blk_list ? blk_list[t] : t;
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(k_s + st * TILE_BYTES + c * HALF_BYTES, &p.tma_k128, &k_full[st], head * D + c * 64, kb * BT);
                mbar_wait(&v_empty[st], ph ^ 1u);
                mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(v_s + st * TILE_BYTES + c * HALF_BYTES, &p.tma_v128, &v_full[st], head * D + c * 64, kb * BT);
            }
        }
    } else if (warp == 1) {
        // one issuer: every poll below sits behind >= 512 clk of queued MMAs (whole warp converged, elected lane issues)
        mbar_wait(qdo_full, 0);
        if (n_blk > 0) {
            mbar_wait2(&k_full[0], 0, &v_full[0], 0);
            tc_fence_after();
            mma_ts_n128(t_s, t_q, smem_u32(k_s));
            umma_commit_e(s_full);
            mma_ts_n128(t_dp, t_do, smem_u32(v_s));
            umma_commit_e(dp_full);
            umma_commit_e(&v_empty[0]);
        }
        for (int t = 0; t < n_blk; ++t) {
            const int st = t % STAGES, tn = t + 1, st_n = tn % STAGES;
            const uint32_t ph_n = (tn / STAGES) & 1;
            if (tn < n_blk) {   // S(t+1) as soon as S(t) is in registers
                mbar_wait2(s_free, t & 1, &k_full[st_n], ph_n);
                tc_fence_after();
                TL_DQ(lane == 0, t, 8);      // S(t) released
                mma_ts_n128(t_s, t_q, smem_u32(k_s + st_n * TILE_BYTES));
                umma_commit_e(s_full);
                mbar_wait(&v_full[st_n], ph_n);   // long complete: keeps the poll out of the dQ -> dP path
                TL_DQ(lane == 0, t, 9);      // S(t+1) issued
            }
            mbar_wait(ds_full, t & 1);
            tc_fence_after();
            TL_DQ(lane == 0, t, 10);         // dS(t) seen
            mma_ts_k128(t_dq, t_dp, smem_u32(k_s + st * TILE_BYTES));
            umma_commit_e(&k_empty[st]);
            if (tn < n_blk) {   // dP(t+1) overwrites dS(t): ordered behind the dQ MMAs just issued
                mma_ts_n128(t_dp, t_do, smem_u32(v_s + st_n * TILE_BYTES));
                umma_commit_e(dp_full);
                umma_commit_e(&v_empty[st_n]);
            }
            TL_DQ(lane == 0, t, 11);         // dQ(t) and dP(t+1) issued
        }
        umma_commit_e(dq_done);
    } else if (warp >= 3) {
        const int c4 = (warp - 3) >> 2;              // 32-column quarter of the 128-wide block
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int q_row = q0 + row;
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const bool row_ok = q_row < q_end;
        {   // Q and dO rows -> TMEM as packed bf16 pairs (word w of a row = elements 2w, 2w+1 = one TMEM column)
            uint32_t qw[16], dw[16];
            const uint4* qsrc = reinterpret_cast<const uint4*>(p.Q + (long long)q_row * p.ldq + head * D + c4 * 32);
            const uint4* dsrc = reinterpret_cast<const uint4*>(p.dO + (long long)q_row * p.lddo + head * D + c4 * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 a = row_ok ? __ldg(qsrc + i) : make_uint4(0, 0, 0, 0);
                const uint4 b = row_ok ? __ldg(dsrc + i) : make_uint4(0, 0, 0, 0);
                qw[4 * i] = a.x; qw[4 * i + 1] = a.y; qw[4 * i + 2] = a.z; qw[4 * i + 3] = a.w;
                dw[4 * i] = b.x; dw[4 * i + 1] = b.y; dw[4 * i + 2] = b.z; dw[4 * i + 3] = b.w;
            }
            tmem_st_32x32b_x16(t_q + lane_addr + c4 * 16, qw);
            tmem_st_32x32b_x16(t_do + lane_addr + c4 * 16, dw);
            uint32_t zero[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) zero[i] = 0u;
            tmem_st_32x32b_x32(t_dq + lane_addr + c4 * 32, zero);   // every dQ MMA accumulates
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(qdo_full);
        }
        const float lse2 = row_ok ? p.LSE[(long long)head * p.n_q + q_row] * LOG2E : 0.f;
        const float dlt = row_ok ? p.delta[(long long)head * p.n_q + q_row] : 0.f;
        const uint32_t a_s = t_s + lane_addr + c4 * 32, a_dp = t_dp + lane_addr + c4 * 32;
        // invalid rows (beyond the segment) get lse = +inf so that exp2(s - lse) = 0 without a per-element select
        const float neg_lse = row_ok ? -lse2 : -INFINITY;
        for (int t = 0; t < n_blk; ++t) {
            mbar_wait(s_full, t & 1);
            tc_fence_after();
            TL_DQ(warp == 3 && lane == 0, t, 0);     // S(t) seen
            const int valid = blk_list ? 32 : kv_len - t * BT - c4 * 32;   // block-sparse lists only hold whole blocks
            uint32_t sv[32], dp[32], pk[16];
            tmem_ld_32x32b_x32(a_s, sv);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive_warp(s_free);         // S(t + 1) may overwrite the buffer from here on
            TL_DQ(warp == 3 && lane == 0, t, 1);     // S(t) loaded and released
            if (valid >= 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    sv[i] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(sv[i]), p.scale_log2, neg_lse)));
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float pv = fast_exp2(fmaf(__uint_as_float(sv[i]), p.scale_log2, neg_lse));
                    sv[i] = __float_as_uint(i < valid ? pv : 0.f);
                }
            }
            TL_DQ(warp == 3 && lane == 0, t, 2);     // exponentials done
            mbar_wait(dp_full, t & 1);
            tc_fence_after();
            TL_DQ(warp == 3 && lane == 0, t, 3);     // dP(t) seen
            tmem_ld_32x32b_x32(a_dp, dp);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2)
                pk[i >> 1] = pack_bf16x2(__uint_as_float(sv[i]) * (__uint_as_float(dp[i]) - dlt),
                                         __uint_as_float(sv[i + 1]) * (__uint_as_float(dp[i + 1]) - dlt));
            tmem_st_32x32b_x16(a_dp, pk);   // dS over the first half of the dP columns this thread just loaded
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(ds_full);
            TL_DQ(warp == 3 && lane == 0, t, 4);     // dS(t) stored and signalled
        }
        mbar_wait(dq_done, 0);
        tc_fence_after();
        store_row_bf16(p.dQ + (long long)q_row * p.lddq + head * D, t_dq + lane_addr, p.scale, row_ok, c4, c4 + 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(0); }
}

// ------------------------------------------------------------------------------------ dK, dV
// TMEM: buffer b: S^T at [128 b, 128 b + 64), dP^T at [128 b + 64, 128 b + 128); dV at [256, 384); dK at [384, 512).
__global__ void __launch_bounds__(NUM_THREADS, 1) dkv_kernel(const __grid_constant__ BwdParams p) {
    constexpr int STAGES = DKV_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* k_s = smem;
    uint8_t* v_s = k_s + TILE_BYTES;
    uint8_t* q_s = v_s + TILE_BYTES;               // [STAGES][SUBT_BYTES]
    uint8_t* do_s = q_s + STAGES * SUBT_BYTES;     // [STAGES][SUBT_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(do_s + STAGES * SUBT_BYTES);
    uint64_t* kv_full = bars;
    uint64_t* q_full = bars + 1;
    uint64_t* q_empty = q_full + STAGES;
    uint64_t* do_full = q_empty + STAGES;
    uint64_t* do_empty = do_full + STAGES;
    uint64_t* stat_full = do_empty + STAGES;
    uint64_t* s_full = stat_full + STAGES;    // [2] tcgen05.commit: S^T_b complete
    uint64_t* dp_full = s_full + 2;           // [2] tcgen05.commit: dP^T_b complete
    uint64_t* p_full = dp_full + 2;           // [2] compute group: P^T written over S^T_b
    uint64_t* ds_full = p_full + 2;           // [2] compute group: dS^T written over dP^T_b
    uint64_t* dkv_done = ds_full + 2;
    uint64_t* acc_zero = dkv_done + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_zero + 1);
    float* stat_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);  // [STAGES][2][SUB]: LSE (log2), delta

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kv0 = blockIdx.x * BT, head = blockIdx.y;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tma_q64); tma_prefetch_desc(&p.tma_do64); tma_prefetch_desc(&p.tma_k128); tma_prefetch_desc(&p.tma_v128);
        mbar_init(kv_full, 1);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&do_full[i], 1); mbar_init(&do_empty[i], 1);
            mbar_init(&stat_full[i], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_full[b], 1); mbar_init(&dp_full[b], 1);
            mbar_init(&p_full[b], GROUP_THREADS / 32); mbar_init(&ds_full[b], GROUP_THREADS / 32);
        }
        mbar_init(dkv_done, 2);
        mbar_init(acc_zero, 2 * GROUP_THREADS / 32);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_ptr != 0) __trap();
    constexpr uint32_t tmem_base = 0;
    constexpr uint32_t t_dv = tmem_base + 256, t_dk = tmem_base + 384;

    // every role walks the same list of 64-row query sub-tiles: for each segment that can see this K/V block
    const int* blk_list = nullptr;       // block-sparse: the query blocks that attend this (head, key block)
    int n_list = 0;
    if (p.k_off != nullptr) {
        const int o = p.k_off[head * gridDim.x + blockIdx.x];
        n_list = p.k_off[head * gridDim.x + blockIdx.x + 1] - o;
        blk_list = p.k_idx + o;
    }
    auto for_each_sub = [&](auto&& fn) {
        int u = 0;
        if (blk_list != nullptr) {
            for (int e = 0; e < n_list; ++e) {
                const int q0 = blk_list[e] * BT;
                fn(u++, 0, q0);
                fn(u++, 0, q0 + SUB);
            }
            return u;
        }
        for (int s = 0; s < p.n_seg; ++s) {
            if (kv0 >= p.seg_kv_len[s]) continue;
            for (int q0 = p.seg_q_begin[s]; q0 < p.seg_q_end[s]; q0 += SUB) fn(u++, s, q0);
        }
        return u;
    };
    int n_sub = 0;
    if (blk_list != nullptr) n_sub = 2 * n_list;
    else
        for (int s = 0; s < p.n_seg; ++s)
            if (kv0 < p.seg_kv_len[s]) n_sub += (p.seg_q_end[s] - p.seg_q_begin[s] + SUB - 1) / SUB;

    if (warp == 0) {
        // whole warp: lane 0 drives TMA, all lanes stage the sub-tile's 64 LSE / delta values (two columns per lane)
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * TILE_BYTES);
            for (int c = 0; c < 2; ++c) {
                tma_load_2d(k_s + c * HALF_BYTES, &p.tma_k128, kv_full, head * D + c * 64, kv0);
                tma_load_2d(v_s + c * HALF_BYTES, &p.tma_v128, kv_full, head * D + c * 64, kv0);
            }
        }
        __syncwarp();
        for_each_sub([&](int u, int s, int q0) {
            const int st = u % STAGES;
            const uint32_t ph = (u / STAGES) & 1;
            // issue the global loads first: their latency overlaps the wait for the slot
            const int qc = q0 + 2 * lane;
            const int q_lim = p.seg_q_end[s];
            const float* lse_g = p.LSE + (long long)head * p.n_q;
            const float* dl_g = p.delta + (long long)head * p.n_q;
            float2 l2, dl;   // exp2(-inf) = 0 masks the columns beyond the segment
            l2.x = qc < q_lim ? __ldg(lse_g + qc) * LOG2E : INFINITY;
            l2.y = qc + 1 < q_lim ? __ldg(lse_g + qc + 1) * LOG2E : INFINITY;
            dl.x = qc < q_lim ? __ldg(dl_g + qc) : 0.f;
            dl.y = qc + 1 < q_lim ? __ldg(dl_g + qc + 1) : 0.f;
            // do_empty[st] / q_empty[st] complete after the MMAs of sub-tile u - STAGES, which were issued after every
            // compute thread arrived on ds_full, i.e. after the last read of this slot's statistics
            mbar_wait(&do_empty[st], ph ^ 1u);
            mbar_wait(&q_empty[st], ph ^ 1u);
            if (lane == 0) {
                mbar_arrive_expect_tx(&q_full[st], SUBT_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(q_s + st * SUBT_BYTES + c * SUBH_BYTES, &p.tma_q64, &q_full[st], head * D + c * 64, q0);
                mbar_arrive_expect_tx(&do_full[st], SUBT_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(do_s + st * SUBT_BYTES + c * SUBH_BYTES, &p.tma_do64, &do_full[st], head * D + c * 64, q0);
            }
            float* slot = stat_s + st * 2 * SUB;
            *reinterpret_cast<float2*>(slot + 2 * lane) = l2;
            *reinterpret_cast<float2*>(slot + SUB + 2 * lane) = dl;
            __syncwarp();
            if (lane == 0) mbar_arrive(&stat_full[st]);
        });
    } else if (warp <= 2) {
        // issuer w owns buffer w and the sub-tiles u = w, w+2, ...; whole warp converged, the elected lane issues
        const int w = warp - 1;
        mbar_wait(kv_full, 0);
        mbar_wait(acc_zero, 0);
        tc_fence_after();
        const uint32_t ka = smem_u32(k_s), va = smem_u32(v_s);
        // Per own sub-tile u: dV(u) as soon as P^T(u) is there, then S^T(u+2) (re-uses the S^T buffer: ordered behind
        // dV(u)), then dK(u) once dS^T(u) is there, then dP^T(u+2).  S^T and dP^T are signalled separately so the
        // exponentials of a sub-tile overlap its own dP^T MMAs.
        long long w_ds = 0;
        [[maybe_unused]] const long long w_tot0 = DBG_CLK();
        if (w < n_sub) {
            mbar_wait2(&q_full[w], 0, &do_full[w], 0);
            tc_fence_after();
            mma_ss_n64(tmem_base + w * 128, ka, smem_u32(q_s + w * SUBT_BYTES));
            umma_commit_e(&s_full[w]);
            mma_ss_n64(tmem_base + w * 128 + 64, va, smem_u32(do_s + w * SUBT_BYTES));
            umma_commit_e(&dp_full[w]);
        }
        for (int u = w; u < n_sub; u += 2) {
            const int st = u % STAGES, un = u + 2, st_n = un % STAGES;
            const uint32_t j = (uint32_t)(u >> 1), ph_n = (un / STAGES) & 1;
            long long c0 = DBG_CLK();
            if (un < n_sub) mbar_wait2(&p_full[w], j & 1, &q_full[st_n], ph_n);
            else mbar_wait(&p_full[w], j & 1);
            w_ds += DBG_CLK() - c0;
            tc_fence_after();
            TL_DKV(lane == 0, u, 8);         // P^T(u) seen
            mma_ts_k64(t_dv, tmem_base + w * 128, smem_u32(do_s + st * SUBT_BYTES));
            umma_commit_e(&do_empty[st]);
            if (un < n_sub) {
                mma_ss_n64(tmem_base + w * 128, ka, smem_u32(q_s + st_n * SUBT_BYTES));
                umma_commit_e(&s_full[w]);
            }
            TL_DKV(lane == 0, u, 9);         // dV(u) and S^T(u+2) issued
            c0 = DBG_CLK();
            if (un < n_sub) mbar_wait2(&ds_full[w], j & 1, &do_full[st_n], ph_n);
            else mbar_wait(&ds_full[w], j & 1);
            w_ds += DBG_CLK() - c0;
            tc_fence_after();
            TL_DKV(lane == 0, u, 10);        // dS^T(u) seen
            mma_ts_k64(t_dk, tmem_base + w * 128 + 64, smem_u32(q_s + st * SUBT_BYTES));
            umma_commit_e(&q_empty[st]);
            if (un < n_sub) {
                mma_ss_n64(tmem_base + w * 128 + 64, va, smem_u32(do_s + st_n * SUBT_BYTES));
                umma_commit_e(&dp_full[w]);
            }
            TL_DKV(lane == 0, u, 11);        // dK(u) and dP^T(u+2) issued
        }
        umma_commit_e(dkv_done);
        if (DBG_ON && warp == 1 && lane == 0) { DBG_SET(16, DBG_CLK() - w_tot0); DBG_SET(17, w_ds); DBG_SET(18, (n_sub + 1) / 2); }
    } else {
        const int g = (warp - 3) >> 3;
        const int half = ((warp - 3) >> 2) & 1;   // query columns [32 half, 32 half + 32) of the 64-wide sub-tile
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;  // K/V row inside the block
        const int kv_row = kv0 + row;
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const uint32_t t_s = tmem_base + g * 128 + lane_addr + half * 32, t_dp = t_s + 64;
        {   // every dV / dK MMA accumulates: zero this thread's quarter of both accumulator rows
            uint32_t zero[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) zero[i] = 0u;
            tmem_st_32x32b_x32(t_dv + lane_addr + (g * 2 + half) * 32, zero);
            tmem_st_32x32b_x32(t_dk + lane_addr + (g * 2 + half) * 32, zero);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(acc_zero);
        }
        long long w_sdp = 0, w_pre = 0, w_math = 0, w_st = 0;
        [[maybe_unused]] const long long c_all0 = DBG_CLK();
        for_each_sub([&](int u, int s, int) {
            if ((u & 1) != g) return;
            const int st = u % STAGES;
            long long c0 = DBG_CLK();
            // rows beyond the segment's kv_len contribute nothing.  Block-uniform: only the K/V block that straddles the
            // end of a segment takes the masked variant; folded into one loop ptxas predicates an extra add per element.
            const bool all_rows = kv0 + BT <= p.seg_kv_len[s];
            const float row_bias = kv_row < p.seg_kv_len[s] ? 0.f : -INFINITY;
            const uint32_t lse_a = smem_u32(stat_s + st * 2 * SUB + half * 32), dl_a = lse_a + SUB * 4;
            mbar_wait(&stat_full[st], (u / STAGES) & 1);
            long long c1 = DBG_CLK(); w_pre += c1 - c0;
            TL_DKV(half == 0 && quarter == 3 && lane == 0, u, 0);    // statistics of sub-tile u staged
            mbar_wait(&s_full[g], (u >> 1) & 1);
            tc_fence_after();
            TL_DKV(half == 0 && quarter == 3 && lane == 0, u, 1);    // S^T(u) seen
            long long c2 = DBG_CLK(); w_sdp += c2 - c1;
            // two passes of 16 columns keep the live set small; a pass stores its bf16 P^T over S^T columns this thread
            // has already consumed.  The 32 probabilities stay in registers for dS^T.
            float pf[32];
            auto exp_pass = [&](auto masked_tag) {
                constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
                for (int hp = 0; hp < 2; ++hp) {
                    uint32_t sv[16], pk[8];
                    tmem_ld_32x32b_x16(t_s + hp * 16, sv);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        // per-column statistics are warp-uniform broadcast reads: 16-byte loads keep the LSU off the
                        // shared-memory banks the MMA operand fetches need
#if B200TTA_EXPERIMENT_NO_STATS_LDS
                        const float4 l4 = make_float4(10.f, 10.f, 10.f, 10.f);   // timing experiment only: WRONG results
#else
                        const float4 l4 = lds_f32x4(lse_a + (hp * 16 + i) * 4);
#endif
                        float a0 = fmaf(__uint_as_float(sv[i]), p.scale_log2, -l4.x);
                        float a1 = fmaf(__uint_as_float(sv[i + 1]), p.scale_log2, -l4.y);
                        float a2 = fmaf(__uint_as_float(sv[i + 2]), p.scale_log2, -l4.z);
                        float a3 = fmaf(__uint_as_float(sv[i + 3]), p.scale_log2, -l4.w);
                        if (MASKED) { a0 += row_bias; a1 += row_bias; a2 += row_bias; a3 += row_bias; }
                        const float p0 = fast_exp2(a0), p1 = fast_exp2(a1), p2 = fast_exp2(a2), p3 = fast_exp2(a3);
                        pf[hp * 16 + i] = p0; pf[hp * 16 + i + 1] = p1; pf[hp * 16 + i + 2] = p2; pf[hp * 16 + i + 3] = p3;
                        pk[i >> 1] = pack_bf16x2(p0, p1);
                        pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
                    }
                    tmem_st_32x32b_x8(t_s + hp * 8, pk);     // P^T (bf16)
                }
            };
            if (all_rows) exp_pass(std::false_type{}); else exp_pass(std::true_type{});
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(&p_full[g]);                     // dV(u) can go while dS^T is still being computed
            long long c3 = DBG_CLK(); w_math += c3 - c2;
            TL_DKV(half == 0 && quarter == 3 && lane == 0, u, 2);    // P^T(u) stored and signalled
            mbar_wait(&dp_full[g], (u >> 1) & 1);
            tc_fence_after();
            TL_DKV(half == 0 && quarter == 3 && lane == 0, u, 3);    // dP^T(u) seen
#pragma unroll
            for (int hp = 0; hp < 2; ++hp) {
                uint32_t dp[16], dk[8];
                tmem_ld_32x32b_x16(t_dp + hp * 16, dp);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
#if B200TTA_EXPERIMENT_NO_STATS_LDS
                    const float4 d4 = make_float4(0.01f, 0.01f, 0.01f, 0.01f);
#else
                    const float4 d4 = lds_f32x4(dl_a + (hp * 16 + i) * 4);
#endif
                    dk[i >> 1] = pack_bf16x2(pf[hp * 16 + i] * (__uint_as_float(dp[i]) - d4.x),
                                             pf[hp * 16 + i + 1] * (__uint_as_float(dp[i + 1]) - d4.y));
                    dk[(i >> 1) + 1] = pack_bf16x2(pf[hp * 16 + i + 2] * (__uint_as_float(dp[i + 2]) - d4.z),
                                                   pf[hp * 16 + i + 3] * (__uint_as_float(dp[i + 3]) - d4.w));
                }
                tmem_st_32x32b_x8(t_dp + hp * 8, dk);    // dS^T (bf16) over dP^T columns this thread already loaded
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(&ds_full[g]);
            TL_DKV(half == 0 && quarter == 3 && lane == 0, u, 4);    // dS^T(u) stored and signalled
            w_st += DBG_CLK() - c3;
        });
        if (DBG_ON && warp == 3 && lane == 0) {
            DBG_SET(20, DBG_CLK() - c_all0); DBG_SET(21, w_sdp); DBG_SET(22, w_pre); DBG_SET(23, w_math); DBG_SET(24, w_st);
        }
        mbar_wait(dkv_done, 0);
        tc_fence_after();
        const bool ok = kv_row < p.n_kv;
        const int oc = g * 2 + half;
        if (n_sub > 0) {
            store_row_bf16(p.dV + (long long)kv_row * p.lddv + head * D, t_dv + lane_addr, 1.0f, ok, oc, oc + 1);
            store_row_bf16(p.dK + (long long)kv_row * p.lddk + head * D, t_dk + lane_addr, p.scale, ok, oc, oc + 1);
        } else if (ok) {  // a K/V block no query sees: gradients are zero
            for (int c = oc * 4; c < oc * 4 + 4; ++c) {
                *reinterpret_cast<uint4*>(p.dV + (long long)kv_row * p.lddv + head * D + c * 8) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(p.dK + (long long)kv_row * p.lddk + head * D + c * 8) = make_uint4(0, 0, 0, 0);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}


// ------------------------------------------------------------------------------------ fused dK, dV, dQ: five products
// One CTA per (128-row K/V block, head), K and V resident in shared memory, 128-row Q / dO tiles stream through.  Per
// (query tile, key block) pair exactly the five necessary products run, all with N = 128 (tensor rate; the 64-wide SS
// forms of the split dkv kernel cost 48 instead of 32 clk per instruction):
//     S^T  = K Q^T     SS   -> TMEM [0,128)
//     dP^T = V dO^T    SS   -> TMEM [128,256)
//     dV  += P^T dO    TS   A = P^T (bf16) written over the S^T columns each thread consumed itself
//     dQ^T = K^T dS^T  SS   A = K read MN-major, B = dS staged in shared memory as [q][key]; D -> TMEM [128,256) (over dP^T)
//     dK  += dS^T Q    SS   A = the same dS tile read MN-major, B = Q read MN-major
// TMEM is full with S^T, dP^T, dV, dK (4 x 128 columns), so the dQ^T partial product re-uses the dP^T columns once every
// compute thread holds its dP^T values in registers (dS leaves through shared memory only).
// dQ^T (lane = head-dim index, column = query row) goes TMEM -> registers -> a 32 KB shared-memory staging buffer (half a
// tile at a time, rows = queries) -> TMA fp32 reduce-add into the accumulator [n_q, heads * 128] that a small kernel
// converts to bf16 afterwards.  The reduce-adds are asynchronous (direct red.global from the compute threads blocked them
// for ~2 900 clk per tile: the chip-wide atomic rate is ~6 TB/s = 64 KB per ~3 000 clk per SM, scratch/red_bw.cu) and
// that rate is the floor of this design: 151 GB of partial tiles at the headline shape = 25 ms.
// Software pipeline of the 16 compute warps (thread = TMEM lane x 32-column quarter), tile t:
//     dS(t) | stage dQ(t-1) second half | first half of the exponentials of tile t+1 | dQ^T(t) -> registers, stage first
//     half | second half of the exponentials of tile t+1
// and of the single MMA issuer:  dV(t) | S^T(t+1) | dQ^T(t) | dK(t) | dP^T(t+1)   (S^T(t+1) runs under dS(t), dP^T(t+1)
// under the second half of the exponentials).  Shared memory: K, V, 2 x Q, 1 x dO (dO(t) is dead after dV(t)), dS, stage.
constexpr int F_STG_BYTES = 64 * D * 4;
constexpr int F_SMEM_BYTES = 2 * TILE_BYTES + 2 * TILE_BYTES + TILE_BYTES + TILE_BYTES + F_STG_BYTES + 512 + 2 * 2 * BT * 4;
static_assert(F_SMEM_BYTES <= 232448, "fused backward: shared memory budget");

// D[128 x 128] = A[128 x 128(d)] (smem, K-major) * B[128 x 128(d)]^T (smem, K-major)
__device__ __forceinline__ void mma_ss_n128(uint32_t d_tmem, uint32_t a_smem, uint32_t b_smem) {
    constexpr uint32_t idesc = umma_idesc_bf16(BT, BT, 0, 0);
    const uint64_t ad = umma_desc_kmajor(a_smem), bd = umma_desc_kmajor(b_smem);
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
        const uint32_t off = (ks >> 2) * HALF_BYTES + (ks & 3) * 32;
        umma_ss_e(d_tmem, umma_desc_advance(ad, off), umma_desc_advance(bd, off), idesc, ks ? 1u : 0u);
    }
}
// D[128(d) x 128(q)] = A[128(d) x 128(keys)] (K tile read MN-major) * B[128(q) x 128(keys)]^T (dS tile, K-major)
__device__ __forceinline__ void mma_ss_dq(uint32_t d_tmem, uint32_t k_smem, uint32_t ds_smem) {
    constexpr uint32_t idesc = umma_idesc_bf16(D, BT, 1, 0);
    const uint64_t ad = umma_desc_mnmajor(k_smem, HALF_BYTES), bd = umma_desc_kmajor(ds_smem);
#pragma unroll
    for (int ks = 0; ks < BT / 16; ++ks)
        umma_ss_e(d_tmem, umma_desc_advance(ad, ks * 2048), umma_desc_advance(bd, (ks >> 2) * HALF_BYTES + (ks & 3) * 32),
                  idesc, ks ? 1u : 0u);
}
// D[128(keys) x 128(d)] += A[128(keys) x 128(q)] (dS tile [q][key] read MN-major) * B[128(q) x 128(d)] (Q tile, MN-major)
__device__ __forceinline__ void mma_ss_dk(uint32_t d_tmem, uint32_t ds_smem, uint32_t q_smem) {
    constexpr uint32_t idesc = umma_idesc_bf16(BT, D, 1, 1);
    const uint64_t ad = umma_desc_mnmajor(ds_smem, HALF_BYTES), bd = umma_desc_mnmajor(q_smem, HALF_BYTES);
#pragma unroll
    for (int ks = 0; ks < BT / 16; ++ks)
        umma_ss_e(d_tmem, umma_desc_advance(ad, ks * 2048), umma_desc_advance(bd, ks * 2048), idesc, 1u);
}

__global__ void __launch_bounds__(NUM_THREADS, 1) dkvq_kernel(const __grid_constant__ BwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_f[];
    uint8_t* k_s = smem_f;
    uint8_t* v_s = k_s + TILE_BYTES;
    uint8_t* q_s = v_s + TILE_BYTES;               // [2][TILE_BYTES]
    uint8_t* do_s = q_s + 2 * TILE_BYTES;          // single stage
    uint8_t* ds_s = do_s + TILE_BYTES;             // dS as [q][key]: two [128 x 64] swizzled halves (keys 0-63 / 64-127)
    uint8_t* stg_s = ds_s + TILE_BYTES;            // [64 query rows][128] fp32: half of a dQ partial tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg_s + F_STG_BYTES);
    uint64_t* kv_full = bars;
    uint64_t* q_full = bars + 1;              // [2]
    uint64_t* q_empty = q_full + 2;           // [2] tcgen05.commit behind dK(t)
    uint64_t* stat_full = q_empty + 2;        // [2]
    uint64_t* do_full = stat_full + 2;
    uint64_t* do_empty = do_full + 1;         // tcgen05.commit behind dV(t)
    uint64_t* s_full = do_empty + 1;          // tcgen05.commit: S^T(t) complete
    uint64_t* dp_full = s_full + 1;           // tcgen05.commit: dP^T(t) complete
    uint64_t* p_full = dp_full + 1;           // compute: P^T(t) stored
    uint64_t* ds_full = p_full + 1;           // compute: dS(t) in shared memory
    uint64_t* ds_free = ds_full + 1;          // tcgen05.commit behind dK(t): the dS tile may be overwritten
    uint64_t* dq_full = ds_free + 1;          // tcgen05.commit: dQ^T(t) complete
    uint64_t* dq_free = dq_full + 1;          // compute: dQ^T(t) is in registers
    uint64_t* dkv_done = dq_free + 1;
    uint64_t* acc_zero = dkv_done + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_zero + 1);
    float* stat_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);  // [2][2][BT]: -LSE (log2), -delta

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kv0 = blockIdx.x * BT, head = blockIdx.y;
    if ((smem_u32(smem_f) & 1023u) != 0) __trap();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tma_q128); tma_prefetch_desc(&p.tma_do128); tma_prefetch_desc(&p.tma_k128); tma_prefetch_desc(&p.tma_v128);
        tma_prefetch_desc(&p.tma_dqacc);
        mbar_init(kv_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&stat_full[i], 1); }
        mbar_init(do_full, 1); mbar_init(do_empty, 1);
        mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(dq_full, 1); mbar_init(ds_free, 1);
        mbar_init(p_full, 2 * GROUP_THREADS / 32); mbar_init(ds_full, 2 * GROUP_THREADS / 32); mbar_init(dq_free, 2 * GROUP_THREADS / 32);
        mbar_init(dkv_done, 1);
        mbar_init(acc_zero, 2 * GROUP_THREADS / 32);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_ptr != 0) __trap();
    constexpr uint32_t t_s = 0, t_dp = 128, t_dv = 256, t_dk = 384;

    // Every role walks the same list of 128-row query tiles: for each segment that can see this K/V block.  The walk of
    // a segment starts at a tile that depends on the key block and wraps around: the CTAs resident at the same time are
    // consecutive key blocks of one head, and in lock-step they would all reduce-add into the SAME dQ tile at once
    // (same-address atomics serialise in L2: scratch/red_bw.cu "same" vs "staggered").
    int n_tiles = 0;
    for (int s = 0; s < p.n_seg; ++s)
        if (kv0 < p.seg_kv_len[s]) n_tiles += (p.seg_q_end[s] - p.seg_q_begin[s] + BT - 1) / BT;
    auto tile_at = [&](int t, int& s_out, int& q0_out) {
        for (int s = 0; s < p.n_seg; ++s) {
            if (kv0 >= p.seg_kv_len[s]) continue;
            const int n = (p.seg_q_end[s] - p.seg_q_begin[s] + BT - 1) / BT;
            if (t < n) {
                int i = (int)(blockIdx.x % (unsigned)n) + t;
                if (i >= n) i -= n;
                s_out = s; q0_out = p.seg_q_begin[s] + i * BT;
                return;
            }
            t -= n;
        }
        s_out = 0; q0_out = 0;
    };

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * TILE_BYTES);
            for (int c = 0; c < 2; ++c) {
                tma_load_2d(k_s + c * HALF_BYTES, &p.tma_k128, kv_full, head * D + c * 64, kv0);
                tma_load_2d(v_s + c * HALF_BYTES, &p.tma_v128, kv_full, head * D + c * 64, kv0);
            }
        }
        __syncwarp();
        for (int t = 0; t < n_tiles; ++t) {
            int s, q0;
            tile_at(t, s, q0);
            const int st = t & 1;
            const int qc = q0 + 4 * lane;     // four query columns per lane
            const int q_lim = p.seg_q_end[s];
            const float* lse_g = p.LSE + (long long)head * p.n_q;
            const float* dl_g = p.delta + (long long)head * p.n_q;
            float4 l4, d4;   // negated: the compute threads fold them into packed FMAs / ADDs; exp2(s - inf) = 0 masks columns
            l4.x = qc < q_lim ? -__ldg(lse_g + qc) * LOG2E : -INFINITY;
            l4.y = qc + 1 < q_lim ? -__ldg(lse_g + qc + 1) * LOG2E : -INFINITY;
            l4.z = qc + 2 < q_lim ? -__ldg(lse_g + qc + 2) * LOG2E : -INFINITY;
            l4.w = qc + 3 < q_lim ? -__ldg(lse_g + qc + 3) * LOG2E : -INFINITY;
            d4.x = qc < q_lim ? -__ldg(dl_g + qc) : 0.f;
            d4.y = qc + 1 < q_lim ? -__ldg(dl_g + qc + 1) : 0.f;
            d4.z = qc + 2 < q_lim ? -__ldg(dl_g + qc + 2) : 0.f;
            d4.w = qc + 3 < q_lim ? -__ldg(dl_g + qc + 3) : 0.f;
            // q_empty[st] completes behind dK(t-2), issued after every compute thread arrived on ds_full(t-2), i.e. after
            // the last read of this slot's statistics
            mbar_wait(&q_empty[st], ((t >> 1) & 1) ^ 1u);
            if (lane == 0) {
                mbar_arrive_expect_tx(&q_full[st], TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(q_s + st * TILE_BYTES + c * HALF_BYTES, &p.tma_q128, &q_full[st], head * D + c * 64, q0);
            }
            float* slot = stat_s + st * 2 * BT;
            *reinterpret_cast<float4*>(slot + 4 * lane) = l4;
            *reinterpret_cast<float4*>(slot + BT + 4 * lane) = d4;
            __syncwarp();
            if (lane == 0) mbar_arrive(&stat_full[st]);
            mbar_wait(do_empty, (t & 1) ^ 1u);          // dV(t-1) is done with the single dO buffer
            if (lane == 0) {
                mbar_arrive_expect_tx(do_full, TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(do_s + c * HALF_BYTES, &p.tma_do128, do_full, head * D + c * 64, q0);
            }
        }
    } else if (warp == 1) {
        // single issuer, whole warp converged, the elected lane issues; all MMAs are ordered in this one stream
        mbar_wait(kv_full, 0);
        mbar_wait(acc_zero, 0);
        tc_fence_after();
        const uint32_t ka = smem_u32(k_s), va = smem_u32(v_s), dsa = smem_u32(ds_s), doa = smem_u32(do_s);
        if (n_tiles > 0) {
            mbar_wait2(&q_full[0], 0, do_full, 0);
            tc_fence_after();
            mma_ss_n128(t_s, ka, smem_u32(q_s));
            umma_commit_e(s_full);
            mma_ss_n128(t_dp, va, doa);
            umma_commit_e(dp_full);
        }
        for (int t = 0; t < n_tiles; ++t) {
            const uint32_t par = (uint32_t)t & 1u;
            const int tn = t + 1;
            const uint32_t qa = smem_u32(q_s + (t & 1) * TILE_BYTES);
            mbar_wait(p_full, par);
            tc_fence_after();
            TL_F(lane == 0, t, 10);          // P^T(t) seen
            mma_ts_k128(t_dv, t_s, doa);
            umma_commit_e(do_empty);
            if (tn < n_tiles) {              // S^T(t+1) overwrites P^T(t): ordered behind dV(t) in this stream
                mbar_wait(&q_full[tn & 1], (tn >> 1) & 1);
                tc_fence_after();
                mma_ss_n128(t_s, ka, smem_u32(q_s + (tn & 1) * TILE_BYTES));
                umma_commit_e(s_full);
            }
            TL_F(lane == 0, t, 11);          // dV(t), S^T(t+1) issued
            mbar_wait(ds_full, par);
            tc_fence_after();
            TL_F(lane == 0, t, 12);          // dS(t) seen
            mma_ss_dq(t_dp, ka, dsa);
            umma_commit_e(dq_full);
            mma_ss_dk(t_dk, dsa, qa);
            umma_commit_e(&q_empty[t & 1]);
            umma_commit_e(ds_free);
            TL_F(lane == 0, t, 13);          // dQ^T(t), dK(t) issued
            if (tn < n_tiles) {              // dP^T(t+1) overwrites dQ^T(t)
                mbar_wait2(dq_free, par, do_full, (uint32_t)tn & 1u);
                tc_fence_after();
                TL_F(lane == 0, t, 8);       // dQ^T(t) drained, dO(t+1) there
                mma_ss_n128(t_dp, va, doa);
                umma_commit_e(dp_full);
                TL_F(lane == 0, t, 9);       // dP^T(t+1) issued
            }
        }
        umma_commit_e(dkv_done);
    } else if (warp >= 3) {
        const int c4 = (warp - 3) >> 2;              // 32-column quarter (query rows of the tile)
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;         // TMEM lane: key row for S^T / dP^T / dV / dK, head-dim index for dQ^T
        const int kv_row = kv0 + row;
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const bool tl = c4 == 0 && quarter == 3 && lane == 0;
        const bool leader = quarter == 0 && lane == 0;   // per quarter: issues (and waits for) its TMA reduce-adds
        {   // every dV / dK MMA accumulates: zero this thread's quarter of both accumulator rows
            uint32_t zero[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) zero[i] = 0u;
            tmem_st_32x32b_x32(t_dv + lane_addr + c4 * 32, zero);
            tmem_st_32x32b_x32(t_dk + lane_addr + c4 * 32, zero);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(acc_zero);
        }
        // dS in shared memory as [q][key] (K-major B operand of dQ^T, MN-major A operand of dK): key `row` sits in half
        // (row >> 6), 16-byte chunk ((row & 63) >> 3) of the 128-byte row, swizzled with (q & 7).  Neighbouring lanes (keys
        // 2m, 2m+1) exchange one packed register so that each stores a 32-bit word: the even lane the pair of keys of
        // query 2j, the odd lane that of query 2j + 1.
        const uint32_t odd = uint32_t(lane & 1);
        const uint32_t ds_thr = smem_u32(ds_s) + uint32_t(row >> 6) * HALF_BYTES + (uint32_t(c4 * 32) + odd) * 128u +
                                (uint32_t(row & 7) >> 1) * 4u;
        const uint32_t ds_cx = (uint32_t(row & 63) >> 3) ^ odd;
        const uint32_t prmt_sel = odd ? 0x3276u : 0x5410u;
        const uint32_t stg_thr = smem_u32(stg_s) + uint32_t(c4 * 16) * 512u + uint32_t(row) * 4u;
        uint32_t sv[32];          // P(t) (fp32) from the exponentials until dS(t)
        uint32_t dqb[16];         // second half of this thread's dQ^T(t) columns, staged one phase later
        int s_cur = 0, q0_cur = 0, s_nxt = 0, q0_nxt = 0, q0_prev = 0;

        // exponentials of one 16-column half of tile t's S^T quarter -> sv, P^T (bf16) over columns this thread has read
        auto exp_half = [&](int st, int s, int hf) {
            const bool all_rows = kv0 + BT <= p.seg_kv_len[s];
            const float row_bias = kv_row < p.seg_kv_len[s] ? 0.f : -INFINITY;
            const uint32_t lse_a = smem_u32(stat_s + st * 2 * BT + c4 * 32 + hf * 16);
            uint32_t raw[16], pk[8];
            tmem_ld_32x32b_x16(t_s + lane_addr + c4 * 32 + hf * 16, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 l4 = lds_f32x4(lse_a + i * 4);
                float a0, a1, a2, a3;
                ffma2v(a0, a1, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]), p.scale_log2, l4.x, l4.y);
                ffma2v(a2, a3, __uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3]), p.scale_log2, l4.z, l4.w);
                if (!all_rows) { a0 += row_bias; a1 += row_bias; a2 += row_bias; a3 += row_bias; }
                const float p0 = fast_exp2(a0), p1 = fast_exp2(a1), p2 = fast_exp2(a2), p3 = fast_exp2(a3);
                sv[hf * 16 + i] = __float_as_uint(p0); sv[hf * 16 + i + 1] = __float_as_uint(p1);
                sv[hf * 16 + i + 2] = __float_as_uint(p2); sv[hf * 16 + i + 3] = __float_as_uint(p3);
                pk[i >> 1] = pack_bf16x2(p0, p1);
                pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
            }
            tmem_st_32x32b_x8(t_s + lane_addr + c4 * 32 + hf * 8, pk);
        };
        // 16 query rows x this thread's head-dim index of a dQ^T partial tile -> staging buffer -> TMA reduce-add
        // (each 32-column quarter = 4 warps owns one 8 KB box of the buffer and its own reduce-add stream: no hand-shake
        // across quarters, and a quarter only ever waits for its own previous box to have been read)
        auto stage = [&](const uint32_t (&v)[16], int q0, int phase) {
#if B200TTA_DQ_DIRECT_RED
            // variant: reduce-add straight from the registers (no staging traffic in shared memory, but the LSU queue
            // back-pressures the issuing warps when the L2 atomic units are behind)
            float* a = p.dq_acc + ((long long)head * p.n_q + q0 + c4 * 32 + phase * 16) * D + row;
            const int n_ok = p.n_q - (q0 + c4 * 32 + phase * 16);
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < n_ok) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a + i * D), "f"(__uint_as_float(v[i])) : "memory");
            return;
#endif
            if (leader) tma_store_wait_read<0>();
            named_bar_sync(2 + c4, 128);
#pragma unroll
            for (int i = 0; i < 16; ++i) sts_b32(stg_thr + i * 512, v[i]);
            fence_proxy_async();
            named_bar_sync(2 + c4, 128);
            if (leader) {
                tma_reduce_add_3d(&p.tma_dqacc, smem_u32(stg_s) + c4 * 8192, 0, q0 + c4 * 32 + phase * 16, head);
                tma_store_commit();
            }
        };

        if (n_tiles > 0) {
            tile_at(0, s_cur, q0_cur);
            mbar_wait(&stat_full[0], 0);
            mbar_wait(s_full, 0);
            tc_fence_after();
            exp_half(0, s_cur, 0);
            exp_half(0, s_cur, 1);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(p_full);
        }
        for (int t = 0; t < n_tiles; ++t) {
            const uint32_t par = (uint32_t)t & 1u;
            const int st = t & 1, tn = t + 1;
            if (tn < n_tiles) tile_at(tn, s_nxt, q0_nxt);
            // ---- dS(t) = P o (dP - delta) -> shared memory
            mbar_wait(dp_full, par);
            tc_fence_after();
            TL_F(tl, t, 3);                  // dP^T(t) seen
            {
                uint32_t dp[32], pk[16];
                tmem_ld_32x32b_x32(t_dp + lane_addr + c4 * 32, dp);
                tmem_ld_wait();
                TL_F(tl, t, 14);             // dP^T(t) in registers
                const uint32_t dl_a = smem_u32(stat_s + st * 2 * BT + BT + c4 * 32);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 d4 = lds_f32x4(dl_a + i * 4);
                    float e0 = __uint_as_float(dp[i]), e1 = __uint_as_float(dp[i + 1]);
                    float e2 = __uint_as_float(dp[i + 2]), e3 = __uint_as_float(dp[i + 3]);
                    fadd2(e0, e1, d4.x, d4.y);
                    fadd2(e2, e3, d4.z, d4.w);
                    fmul2(e0, e1, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
                    fmul2(e2, e3, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
                    pk[i >> 1] = pack_bf16x2(e0, e1);
                    pk[(i >> 1) + 1] = pack_bf16x2(e2, e3);
                }
                tc_fence_before();
                if (t > 0) mbar_wait(ds_free, par ^ 1u);       // dQ^T(t-1), dK(t-1) have read the dS tile
                TL_F(tl, t, 15);             // dS(t) computed, dS tile free
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t recv = __shfl_xor_sync(0xffffffffu, pk[j], 1);
                    uint32_t word;
                    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(word) : "r"(pk[j]), "r"(recv), "r"(prmt_sel));
                    sts_b32(ds_thr + j * 256 + ((ds_cx ^ uint32_t((2 * j) & 7)) << 4), word);
                }
                fence_proxy_async();
                mbar_arrive_warp(ds_full);
            }
            TL_F(tl, t, 4);                  // dS(t) stored and signalled
            // ---- second half of dQ^T(t-1): its first half's reduce-adds have had a dS phase to read the staging buffer
            if (t > 0) stage(dqb, q0_prev, 1);
            TL_F(tl, t, 5);
            // ---- first half of the exponentials of tile t+1 (S^T(t+1) was issued right behind dV(t))
            if (tn < n_tiles) {
                mbar_wait(&stat_full[tn & 1], (tn >> 1) & 1);
                mbar_wait(s_full, (uint32_t)tn & 1u);
                tc_fence_after();
                TL_F(tl, t, 0);              // S^T(t+1) seen
                exp_half(tn & 1, s_nxt, 0);
            }
            TL_F(tl, t, 1);
            // ---- dQ^T(t): TMEM -> registers, hand the columns back, stage the first half
            mbar_wait(dq_full, par);
            tc_fence_after();
            TL_F(tl, t, 6);                  // dQ^T(t) seen
            {
                uint32_t dqa[16];
                tmem_ld_32x32b_x16(t_dp + lane_addr + c4 * 32, dqa);
                tmem_ld_32x32b_x16(t_dp + lane_addr + c4 * 32 + 16, dqb);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive_warp(dq_free);                      // dP^T(t+1) may overwrite the columns
                stage(dqa, q0_cur, 0);
            }
            TL_F(tl, t, 7);                  // first half staged
            // ---- second half of the exponentials of tile t+1
            if (tn < n_tiles) {
                exp_half(tn & 1, s_nxt, 1);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive_warp(p_full);
            }
            TL_F(tl, t, 2);                  // P^T(t+1) stored and signalled
            q0_prev = q0_cur;
            s_cur = s_nxt; q0_cur = q0_nxt;
        }
        if (n_tiles > 0) stage(dqb, q0_prev, 1);
        if (leader) tma_store_wait<0>();                       // the staging buffer must outlive the reduce-adds
        mbar_wait(dkv_done, 0);
        tc_fence_after();
        const bool ok = kv_row < p.n_kv;
        if (n_tiles > 0) {
            store_row_bf16(p.dV + (long long)kv_row * p.lddv + head * D, t_dv + lane_addr, 1.0f, ok, c4, c4 + 1);
            store_row_bf16(p.dK + (long long)kv_row * p.lddk + head * D, t_dk + lane_addr, p.scale, ok, c4, c4 + 1);
        } else if (ok) {
            for (int c = c4 * 4; c < c4 * 4 + 4; ++c) {
                *reinterpret_cast<uint4*>(p.dV + (long long)kv_row * p.lddv + head * D + c * 8) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(p.dK + (long long)kv_row * p.lddk + head * D + c * 8) = make_uint4(0, 0, 0, 0);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(0); }
}

// dQ (bf16, [n_q, heads, 128] strided) = scale * accumulator (fp32 [heads, n_q, 128]); 8 elements per thread
__global__ void __launch_bounds__(256) dq_convert_kernel(__nv_bfloat16* __restrict__ dq, long long lddq,
                                                         const float* __restrict__ acc, long long n_q, int heads, float scale) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;     // index into [n_q][heads][128]
    if (i >= n_q * heads * D) return;
    const long long row = i / (heads * D);
    const int col = (int)(i - row * heads * D), h = col / D, d = col % D;
    const float* src = acc + ((long long)h * n_q + row) * D + d;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
    uint4 o;
    o.x = pack_bf16x2(a.x * scale, a.y * scale); o.y = pack_bf16x2(a.z * scale, a.w * scale);
    o.z = pack_bf16x2(b.x * scale, b.y * scale); o.w = pack_bf16x2(b.z * scale, b.w * scale);
    *reinterpret_cast<uint4*>(dq + row * lddq + col) = o;
}

}  // namespace
}  // namespace b200

using namespace b200;

#if B200TTA_ATTN_DEBUG
extern "C" int b200tta_debug_read(long long* out32) {
    return cudaMemcpyFromSymbol(out32, g_dbg, sizeof(long long) * 32) == cudaSuccess ? 0 : -3;
}
extern "C" int b200tta_debug_bwd_timeline(long long* dq_out, long long* dkv_out) {   // [24][16] and [48][16] clock values
    if (cudaMemcpyFromSymbol(dq_out, g_tl_dq, sizeof(g_tl_dq)) != cudaSuccess) return -3;
    return cudaMemcpyFromSymbol(dkv_out, g_tl_dkv, sizeof(g_tl_dkv)) == cudaSuccess ? 0 : -3;
}
#endif

static int attn_bwd_impl(void* dQ, int64_t lddq, void* dK, int64_t lddk, void* dV, int64_t lddv, const void* dO,
                         int64_t lddo, const void* O, int64_t ldo, const float* LSE, float* delta, const void* Q,
                         int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, int32_t n_q,
                         int32_t n_kv, int32_t heads, float softmax_scale, const b200tta_attn_seg* segs,
                         int32_t n_seg, const int32_t* q_off, const int32_t* q_idx, const int32_t* k_off,
                         const int32_t* k_idx, b200tta_stream_t stream) {
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
    const uint64_t inner = (uint64_t)heads * D;
    if (int rc = make_tmap_2d_bf16(&p.tma_k128, K, inner, (uint64_t)n_kv, (uint64_t)ldk * 2, 64, BT)) return rc;
    if (int rc = make_tmap_2d_bf16(&p.tma_v128, V, inner, (uint64_t)n_kv, (uint64_t)ldv * 2, 64, BT)) return rc;
    if (int rc = make_tmap_2d_bf16(&p.tma_q64, Q, inner, (uint64_t)n_q, (uint64_t)ldq * 2, 64, SUB)) return rc;
    if (int rc = make_tmap_2d_bf16(&p.tma_do64, dO, inner, (uint64_t)n_q, (uint64_t)lddo * 2, 64, SUB)) return rc;
    p.Q = (const __nv_bfloat16*)Q; p.dO = (const __nv_bfloat16*)dO; p.ldq = ldq; p.lddo = lddo;
    p.dQ = (__nv_bfloat16*)dQ; p.dK = (__nv_bfloat16*)dK; p.dV = (__nv_bfloat16*)dV;
    p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
    p.LSE = LSE; p.delta = delta; p.n_q = n_q; p.n_kv = n_kv; p.heads = heads;
    p.scale = softmax_scale; p.scale_log2 = softmax_scale * LOG2E;
    p.n_seg = n_seg;
    p.q_off = q_off; p.q_idx = q_idx; p.k_off = k_off; p.k_idx = k_idx;
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
        B200_CUDA(cudaFuncSetAttribute(dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM_BYTES));
        B200_CUDA(cudaFuncSetAttribute(dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM_BYTES));
        attr = true;
    }
    {
        const long long warps = (long long)n_q * heads;
        delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(delta, (const __nv_bfloat16*)dO, lddo,
                                                                          (const __nv_bfloat16*)O, ldo, n_q, heads);
        B200_LAUNCHED();
    }
    if (!B200TTA_ATTN_DEBUG || !getenv("B200TTA_DEBUG_NO_DQ")) dq_kernel<<<dim3(items, heads), NUM_THREADS, DQ_SMEM_BYTES, st>>>(p);
    B200_LAUNCHED();
    if (!B200TTA_ATTN_DEBUG || !getenv("B200TTA_DEBUG_NO_DKV")) dkv_kernel<<<dim3((n_kv + BT - 1) / BT, heads), NUM_THREADS, DKV_SMEM_BYTES, st>>>(p);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_attn_bwd_fused(void* dQ, int64_t lddq, void* dK, int64_t lddk, void* dV, int64_t lddv, const void* dO,
                                      int64_t lddo, const void* O, int64_t ldo, const float* LSE, float* delta, const void* Q,
                                      int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, int32_t n_q,
                                      int32_t n_kv, int32_t heads, float softmax_scale, const b200tta_attn_seg* segs,
                                      int32_t n_seg, float* dq_acc, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(dQ && dK && dV && dO && O && LSE && delta && Q && K && V && dq_acc && n_q > 0 && n_kv > 0 && heads > 0,
                 "attn_bwd_fused: null/empty argument");
    B200_REQUIRE(n_seg >= 1 && n_seg <= MAX_SEGS && segs, "attn_bwd_fused: n_seg=%d not in [1,%d]", n_seg, MAX_SEGS);
    B200_REQUIRE(aligned16(dQ) && aligned16(dK) && aligned16(dV) && aligned16(dO) && aligned16(O) && aligned16(Q) &&
                     aligned16(K) && aligned16(V) && aligned16(dq_acc) && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 &&
                     lddo % 8 == 0 && ldo % 8 == 0 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0,
                 "attn_bwd_fused: tensors must be 16-byte aligned [tokens, heads, 128] views");
    cudaStream_t st = (cudaStream_t)stream;
    BwdParams p;
    memset(&p, 0, sizeof(p));
    const uint64_t inner = (uint64_t)heads * D;
    if (int rc = make_tmap_2d_bf16(&p.tma_k128, K, inner, (uint64_t)n_kv, (uint64_t)ldk * 2, 64, BT)) return rc;
    if (int rc = make_tmap_2d_bf16(&p.tma_v128, V, inner, (uint64_t)n_kv, (uint64_t)ldv * 2, 64, BT)) return rc;
    if (int rc = make_tmap_2d_bf16(&p.tma_q128, Q, inner, (uint64_t)n_q, (uint64_t)ldq * 2, 64, BT)) return rc;
    if (int rc = make_tmap_2d_bf16(&p.tma_do128, dO, inner, (uint64_t)n_q, (uint64_t)lddo * 2, 64, BT)) return rc;
    p.Q = (const __nv_bfloat16*)Q; p.dO = (const __nv_bfloat16*)dO; p.ldq = ldq; p.lddo = lddo;
    p.dQ = (__nv_bfloat16*)dQ; p.dK = (__nv_bfloat16*)dK; p.dV = (__nv_bfloat16*)dV;
    p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
    p.LSE = LSE; p.delta = delta; p.n_q = n_q; p.n_kv = n_kv; p.heads = heads;
    p.scale = softmax_scale; p.scale_log2 = softmax_scale * LOG2E;
    p.n_seg = n_seg;
    p.dq_acc = dq_acc;
    if (int rc = make_tmap_3d_f32_plain(&p.tma_dqacc, dq_acc, D, (uint64_t)n_q, (uint64_t)heads, D * sizeof(float),
                                        (uint64_t)n_q * D * sizeof(float), D, 16, 1)) return rc;
    for (int s = 0; s < n_seg; ++s) {
        B200_REQUIRE(segs[s].q_begin >= 0 && segs[s].q_end > segs[s].q_begin && segs[s].q_end <= n_q &&
                         segs[s].kv_len > 0 && segs[s].kv_len <= n_kv,
                     "attn_bwd_fused: bad segment %d", s);
        p.seg_q_begin[s] = segs[s].q_begin; p.seg_q_end[s] = segs[s].q_end; p.seg_kv_len[s] = segs[s].kv_len;
    }
    static bool attr = false;
    if (!attr) {
        B200_CUDA(cudaFuncSetAttribute(dkvq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES));
        attr = true;
    }
    const long long warps = (long long)n_q * heads;
    delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(delta, (const __nv_bfloat16*)dO, lddo,
                                                                      (const __nv_bfloat16*)O, ldo, n_q, heads);
    B200_LAUNCHED();
    B200_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)n_q * heads * D * sizeof(float), st));
    dkvq_kernel<<<dim3((n_kv + BT - 1) / BT, heads), NUM_THREADS, F_SMEM_BYTES, st>>>(p);
    B200_LAUNCHED();
    const long long total = (long long)n_q * heads * D;
    dq_convert_kernel<<<(unsigned)((total / 8 + 255) / 256), 256, 0, st>>>(p.dQ, lddq, dq_acc, n_q, heads, softmax_scale);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_attn_bwd(void* dQ, int64_t lddq, void* dK, int64_t lddk, void* dV, int64_t lddv, const void* dO,
                                int64_t lddo, const void* O, int64_t ldo, const float* LSE, float* delta, const void* Q,
                                int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, int32_t n_q,
                                int32_t n_kv, int32_t heads, float softmax_scale, const b200tta_attn_seg* segs,
                                int32_t n_seg, b200tta_stream_t stream) {
    return attn_bwd_impl(dQ, lddq, dK, lddk, dV, lddv, dO, lddo, O, ldo, LSE, delta, Q, ldq, K, ldk, V, ldv, n_q, n_kv, heads,
                         softmax_scale, segs, n_seg, nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int b200tta_attn_bsa_bwd(void* dQ, int64_t lddq, void* dK, int64_t lddk, void* dV, int64_t lddv, const void* dO,
                                    int64_t lddo, const void* O, int64_t ldo, const float* LSE, float* delta,
                                    const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv,
                                    int32_t n_tok, int32_t heads, float softmax_scale, const int32_t* q_off,
                                    const int32_t* q_idx, const int32_t* k_off, const int32_t* k_idx,
                                    b200tta_stream_t stream) {
    B200_REQUIRE(q_off && q_idx && k_off && k_idx, "attn_bsa_bwd: null block list");
    B200_REQUIRE(n_tok > 0 && n_tok % BT == 0, "attn_bsa_bwd: n_tok=%d must be a multiple of the 128-token block", n_tok);
    const b200tta_attn_seg seg = {0, n_tok, n_tok};
    return attn_bwd_impl(dQ, lddq, dK, lddk, dV, lddv, dO, lddo, O, ldo, LSE, delta, Q, ldq, K, ldk, V, ldv, n_tok, n_tok,
                         heads, softmax_scale, &seg, 1, q_off, q_idx, k_off, k_idx, stream);
}
