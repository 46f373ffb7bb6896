// Flash-style attention forward for sm_100a, head_dim 128, bf16.
//
// One CTA owns TWO 128-row query tiles of one head (256 query rows share every K/V tile that comes through
// shared memory) and loops over 128-row K/V blocks:
//   warp 0      TMA producer: Q0,Q1 once; K_j and V_j through two independent 2-deep rings
//   warp 1      MMA issuer (one lane): S_t = Q_t K_j^T (SS), O_t += P_t V_j (A = P from TMEM, B = V MN-major)
//   warps 2-5   softmax warpgroup of tile 0 (one thread per query row)
//   warps 6-9   softmax warpgroup of tile 1
// TMEM (512 columns): S0 | S1 | O0 | O1, 128 fp32 columns each.  P_t (bf16, 64 columns) is written over the first
// half of S_t by the softmax warpgroup, so the tensor core reads its A operand straight from TMEM while the other
// tile's softmax runs (the two tiles ping-pong).  Online softmax keeps a possibly stale row maximum and rescales
// O_t in TMEM only when the maximum grew by more than 2^8 (the exponentials stay well inside bf16 range).
//
// Segments implement the clean-context / noised split of LongCat's self-attention without a second pass over K/V:
// a CTA's query rows belong to one segment and simply stop at that segment's kv_len.
#include "host_common.h"
#include "ptx.cuh"

namespace b200 {
namespace {

constexpr int D = 128;
constexpr int BM = 128;              // rows per query tile
constexpr int BN = 128;              // K/V rows per block
constexpr int KV_STAGES = 2;
constexpr int TILE_BYTES = 128 * 128 * 2;   // one [128 x 128] bf16 tile = two [128 x 64] swizzled sub-tiles
constexpr int SUB_BYTES = TILE_BYTES / 2;
constexpr int fwd_threads(int tiles) { return 32 * (2 + 4 * tiles); }   // producer, issuer, one softmax warpgroup per tile
constexpr int fwd_smem(int tiles) { return tiles * TILE_BYTES + 2 * KV_STAGES * TILE_BYTES + 1024 + 256; }
constexpr int MAX_SEGS = 4;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

struct FwdParams {
    CUtensorMap tma_q, tma_k, tma_v;
    __nv_bfloat16* O; long long ldo;
    float* LSE;
    int n_q, heads;
    float scale_log2;  // softmax_scale * log2(e)
    int n_seg;
    int seg_q_begin[MAX_SEGS], seg_q_end[MAX_SEGS], seg_kv_len[MAX_SEGS], seg_item0[MAX_SEGS + 1];
    // block-sparse variant (NULL = dense; one query tile per CTA): CSR list of the 128-token key blocks attended by
    // (head, query block), ascending; tokens in block-major order
    const int* q_off; const int* q_idx;
};

#ifndef B200TTA_ATTN_DEBUG
#define B200TTA_ATTN_DEBUG 0
#endif
#if B200TTA_ATTN_DEBUG
// developer timeline (scratch/fwd_timeline.py): SM clock at the hand-over points of one CTA for TL_BLOCKS K/V blocks.
// slots per block: tile t softmax (warp quarter 0) 5t+{0 s_full seen, 1 S loaded, 2 max done, 3 first half of P
// signalled, 4 P complete}; issuer 10+3t+{0 p_half seen, 1 p_full seen, 2 next S issued}
constexpr int TL_J0 = 100, TL_BLOCKS = 24, TL_SLOTS = 16, TL_ITEM = 40;
__device__ long long g_timeline[TL_BLOCKS * TL_SLOTS];
#define TL_MARK(cond, j, slot)                                                                             \
    do {                                                                                                   \
        if ((cond) && blockIdx.x == TL_ITEM && blockIdx.y == 0 && (j) >= TL_J0 && (j) < TL_J0 + TL_BLOCKS) \
            g_timeline[((j) - TL_J0) * TL_SLOTS + (slot)] = clock64();                                      \
    } while (0)
#else
#define TL_MARK(cond, j, slot) do { } while (0)
#endif

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// TILES = query tiles per CTA: 2 for dense attention (256 query rows share every K/V tile), 1 for the block-sparse
// variant (every 128-token query block has its own key-block list).
template <int TILES>
__global__ void __launch_bounds__(fwd_threads(TILES), 1) attn_fwd_kernel(const __grid_constant__ FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_smem = smem;                                   // [TILES][TILE_BYTES]
    uint8_t* k_smem = q_smem + TILES * TILE_BYTES;            // [KV_STAGES][TILE_BYTES]
    uint8_t* v_smem = k_smem + KV_STAGES * TILE_BYTES;        // [KV_STAGES][TILE_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + KV_STAGES * TILE_BYTES);
    uint64_t* q_full = bars;                 // 1
    uint64_t* k_full = bars + 1;             // KV_STAGES
    uint64_t* k_empty = k_full + KV_STAGES;
    uint64_t* v_full = k_empty + KV_STAGES;
    uint64_t* v_empty = v_full + KV_STAGES;
    uint64_t* s_full = v_empty + KV_STAGES;  // TILES
    uint64_t* p_full = s_full + TILES;       // TILES
    uint64_t* o_done = p_full + TILES;       // TILES
    uint64_t* p_half = o_done + TILES;       // TILES: first 64 K/V columns of P_t written
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(p_half + TILES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- work item -> (segment, first query row), head
    const int item = blockIdx.x, head = blockIdx.y;
    int seg = 0;
    while (seg + 1 < p.n_seg && item >= p.seg_item0[seg + 1]) ++seg;
    const int q0 = p.seg_q_begin[seg] + (item - p.seg_item0[seg]) * (BM * TILES);
    const int q_end = p.seg_q_end[seg];
    const int kv_len = p.seg_kv_len[seg];
    int n_blocks = (kv_len + BN - 1) / BN;
    const int* blk_list = nullptr;
    if (p.q_off != nullptr) {   // block-sparse (TILES == 1): item = query block
        const int o = p.q_off[head * gridDim.x + item];
        n_blocks = p.q_off[head * gridDim.x + item + 1] - o;
        blk_list = p.q_idx + o;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tma_q);
        tma_prefetch_desc(&p.tma_k);
        tma_prefetch_desc(&p.tma_v);
        mbar_init(q_full, 1);
        for (int i = 0; i < KV_STAGES; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        }
        for (int t = 0; t < TILES; ++t) {
            mbar_init(&s_full[t], 1); mbar_init(&p_full[t], BM / 32); mbar_init(&o_done[t], 1); mbar_init(&p_half[t], BM / 32);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // the CTA owns all 512 TMEM columns (one CTA per SM by shared-memory footprint), so the allocation starts at
    // column 0 / lane 0: use the constant so that MMA operand addresses are compile-time uniform values
    if (*tmem_ptr != 0) __trap();
    constexpr uint32_t tmem_base = 0;
    constexpr uint32_t tmem_s0 = tmem_base, tmem_o0 = tmem_base + TILES * BN;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, TILES * TILE_BYTES);
            for (int t = 0; t < TILES; ++t)
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(q_smem + t * TILE_BYTES + c * SUB_BYTES, &p.tma_q, q_full, head * D + c * 64, q0 + t * BM);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_blocks; ++j) {
                const int kb = blk_list ? blk_list[j] : j;
                mbar_wait(&k_empty[stage], phase ^ 1u);
                mbar_arrive_expect_tx(&k_full[stage], TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(k_smem + stage * TILE_BYTES + c * SUB_BYTES, &p.tma_k, &k_full[stage], head * D + c * 64, kb * BN);
                mbar_wait(&v_empty[stage], phase ^ 1u);
                mbar_arrive_expect_tx(&v_full[stage], TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(v_smem + stage * TILE_BYTES + c * SUB_BYTES, &p.tma_v, &v_full[stage], head * D + c * 64, kb * BN);
                if (++stage == KV_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (whole warp converged, elected lane issues)
        {
            constexpr uint32_t idesc_s = umma_idesc_bf16(BM, BN, 0, 0);  // S = Q K^T : both K-major
            constexpr uint32_t idesc_o = umma_idesc_bf16(BM, D, 0, 1);   // O += P V  : A in TMEM, B = V MN-major
            const uint32_t q_base = smem_u32(q_smem), k_base = smem_u32(k_smem), v_base = smem_u32(v_smem);
            auto issue_s = [&](int t, int kstage) {
                const uint64_t qd = umma_desc_kmajor(q_base + t * TILE_BYTES), kd = umma_desc_kmajor(k_base + kstage * TILE_BYTES);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_ss_e(tmem_s0 + t * BN, umma_desc_advance(qd, c * SUB_BYTES + ks * 32),
                                  umma_desc_advance(kd, c * SUB_BYTES + ks * 32), idesc_s, (c | ks) ? 1u : 0u);
                umma_commit_e(&s_full[t]);
            };
            // PV in two halves of 64 K/V rows: the first is issued under the second half of the tile's exponentials
            auto issue_pv = [&](int t, int vstage, bool accumulate, int half) {
                const uint64_t vd = umma_desc_mnmajor(v_base + vstage * TILE_BYTES, SUB_BYTES);
#pragma unroll
                for (int ks = half * 4; ks < half * 4 + 4; ++ks)
                    umma_ts_e(tmem_o0 + t * D, tmem_s0 + t * BN + ks * 8, umma_desc_advance(vd, ks * 2048),
                              idesc_o, (accumulate || ks) ? 1u : 0u);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_s(0, 0);
            if (TILES == 1) umma_commit_e(&k_empty[0]);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_blocks; ++j) {
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == KV_STAGES) { nstage = 0; nphase ^= 1u; }
                mbar_wait(&v_full[stage], phase);
                for (int t = 0; t < TILES; ++t) {
                    if (TILES > 1 && j == 0 && t == 1) {
                        // tile 1 starts one softmax later than tile 0: the two softmax groups then run in anti-phase
                        // (one uses the MUFU while the other tile's MMAs run) instead of contending in phase
                        issue_s(1, 0);
                        umma_commit_e(&k_empty[0]);
                    }
                    mbar_wait(&p_half[t], j & 1);
                    tc_fence_after();
                    TL_MARK(lane == 0, j, 10 + 3 * t);
                    issue_pv(t, stage, j > 0, 0);
                    mbar_wait(&p_full[t], j & 1);
                    tc_fence_after();
                    TL_MARK(lane == 0, j, 11 + 3 * t);
                    issue_pv(t, stage, j > 0, 1);
                    if (t == TILES - 1) umma_commit_e(&v_empty[stage]);
                    if (j + 1 < n_blocks) {
                        if (t == 0) { mbar_wait(&k_full[nstage], nphase); tc_fence_after(); }
                        issue_s(t, nstage);
                        TL_MARK(lane == 0, j, 12 + 3 * t);
                        if (t == TILES - 1) umma_commit_e(&k_empty[nstage]);
                    } else {
                        umma_commit_e(&o_done[t]);
                    }
                }
                stage = nstage;
                phase = nphase;
            }
        }
    } else {
        // ------------------------------------------------------------ softmax warpgroups
        const int t = (warp - 2) >> 2;     // tile
        const int quarter = warp & 3;      // TMEM lane quarter of this warp
        const int row_in_tile = quarter * 32 + lane;
        const int q_row = q0 + t * BM + row_in_tile;
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const uint32_t s_addr = tmem_s0 + t * BN + lane_addr;
        const uint32_t o_addr = tmem_o0 + t * D + lane_addr;
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < n_blocks; ++j) {
            mbar_wait(&s_full[t], j & 1);
            tc_fence_after();
            TL_MARK(quarter == 0 && lane == 0, j, 5 * t);
            const int valid = blk_list ? BN : kv_len - j * BN;  // columns >= valid are masked (only on the last dense block)
            // the whole 128-column row of S in registers: four back-to-back TMEM loads, one wait
            uint32_t sr[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(s_addr + c * 32, sr[c]);
            tmem_ld_wait();
            TL_MARK(quarter == 0 && lane == 0, j, 5 * t + 1);
            // Lazy reference maximum: after the first block the row maximum is NOT computed (65 FMNMX3 = ~310 of the
            // ~1 860 clk of a tile's softmax).  A stale-high reference is harmless (bf16 and fp32 share the exponent
            // range; contributions 2^-126 below the running maximum are nil), a stale-low one shows up as a large row
            // sum: the exact path (maximum, O / l rescale, P again) runs only when some row's block sum exceeds
            // LAZY_SUM_LIMIT, which every element above 2^RESCALE_THRESHOLD forces.
            auto row_max = [&](bool masked) -> float {
                float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                if (!masked) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx[c] = fmaxf(mx[c], __uint_as_float(sr[c][i]));
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i < valid) mx[c] = fmaxf(mx[c], __uint_as_float(sr[c][i]));
                }
                return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * p.scale_log2;
            };
            // moves the reference maximum when it grew by > threshold; returns (warp-uniform) whether O_t / l were rescaled
            auto update_max = [&](float mb) -> bool {
                const bool grow = mb > m + RESCALE_THRESHOLD;
                if (j == 0) { m = mb; return false; }
                if (!__any_sync(0xffffffffu, grow)) return false;
                float alpha = 1.0f;
                if (grow) { alpha = fast_exp2(m - mb); m = mb; }
                l *= alpha;
                // O_t *= alpha (PV_t(j-1) has retired: S_t(j) was issued after it and has signalled s_full)
#pragma unroll 1
                for (int c = 0; c < D / 16; ++c) {
                    uint32_t r[16];
                    tmem_ld_32x32b_x16(o_addr + c * 16, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                    tmem_st_32x32b_x16(o_addr + c * 16, r);
                }
                tmem_st_wait();
                return true;
            };
            // P = exp2(s * scale - m) -> bf16 pairs over the first half of S_t; returns the row sum of the block
            auto exp_fast = [&]() -> float {
                const float neg_m = -m;
                float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const int e = (c & 1) * 16 + i;
                        float p0, p1;
                        ffma2(p0, p1, __uint_as_float(sr[c >> 1][e]), __uint_as_float(sr[c >> 1][e + 1]), p.scale_log2, neg_m);
                        p0 = fast_exp2(p0);
                        p1 = fast_exp2(p1);
                        fadd2(ls[c & 1], ls[2 + (c & 1)], p0, p1);
                        pk[i >> 1] = pack_bf16x2(p0, p1);
                    }
                    tmem_st_32x32b_x8(s_addr + c * 8, pk);
                }
                return (ls[0] + ls[1]) + (ls[2] + ls[3]);
            };
            auto exp_masked = [&]() -> float {
                const float neg_m = -m;
                float ls = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {   // fully unrolled: sr[] must stay in registers
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float p0 = fast_exp2(fmaf(__uint_as_float(sr[c][i]), p.scale_log2, neg_m));
                        float p1 = fast_exp2(fmaf(__uint_as_float(sr[c][i + 1]), p.scale_log2, neg_m));
                        if (c * 32 + i >= valid) p0 = 0.f;
                        if (c * 32 + i + 1 >= valid) p1 = 0.f;
                        ls += p0 + p1;
                        pk[i >> 1] = pack_bf16x2(p0, p1);
                    }
                    tmem_st_32x32b_x16(s_addr + c * 16, pk);
                }
                return ls;
            };
            constexpr float LAZY_SUM_LIMIT = 256.0f;   // = 2^RESCALE_THRESHOLD: one element above the threshold is enough
            float lsum;
            if (j == 0 || valid < BN) {                // first block and the masked last block: exact path
                update_max(row_max(valid < BN));
                TL_MARK(quarter == 0 && lane == 0, j, 5 * t + 2);
                lsum = valid < BN ? exp_masked() : exp_fast();
            } else {
                TL_MARK(quarter == 0 && lane == 0, j, 5 * t + 2);
                lsum = exp_fast();
                if (__any_sync(0xffffffffu, !(lsum <= LAZY_SUM_LIMIT))) {   // also catches inf / NaN sums
                    tmem_st_wait();                                        // P stores retire before O_t is touched
                    if (update_max(row_max(false))) lsum = exp_fast();     // P again with the new reference maximum
                }
            }
            l += lsum;
            tmem_st_wait();
            tc_fence_before();
            // one arrival per warp: 256 per-thread arrivals per tile and block are 256 shared-memory operations in the
            // queue the MUFU instructions go through, and each one wakes every warp sleeping on a barrier of this CTA
            mbar_arrive_warp(&p_half[t]);    // no early hand-over of the first half here: P is final only after the check
            TL_MARK(quarter == 0 && lane == 0, j, 5 * t + 3);
            if (lane == 0) mbar_arrive(&p_full[t]);
            TL_MARK(quarter == 0 && lane == 0, j, 5 * t + 4);
        }
        // ---- epilogue: O_t / l -> bf16 ; LSE
        mbar_wait(&o_done[t], 0);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const bool store = q_row < q_end;
        __nv_bfloat16* orow = p.O + (long long)q_row * p.ldo + head * D;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(o_addr + c * 32, r);
            tmem_ld_wait();
            if (store) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(r[i * 8 + 0]) * inv_l, __uint_as_float(r[i * 8 + 1]) * inv_l);
                    o.y = pack_bf16x2(__uint_as_float(r[i * 8 + 2]) * inv_l, __uint_as_float(r[i * 8 + 3]) * inv_l);
                    o.z = pack_bf16x2(__uint_as_float(r[i * 8 + 4]) * inv_l, __uint_as_float(r[i * 8 + 5]) * inv_l);
                    o.w = pack_bf16x2(__uint_as_float(r[i * 8 + 6]) * inv_l, __uint_as_float(r[i * 8 + 7]) * inv_l);
                    *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = o;
                }
            }
        }
        if (store) p.LSE[(long long)head * p.n_q + q_row] = (m + log2f(l)) * 0.6931471805599453f;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace
}  // namespace b200

using namespace b200;

static int attn_fwd_impl(void* O, int64_t ldo, float* LSE, const void* Q, int64_t ldq, const void* K, int64_t ldk,
                         const void* V, int64_t ldv, int32_t n_q, int32_t n_kv, int32_t heads, float softmax_scale,
                         const b200tta_attn_seg* segs, int32_t n_seg, const int32_t* q_off, const int32_t* q_idx,
                         b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(O && LSE && Q && K && V && n_q > 0 && n_kv > 0 && heads > 0, "attn_fwd: null/empty argument");
    B200_REQUIRE(n_seg >= 1 && n_seg <= MAX_SEGS && segs, "attn_fwd: n_seg=%d not in [1,%d]", n_seg, MAX_SEGS);
    B200_REQUIRE(aligned16(O) && aligned16(Q) && aligned16(K) && aligned16(V) && ldo % 8 == 0 && ldq % 8 == 0 &&
                     ldk % 8 == 0 && ldv % 8 == 0 && ldq >= (int64_t)heads * D && ldk >= (int64_t)heads * D &&
                     ldv >= (int64_t)heads * D && ldo >= (int64_t)heads * D,
                 "attn_fwd: tensors must be 16-byte aligned [tokens, heads, 128] views");
    const int tiles = q_off ? 1 : 2;
    FwdParams p;
    memset(&p, 0, sizeof(p));
    if (int rc = make_tmap_2d_bf16(&p.tma_q, Q, (uint64_t)heads * D, (uint64_t)n_q, (uint64_t)ldq * 2, 64, BM)) return rc;
    if (int rc = make_tmap_2d_bf16(&p.tma_k, K, (uint64_t)heads * D, (uint64_t)n_kv, (uint64_t)ldk * 2, 64, BN)) return rc;
    if (int rc = make_tmap_2d_bf16(&p.tma_v, V, (uint64_t)heads * D, (uint64_t)n_kv, (uint64_t)ldv * 2, 64, BN)) return rc;
    p.O = (__nv_bfloat16*)O; p.ldo = ldo; p.LSE = LSE; p.n_q = n_q; p.heads = heads;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    p.n_seg = n_seg;
    p.q_off = q_off; p.q_idx = q_idx;
    int items = 0;
    for (int s = 0; s < n_seg; ++s) {
        B200_REQUIRE(segs[s].q_begin >= 0 && segs[s].q_end > segs[s].q_begin && segs[s].q_end <= n_q &&
                         segs[s].kv_len > 0 && segs[s].kv_len <= n_kv,
                     "attn_fwd: bad segment %d [%d,%d) kv %d", s, segs[s].q_begin, segs[s].q_end, segs[s].kv_len);
        p.seg_q_begin[s] = segs[s].q_begin; p.seg_q_end[s] = segs[s].q_end; p.seg_kv_len[s] = segs[s].kv_len;
        p.seg_item0[s] = items;
        items += (segs[s].q_end - segs[s].q_begin + BM * tiles - 1) / (BM * tiles);
    }
    p.seg_item0[n_seg] = items;
    static bool attr = false;
    if (!attr) {
        B200_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(2)));
        B200_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(1)));
        attr = true;
    }
    if (tiles == 2) attn_fwd_kernel<2><<<dim3(items, heads), fwd_threads(2), fwd_smem(2), (cudaStream_t)stream>>>(p);
    else attn_fwd_kernel<1><<<dim3(items, heads), fwd_threads(1), fwd_smem(1), (cudaStream_t)stream>>>(p);
    B200_LAUNCHED();
    return B200TTA_OK;
}

#if B200TTA_ATTN_DEBUG
extern "C" int b200tta_debug_fwd_timeline(long long* host_out) {   // TL_BLOCKS x TL_SLOTS clock values
    return cudaMemcpyFromSymbol(host_out, b200::g_timeline, sizeof(long long) * 24 * 16) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int b200tta_attn_fwd(void* O, int64_t ldo, float* LSE, const void* Q, int64_t ldq, const void* K,
                                int64_t ldk, const void* V, int64_t ldv, int32_t n_q, int32_t n_kv, int32_t heads,
                                float softmax_scale, const b200tta_attn_seg* segs, int32_t n_seg,
                                b200tta_stream_t stream) {
    return attn_fwd_impl(O, ldo, LSE, Q, ldq, K, ldk, V, ldv, n_q, n_kv, heads, softmax_scale, segs, n_seg, nullptr, nullptr,
                         stream);
}

extern "C" int b200tta_attn_bsa_fwd(void* O, int64_t ldo, float* LSE, const void* Q, int64_t ldq, const void* K,
                                    int64_t ldk, const void* V, int64_t ldv, int32_t n_tok, int32_t heads,
                                    float softmax_scale, const int32_t* q_off, const int32_t* q_idx,
                                    b200tta_stream_t stream) {
    B200_REQUIRE(q_off && q_idx, "attn_bsa_fwd: null block list");
    B200_REQUIRE(n_tok > 0 && n_tok % BM == 0, "attn_bsa_fwd: n_tok=%d must be a multiple of the 128-token block", n_tok);
    const b200tta_attn_seg seg = {0, n_tok, n_tok};
    return attn_fwd_impl(O, ldo, LSE, Q, ldq, K, ldk, V, ldv, n_tok, n_tok, heads, softmax_scale, &seg, 1, q_off, q_idx, stream);
}
