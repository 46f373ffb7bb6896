// Multi-segment bf16 GEMM for sm_100a: TMA -> shared (128B swizzle) -> tcgen05.mma -> TMEM -> fused epilogue.
//
//   D[M,N] = epilogue( sum_s A_s[M,K_s] * op(B_s) )
//
// Two kernels share the epilogues and the tile walk:
//   gemm2_kernel   (default for 256-wide N tiles) a CLUSTER OF TWO CTAs owns a 256 x 256 tile: CTA r stages its 128 rows
//                  of A and half of the B tile, the leader issues tcgen05.mma.cta_group::2 (M = 256) from TWO issuer
//                  warps that alternate k-blocks, tcgen05.commit multicasts the stage release to both CTAs, each CTA
//                  drains its own 128 accumulator rows.  See the comment above the kernel for the why.
//   gemm_kernel<BN> one persistent CTA per SM, 6 warps: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane)
//                  + TMEM owner, warps 2..5 = epilogue (one TMEM lane quarter each).  Used for N < 256 (BN = 64) and,
//                  with B200TTA_DETERMINISTIC=1, for everything (fixed summation order).
// Three pipelines in both: smem full/empty ring (TMA <-> MMA), two TMEM accumulator buffers (MMA <-> epilogue) and the
// static tile schedule (grouped along M so that a wave of tiles re-uses A and B through L2).
//
// The K loop runs over up to 4 segments; segment s may present B either K-major ([N,K], forward
// x @ W^T) or MN-major ([K,N], backward dy @ W -- no transposed copy of the frozen weight), and
// a rank-r LoRA product is simply one more (short) segment accumulating into the same TMEM tile.
#include "host_common.h"
#include "ptx.cuh"
#include <stdlib.h>

namespace b200 {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int MAX_SEG = 4;
constexpr int GROUP_M = 16;
constexpr int NUM_THREADS = 192;

struct EpiArgs {
    int mode, tokens_per_frame;
    void* d; long long ldd;
    void* d2; long long ldd2;
    void* d3; long long ldd3;
    const void* bias; int bias_is_f32;
    const void* resid; long long ldr;
    const float* gate; long long ldg;
    const void* aux1; long long ldaux1;
    const void* aux2; long long ldaux2;
};

struct alignas(64) GemmParams {
    CUtensorMap tma_a[MAX_SEG];
    CUtensorMap tma_b[MAX_SEG];
    CUtensorMap tma_bhi[MAX_SEG];
    CUtensorMap tma_b128[MAX_SEG];   // K-major B with a 128-row box: one CTA's half of the N tile in 2-CTA mode
    int k[MAX_SEG];
    int b_mn[MAX_SEG];
    int a_mn[MAX_SEG];   // A stored [K, M] (M contiguous): weight gradients dW = dY^T X read dY [tokens, out] in place
    int has_hi[MAX_SEG];
    int nseg;
    int M, N;
    int m_tiles, n_tiles;
    int group_m2;   // M-grouping of the CTA-pair kernel's tile walk (256-row tiles)
    EpiArgs epi;
};

template <int BN>
struct Cfg {
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void tile_coords(int tile, int m_tiles, int n_tiles, int& m_blk, int& n_blk, int group_m = GROUP_M) {
    const int per_group = group_m * n_tiles;
    const int group = tile / per_group;
    const int first_m = group * group_m;
    const int rows = min(group_m, m_tiles - first_m);
    const int in_group = tile - group * per_group;
    m_blk = first_m + in_group % rows;
    n_blk = in_group / rows;
}

__device__ __forceinline__ float gelu_tanh(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    return 0.5f * x * (1.0f + tanhf(k0 * (x + k1 * x * x * x)));
}

__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* p, float (&out)[32]) {
    const uint4* p4 = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 v = __ldg(p4 + i);
        float2 f;
        f = unpack_bf16x2(v.x); out[i * 8 + 0] = f.x; out[i * 8 + 1] = f.y;
        f = unpack_bf16x2(v.y); out[i * 8 + 2] = f.x; out[i * 8 + 3] = f.y;
        f = unpack_bf16x2(v.z); out[i * 8 + 4] = f.x; out[i * 8 + 5] = f.y;
        f = unpack_bf16x2(v.w); out[i * 8 + 6] = f.x; out[i * 8 + 7] = f.y;
    }
}
__device__ __forceinline__ void load_f32x32(const float* p, float (&out)[32]) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 v = __ldg(p4 + i);
        out[i * 4 + 0] = v.x; out[i * 4 + 1] = v.y; out[i * 4 + 2] = v.z; out[i * 4 + 3] = v.w;
    }
}
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* p, const float (&v)[32]) {
    uint4* p4 = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 o;
        o.x = pack_bf16x2(v[i * 8 + 0], v[i * 8 + 1]);
        o.y = pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
        o.z = pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]);
        o.w = pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
        p4[i] = o;
    }
}
__device__ __forceinline__ void store_f32x32(float* p, const float (&v)[32]) {
    float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) p4[i] = make_float4(v[i * 4 + 0], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
}

__device__ __forceinline__ void add_bias(const EpiArgs& e, int col, float (&acc)[32]) {
    if (e.bias == nullptr) return;
    float b[32];
    if (e.bias_is_f32) load_f32x32(reinterpret_cast<const float*>(e.bias) + col, b);
    else load_bf16x32(reinterpret_cast<const __nv_bfloat16*>(e.bias) + col, b);
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] += b[i];
}

// One thread owns `row` and 32 consecutive accumulator columns starting at `col`.
__device__ __forceinline__ void epilogue_chunk(const EpiArgs& e, long long row, int col, float (&acc)[32]) {
    switch (e.mode) {
        case B200TTA_EPI_STORE_F32: {
            add_bias(e, col, acc);
            store_f32x32(reinterpret_cast<float*>(e.d) + row * e.ldd + col, acc);
        } break;
        case B200TTA_EPI_GELU: {
            add_bias(e, col, acc);
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = gelu_tanh(acc[i]);
            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(e.d) + row * e.ldd + col, acc);
        } break;
        case B200TTA_EPI_GATE_RESID: {
            add_bias(e, col, acc);
            if (e.d2 != nullptr) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(e.d2) + row * e.ldd2 + col, acc);
            float r[32];
            load_bf16x32(reinterpret_cast<const __nv_bfloat16*>(e.resid) + row * e.ldr + col, r);
            if (e.gate != nullptr) {
                float g[32];
                const long long f = row / e.tokens_per_frame;
                load_f32x32(e.gate + f * e.ldg + col, g);
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = fmaf(g[i], acc[i], r[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] += r[i];
            }
            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(e.d) + row * e.ldd + col, acc);
        } break;
        case B200TTA_EPI_SWIGLU_BWD: {
            float h1[32], h3[32], o1[32];
            load_bf16x32(reinterpret_cast<const __nv_bfloat16*>(e.aux1) + row * e.ldaux1 + col, h1);
            load_bf16x32(reinterpret_cast<const __nv_bfloat16*>(e.aux2) + row * e.ldaux2 + col, h3);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float sig = 1.0f / (1.0f + __expf(-h1[i]));
                const float silu = h1[i] * sig;
                const float dsilu = sig * (1.0f + h1[i] * (1.0f - sig));
                o1[i] = acc[i] * h3[i] * dsilu;
                acc[i] = acc[i] * silu;
            }
            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(e.d) + row * e.ldd + col, o1);
            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(e.d2) + row * e.ldd2 + col, acc);
        } break;
        default: {  // B200TTA_EPI_STORE
            add_bias(e, col, acc);
            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(e.d) + row * e.ldd + col, acc);
        } break;
    }
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tmem_full = empty_bar + C::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.m_tiles * p.n_tiles;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) {
            tma_prefetch_desc(&p.tma_a[s]);
            tma_prefetch_desc(&p.tma_b[s]);
            if (p.has_hi[s]) tma_prefetch_desc(&p.tma_bhi[s]);
        }
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int m_blk, n_blk;
                tile_coords(tile, p.m_tiles, p.n_tiles, m_blk, n_blk);
                const int m0 = m_blk * BM, n0 = n_blk * BN;
                for (int s = 0; s < p.nseg; ++s) {
                    const int kblocks = (p.k[s] + BK - 1) / BK;
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        uint8_t* sa = smem + stage * C::STAGE_BYTES;
                        uint8_t* sb = sa + C::A_BYTES;
                        mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                        if (p.a_mn[s]) {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j)
                                tma_load_2d(sa + j * (64 * BK * 2), &p.tma_a[s], &full_bar[stage], m0 + j * 64, kb * BK);
                        } else {
                            tma_load_2d(sa, &p.tma_a[s], &full_bar[stage], kb * BK, m0);
                        }
                        if (p.b_mn[s]) {
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j)
                                tma_load_2d(sb + j * (64 * BK * 2), &p.tma_b[s], &full_bar[stage], n0 + j * 64, kb * BK);
                        } else if (p.has_hi[s]) {
                            tma_load_2d(sb, &p.tma_b[s], &full_bar[stage], kb * BK, n0 / 2);
                            tma_load_2d(sb + C::B_BYTES / 2, &p.tma_bhi[s], &full_bar[stage], kb * BK, n0 / 2);
                        } else {
                            tma_load_2d(sb, &p.tma_b[s], &full_bar[stage], kb * BK, n0);
                        }
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (whole warp converged, elected lane issues)
        {
            constexpr uint32_t idesc_k = umma_idesc_bf16(BM, BN, 0, 0);
            constexpr uint32_t idesc_mn = umma_idesc_bf16(BM, BN, 0, 1);
            const uint32_t smem_base = smem_u32(smem);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                uint32_t accumulate = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const int kblocks = (p.k[s] + BK - 1) / BK;
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                        const uint32_t sb = sa + C::A_BYTES;
                        const int rem = p.k[s] - kb * BK;
                        const int ksteps = rem >= BK ? BK / 16 : (rem + 15) / 16;
                        const uint64_t ad = umma_desc_kmajor(sa);
                        if (p.a_mn[s]) {      // both operands token-major (dW = dY^T X): requires b_mn as well (host-checked)
                            constexpr uint32_t idesc_mm = umma_idesc_bf16(BM, BN, 1, 1);
                            const uint64_t am = umma_desc_mnmajor(sa, 64 * BK * 2), bd = umma_desc_mnmajor(sb, 64 * BK * 2);
#pragma unroll
                            for (int ks = 0; ks < BK / 16; ++ks) {
                                if (ks < ksteps) {
                                    umma_ss_e(d_tmem, umma_desc_advance(am, ks * 2048), umma_desc_advance(bd, ks * 2048), idesc_mm, accumulate);
                                    accumulate = 1;
                                }
                            }
                        } else if (p.b_mn[s]) {
                            const uint64_t bd = umma_desc_mnmajor(sb, 64 * BK * 2);
#pragma unroll
                            for (int ks = 0; ks < BK / 16; ++ks) {
                                if (ks < ksteps) {
                                    umma_ss_e(d_tmem, umma_desc_advance(ad, ks * 32), umma_desc_advance(bd, ks * 2048), idesc_mn, accumulate);
                                    accumulate = 1;
                                }
                            }
                        } else {
                            const uint64_t bd = umma_desc_kmajor(sb);
#pragma unroll
                            for (int ks = 0; ks < BK / 16; ++ks) {
                                if (ks < ksteps) {
                                    umma_ss_e(d_tmem, umma_desc_advance(ad, ks * 32), umma_desc_advance(bd, ks * 32), idesc_k, accumulate);
                                    accumulate = 1;
                                }
                            }
                        }
                        umma_commit_e(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
                umma_commit_e(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32)
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int m_blk, n_blk;
            tile_coords(tile, p.m_tiles, p.n_tiles, m_blk, n_blk);
            const long long row = (long long)m_blk * BM + quarter * 32 + lane;
            const int n0 = n_blk * BN;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
            if (p.epi.mode == B200TTA_EPI_SWIGLU || p.epi.mode == B200TTA_EPI_GEGLU) {
                // accumulator columns [0, BN/2) = h1, [BN/2, BN) = h3 of output features n0/2 ...
#pragma unroll 1
                for (int c = 0; c < BN / 64; ++c) {
                    uint32_t r1[32], r3[32];
                    tmem_ld_32x32b_x32(t_addr + c * 32, r1);
                    tmem_ld_32x32b_x32(t_addr + BN / 2 + c * 32, r3);
                    tmem_ld_wait();
                    const int col = n0 / 2 + c * 32;
                    if (row < p.M && col < p.N / 2) {
                        float h1[32], h3[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) { h1[i] = __uint_as_float(r1[i]); h3[i] = __uint_as_float(r3[i]); }
                        if (p.epi.d2 != nullptr)
                            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.epi.d2) + row * p.epi.ldd2 + col, h1);
                        if (p.epi.d3 != nullptr)
                            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.epi.d3) + row * p.epi.ldd3 + col, h3);
                        if (p.epi.mode == B200TTA_EPI_GEGLU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) h1[i] = gelu_tanh(h1[i]) * h3[i];
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) h1[i] = h1[i] / (1.0f + __expf(-h1[i])) * h3[i];
                        }
                        store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.epi.d) + row * p.epi.ldd + col, h1);
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_addr + c * 32, r);
                    tmem_ld_wait();
                    const int col = n0 + c * 32;
                    if (row < p.M && col < p.N) {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                        epilogue_chunk(p.epi, row, col, v);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------ 2-CTA form (cta_group::2)
// A cluster of two CTAs (the two SMs of a TPC) owns a 256 x 256 output tile: CTA r stages rows [128 r, 128 r + 128) of A
// and HALF of the B tile (N rows [128 r, 128 r + 128)), the leader issues tcgen05.mma.cta_group::2 with M = 256 and the
// hardware feeds both tensor cores from both shared memories.  Per CTA and k-block that is 8 KB of operand fetch per
// 128 clk (64 B/clk) plus 32 KB of TMA fill per 512 clk (64 B/clk) -- inside the 128 B/clk of shared memory, where the
// 1-CTA 128 x 256 tile needs 96 + 94 B/clk (ncu: 74 % tensor-pipe active).  Each CTA drains its own 128 accumulator rows.
constexpr int STAGES2 = 6;
constexpr int A2_BYTES = BM * BK * 2, B2_BYTES = 128 * BK * 2, STAGE2_BYTES = A2_BYTES + B2_BYTES;
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256;
constexpr int NUM_THREADS2 = 224;   // warp 0 TMA, warps 1 and 6 MMA issuers (leader), warps 2-5 epilogue

__global__ void __launch_bounds__(NUM_THREADS2, 1) gemm2_kernel(const __grid_constant__ GemmParams p) {
    constexpr int BN = 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES2 * STAGE2_BYTES);
    uint64_t* empty_bar = full_bar + STAGES2;
    uint64_t* tmem_full = empty_bar + STAGES2;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int m_tiles2 = (p.M + 2 * BM - 1) / (2 * BM);
    const int num_tiles = m_tiles2 * p.n_tiles;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) {
            tma_prefetch_desc(&p.tma_a[s]);
            tma_prefetch_desc(p.b_mn[s] ? &p.tma_b[s] : (p.has_hi[s] ? (rank ? &p.tma_bhi[s] : &p.tma_b[s]) : &p.tma_b128[s]));
        }
        for (int i = 0; i < STAGES2; ++i) {
            mbar_init(&full_bar[i], 1);     // the leader's own arrive.expect_tx; bytes of both CTAs are counted here
            mbar_init(&empty_bar[i], 1);    // multicast tcgen05.commit of the leader
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 2);    // multicast tcgen05.commit of each of the leader's two issuers
            mbar_init(&tmem_empty[i], 256); // (leader's copy) the epilogue threads of BOTH CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2cta<512>(tmem_ptr);
    tc_fence_before();
    cluster_sync();                         // barrier inits and the TMEM allocation of the peer are visible
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs, own halves)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
                int m_blk, n_blk;
                tile_coords(tile, m_tiles2, p.n_tiles, m_blk, n_blk, p.group_m2);
                const int m0 = m_blk * 2 * BM + rank * BM, n0 = n_blk * BN;
                for (int s = 0; s < p.nseg; ++s) {
                    const int kblocks = (p.k[s] + BK - 1) / BK;
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        uint8_t* sa = smem + stage * STAGE2_BYTES;
                        uint8_t* sb = sa + A2_BYTES;
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE2_BYTES);
                        tma_load_2d_2cta(sa, &p.tma_a[s], &full_bar[stage], kb * BK, m0);
                        if (p.b_mn[s]) {
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                tma_load_2d_2cta(sb + j * (64 * BK * 2), &p.tma_b[s], &full_bar[stage], n0 + (2 * rank + j) * 64, kb * BK);
                        } else if (p.has_hi[s]) {   // w1 | w3 co-tiling: the leader holds the w1 rows, its peer the w3 rows
                            tma_load_2d_2cta(sb, rank ? &p.tma_bhi[s] : &p.tma_b[s], &full_bar[stage], kb * BK, n0 / 2);
                        } else {
                            tma_load_2d_2cta(sb, &p.tma_b128[s], &full_bar[stage], kb * BK, n0 + rank * 128);
                        }
                        if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 6) {
        // ------------------------------------------------------------ MMA issuers: leader CTA only, two warps alternate
        // k-blocks: a barrier poll stalls the polling thread's MMA queue for ~100 clk (scratch/mma_issue.cu) and the other
        // issuer's MMAs cover it.  The two issuers run free (a token that forces the issue order was measured: it
        // serialises them again and loses the gain), so every MMA accumulates, the epilogue leaves the accumulator
        // zeroed, and the fp32 summation order over k-blocks is not fixed run to run (B200TTA_DETERMINISTIC=1 selects
        // the single-issuer 1-CTA kernel instead).
        if (rank == 0) {
            const int w = warp == 1 ? 0 : 1;
            constexpr uint32_t idesc_k = umma_idesc_bf16(2 * BM, BN, 0, 0);
            constexpr uint32_t idesc_mn = umma_idesc_bf16(2 * BM, BN, 0, 1);
            const uint32_t smem_base = smem_u32(smem);
            long long g = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
                mbar_wait(&tmem_empty[acc], acc_phase);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int s = 0; s < p.nseg; ++s) {
                    const int kblocks = (p.k[s] + BK - 1) / BK;
                    for (int kb = 0; kb < kblocks; ++kb, ++g) {
                        if ((g & 1) != w) continue;
                        const int stage = (int)(g % STAGES2);
                        const uint32_t phase = (uint32_t)(g / STAGES2) & 1u;
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + stage * STAGE2_BYTES;
                        const uint32_t sb = sa + A2_BYTES;
                        const int rem = p.k[s] - kb * BK;
                        const int ksteps = rem >= BK ? BK / 16 : (rem + 15) / 16;
                        const uint64_t ad = umma_desc_kmajor(sa);
                        if (p.b_mn[s]) {
                            const uint64_t bd = umma_desc_mnmajor(sb, 64 * BK * 2);
#pragma unroll
                            for (int ks = 0; ks < BK / 16; ++ks)
                                if (ks < ksteps)
                                    umma_ss_2cta_e(d_tmem, umma_desc_advance(ad, ks * 32), umma_desc_advance(bd, ks * 2048), idesc_mn, 1u);
                        } else {
                            const uint64_t bd = umma_desc_kmajor(sb);
#pragma unroll
                            for (int ks = 0; ks < BK / 16; ++ks)
                                if (ks < ksteps)
                                    umma_ss_2cta_e(d_tmem, umma_desc_advance(ad, ks * 32), umma_desc_advance(bd, ks * 32), idesc_k, 1u);
                        }
                        umma_commit_2cta_e(&empty_bar[stage]);   // frees the slot in BOTH CTAs when these MMAs retire
                    }
                }
                umma_commit_2cta_e(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp >= 2 && warp <= 5) {
        // ------------------------------------------------------------ epilogue (warps 2..5 of both CTAs, own 128 rows)
        const int quarter = warp & 3;
        uint32_t zero[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) zero[i] = 0u;
        for (int a2 = 0; a2 < 2; ++a2) {   // every MMA accumulates: hand both accumulators over zeroed
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c)
                tmem_st_32x32b_x32(tmem_base + (uint32_t(quarter * 32) << 16) + a2 * BN + c * 32, zero);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_leader(&tmem_empty[a2]);
        }
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
            int m_blk, n_blk;
            tile_coords(tile, m_tiles2, p.n_tiles, m_blk, n_blk, p.group_m2);
            const long long row = (long long)m_blk * 2 * BM + rank * BM + quarter * 32 + lane;
            const int n0 = n_blk * BN;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
            if (p.epi.mode == B200TTA_EPI_SWIGLU || p.epi.mode == B200TTA_EPI_GEGLU) {
#pragma unroll 1
                for (int c = 0; c < BN / 64; ++c) {
                    uint32_t r1[32], r3[32];
                    tmem_ld_32x32b_x32(t_addr + c * 32, r1);
                    tmem_ld_32x32b_x32(t_addr + BN / 2 + c * 32, r3);
                    tmem_ld_wait();
                    tmem_st_32x32b_x32(t_addr + c * 32, zero);
                    tmem_st_32x32b_x32(t_addr + BN / 2 + c * 32, zero);
                    const int col = n0 / 2 + c * 32;
                    if (row < p.M && col < p.N / 2) {
                        float h1[32], h3[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) { h1[i] = __uint_as_float(r1[i]); h3[i] = __uint_as_float(r3[i]); }
                        if (p.epi.d2 != nullptr)
                            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.epi.d2) + row * p.epi.ldd2 + col, h1);
                        if (p.epi.d3 != nullptr)
                            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.epi.d3) + row * p.epi.ldd3 + col, h3);
                        if (p.epi.mode == B200TTA_EPI_GEGLU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) h1[i] = gelu_tanh(h1[i]) * h3[i];
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) h1[i] = h1[i] / (1.0f + __expf(-h1[i])) * h3[i];
                        }
                        store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.epi.d) + row * p.epi.ldd + col, h1);
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_addr + c * 32, r);
                    tmem_ld_wait();
                    tmem_st_32x32b_x32(t_addr + c * 32, zero);
                    const int col = n0 + c * 32;
                    if (row < p.M && col < p.N) {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                        epilogue_chunk(p.epi, row, col, v);
                    }
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_leader(&tmem_empty[acc]);   // the leader's issuers wait for both CTAs' epilogues
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    cluster_sync();                         // neither CTA may free TMEM / exit while its peer still uses the pair
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta<512>(tmem_base);
    }
}

int launch2(const GemmParams& p, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        B200_CUDA(cudaFuncSetAttribute(gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
        attr_set = true;
    }
    const int m_tiles2 = (p.M + 2 * BM - 1) / (2 * BM);
    const int tiles = m_tiles2 * p.n_tiles;
    const int max_clusters = sm_count() / 2;
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(NUM_THREADS2);
    cfg.dynamicSmemBytes = SMEM2_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA(cudaLaunchKernelEx(&cfg, gemm2_kernel, p));
    B200_LAUNCHED();
    return B200TTA_OK;
}

template <int BN>
int launch(const GemmParams& p, cudaStream_t stream) {
    using C = Cfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        B200_CUDA(cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    gemm_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(p);
    B200_LAUNCHED();
    return B200TTA_OK;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200tta_gemm(int64_t M, int64_t N, const b200tta_gemm_seg* segs, int32_t nseg,
                            const b200tta_gemm_epi* epi, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(M > 0 && N > 0 && M < (1ll << 31) && N < (1ll << 31), "gemm: bad M=%lld N=%lld", (long long)M, (long long)N);
    B200_REQUIRE(nseg >= 1 && nseg <= MAX_SEG, "gemm: nseg=%d not in [1,%d]", nseg, MAX_SEG);
    B200_REQUIRE(segs && epi && epi->d, "gemm: null segs/epi/output");
    const bool swiglu = epi->mode == B200TTA_EPI_SWIGLU || epi->mode == B200TTA_EPI_GEGLU;
    B200_REQUIRE(N % 64 == 0, "gemm: N=%lld must be a multiple of 64", (long long)N);
    const int BN = (N % 256 == 0 || N > 256) ? 256 : 64;
    // (Tried for few-row problems such as the text encoder's 512-token prompts: 128 x 128 single-CTA tiles to fill the chip
    // -- 128 instead of 32 tiles for a 4096-wide layer.  Slower, 7.6 vs 6.0 ms per UMT5-xxl encode: a 128 x 128 tile pulls
    // 128 B/clk per SM from L2 where a CTA pair pulls 64, and the L2 cannot feed 128 SMs at that rate.)
    B200_REQUIRE(BN == 64 || N % 32 == 0, "gemm: N tail");
    B200_REQUIRE(!swiglu || (N % 256 == 0), "gemm: SWIGLU / GEGLU need N (=2F) %% 256 == 0");

    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.nseg = nseg;
    p.M = (int)M;
    p.N = (int)N;
    p.m_tiles = (int)((M + BM - 1) / BM);
    p.n_tiles = (int)((N + BN - 1) / BN);
    { const char* gm = getenv("B200TTA_GEMM_GROUP_M2"); p.group_m2 = gm ? atoi(gm) : 8; }
    int any_a_mn = 0;
    for (int s = 0; s < nseg; ++s) {
        const b200tta_gemm_seg& g = segs[s];
        B200_REQUIRE(g.a && g.b && g.k > 0, "gemm: segment %d has null operand or k<=0", s);
        B200_REQUIRE(aligned16(g.a) && aligned16(g.b) && g.lda % 8 == 0 && g.ldb % 8 == 0,
                     "gemm: segment %d operands must be 16-byte aligned with ld %% 8 == 0 (lda=%lld ldb=%lld)", s,
                     (long long)g.lda, (long long)g.ldb);
        p.k[s] = (int)g.k;
        p.b_mn[s] = g.b_mn_major ? 1 : 0;
        p.a_mn[s] = g.a_mn_major ? 1 : 0;
        any_a_mn |= p.a_mn[s];
        B200_REQUIRE(!p.a_mn[s] || g.b_mn_major, "gemm: an MN-major A needs an MN-major B (weight-gradient form)");
        p.has_hi[s] = g.b_hi ? 1 : 0;
        B200_REQUIRE(!(g.b_hi && g.b_mn_major), "gemm: b_hi requires a K-major B");
        B200_REQUIRE(!g.b_hi || BN == 256, "gemm: b_hi requires 256-wide N tiles");
        if (p.a_mn[s]) {
            if (int rc = make_tmap_2d_bf16(&p.tma_a[s], g.a, (uint64_t)M, (uint64_t)g.k, (uint64_t)g.lda * 2, 64, BK)) return rc;
        } else if (int rc = make_tmap_2d_bf16(&p.tma_a[s], g.a, (uint64_t)g.k, (uint64_t)M, (uint64_t)g.lda * 2, BK, BM)) return rc;
        if (g.b_mn_major) {
            if (int rc = make_tmap_2d_bf16(&p.tma_b[s], g.b, (uint64_t)N, (uint64_t)g.k, (uint64_t)g.ldb * 2, 64, BK)) return rc;
        } else if (g.b_hi) {
            B200_REQUIRE(aligned16(g.b_hi), "gemm: b_hi alignment");
            if (int rc = make_tmap_2d_bf16(&p.tma_b[s], g.b, (uint64_t)g.k, (uint64_t)N / 2, (uint64_t)g.ldb * 2, BK, BN / 2)) return rc;
            if (int rc = make_tmap_2d_bf16(&p.tma_bhi[s], g.b_hi, (uint64_t)g.k, (uint64_t)N / 2, (uint64_t)g.ldb * 2, BK, BN / 2)) return rc;
        } else {
            if (int rc = make_tmap_2d_bf16(&p.tma_b[s], g.b, (uint64_t)g.k, (uint64_t)N, (uint64_t)g.ldb * 2, BK, BN)) return rc;
            if (BN == 256)
                if (int rc = make_tmap_2d_bf16(&p.tma_b128[s], g.b, (uint64_t)g.k, (uint64_t)N, (uint64_t)g.ldb * 2, BK, 128)) return rc;
        }
    }
    EpiArgs& e = p.epi;
    e.mode = epi->mode; e.tokens_per_frame = epi->tokens_per_frame > 0 ? epi->tokens_per_frame : 1;
    e.d = epi->d; e.ldd = epi->ldd; e.d2 = epi->d2; e.ldd2 = epi->ldd2; e.d3 = epi->d3; e.ldd3 = epi->ldd3;
    e.bias = epi->bias; e.bias_is_f32 = epi->bias_is_f32;
    e.resid = epi->resid; e.ldr = epi->ldr; e.gate = epi->gate; e.ldg = epi->ldg;
    e.aux1 = epi->aux1; e.ldaux1 = epi->ldaux1; e.aux2 = epi->aux2; e.ldaux2 = epi->ldaux2;
    B200_REQUIRE(aligned16(e.d) && e.ldd % 8 == 0, "gemm: output must be 16-byte aligned, ldd %% 8 == 0");
    if (e.mode == B200TTA_EPI_GATE_RESID)
        B200_REQUIRE(e.resid && aligned16(e.resid) && e.ldr % 8 == 0 && (!e.gate || (aligned16(e.gate) && e.ldg % 4 == 0)),
                     "gemm: GATE_RESID needs aligned resid/gate");
    if (e.mode == B200TTA_EPI_SWIGLU_BWD)
        B200_REQUIRE(e.aux1 && e.aux2 && e.d2 && aligned16(e.aux1) && aligned16(e.aux2) && aligned16(e.d2) &&
                         e.ldaux1 % 8 == 0 && e.ldaux2 % 8 == 0 && e.ldd2 % 8 == 0,
                     "gemm: SWIGLU_BWD needs h1, h3 (aux1, aux2) and d2");
    B200_REQUIRE(!e.bias || aligned16(e.bias), "gemm: bias alignment");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // 256-wide tiles run on CTA pairs with two free-running issuers (fastest; summation order over k-blocks not fixed run
    // to run).  B200TTA_DETERMINISTIC=1 (read per call) selects the single-issuer 1-CTA kernel: same bits every run.
    const char* det = getenv("B200TTA_DETERMINISTIC");
    if (BN == 256 && !(det && det[0] == '1') && !any_a_mn) return launch2(p, st);   // MN-major A: 1-CTA kernel only
    return BN == 256 ? launch<256>(p, st) : launch<64>(p, st);
}
