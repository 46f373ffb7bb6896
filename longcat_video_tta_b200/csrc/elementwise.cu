// Fused elementwise / row-reduction kernels of the DiT block (HBM-bound): LayerNorm+modulate,
// per-head RMSNorm + 3-D RoPE, gate multiply, noising + patchify, MSE, timestep embedding pieces.
// One warp owns one token row; 128-bit global accesses; reductions by warp shuffle only.
#include "host_common.h"
#include "ptx.cuh"

namespace b200 {
namespace {

constexpr int WARPS_PER_BLOCK = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    float2 t;
    t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
    t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
    t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
// same as unpack8, but opaque to common-subexpression elimination: a kernel that keeps a row PACKED in registers and
// unpacks it once per pass must not have the unpacked copy kept live across passes (that doubles the register need)
// (`tag` differs per pass: identical non-volatile asm statements would be merged, volatile ones cannot be reordered --
// the passes below need both a private unpack per pass and freedom to interleave it with the arithmetic)
template <int TAG>
__device__ __forceinline__ void unpack8_opaque(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t lo, hi;
        asm("shl.b32 %0, %1, 16; // %2" : "=r"(lo) : "r"(w[i]), "n"(TAG));
        asm("and.b32 %0, %1, 0xffff0000; // %2" : "=r"(hi) : "r"(w[i]), "n"(TAG));
        f[2 * i] = __uint_as_float(lo);
        f[2 * i + 1] = __uint_as_float(hi);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    return u;
}
// 8 consecutive parameters starting at element `idx` of a bf16 or f32 vector
__device__ __forceinline__ void load_param8(const void* base, long long idx, int is_bf16, float (&f)[8]) {
    if (is_bf16) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx)), f);
    } else {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
        float4 a = __ldg(p), b = __ldg(p + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
}

// ------------------------------------------------------------------------------------ LayerNorm + modulate
// grid (blocks per frame, frames); a block owns LN_FWD_ROWS consecutive rows of ONE frame and stages that frame's
// (mul_base + scale) and shift as fp32 in shared memory once (32 KB at C = 4096), laid out so that every lane's eight
// values are two conflict-free 16-byte reads.  Round-2 ncu of the previous version (every warp re-read both fp32
// parameter rows, 32 KB per 8 KB token row, through L1 with a 32-byte lane stride): 86 M L1 sectors for 9.6 M of data,
// long-scoreboard stalls 9.8 of 13.3 cycles per issue, 3.0 TB/s.
constexpr int LN_FWD_ROWS = 32;
template <int CH>  // C = CH * 256
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 2) ln_mod_fwd_kernel(
    __nv_bfloat16* __restrict__ Y, long long ldy, const __nv_bfloat16* __restrict__ X, long long ldx,
    const void* __restrict__ scale, const void* __restrict__ shift, long long mod_ld, int params_bf16, float mul_base,
    long long rows, int tpf, float eps) {
    constexpr int C = CH * 256;
    extern __shared__ float4 ln_prm[];   // [CH][4][32]: per chunk c, float4 slot s (scale lo, scale hi, shift lo, shift hi), lane
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long f = blockIdx.y;
    {
        const long long base = f * mod_ld;
        for (int c = warp; c < CH; c += WARPS_PER_BLOCK) {
            float sc[8], sh[8];
            load_param8(scale, base + (c * 32 + lane) * 8, params_bf16, sc);
            load_param8(shift, base + (c * 32 + lane) * 8, params_bf16, sh);
            ln_prm[(c * 4 + 0) * 32 + lane] = make_float4(mul_base + sc[0], mul_base + sc[1], mul_base + sc[2], mul_base + sc[3]);
            ln_prm[(c * 4 + 1) * 32 + lane] = make_float4(mul_base + sc[4], mul_base + sc[5], mul_base + sc[6], mul_base + sc[7]);
            ln_prm[(c * 4 + 2) * 32 + lane] = make_float4(sh[0], sh[1], sh[2], sh[3]);
            ln_prm[(c * 4 + 3) * 32 + lane] = make_float4(sh[4], sh[5], sh[6], sh[7]);
        }
    }
    __syncthreads();
    const long long r0 = f * tpf + (long long)blockIdx.x * LN_FWD_ROWS;
    const long long r1 = min(min(rows, (f + 1) * (long long)tpf), r0 + LN_FWD_ROWS);
    for (long long row = r0 + warp; row < r1; row += WARPS_PER_BLOCK) {
        // the row stays PACKED (bf16) in registers and is unpacked in each of the three passes: 64 instead of 128 data
        // registers at C = 4096
        uint4 raw[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) raw[c] = __ldg(reinterpret_cast<const uint4*>(X + row * ldx + (c * 32 + lane) * 8));
        float s8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            float v[8];
            unpack8_opaque<1>(raw[c], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) s8[i] += v[i];
        }
        const float mean = warp_sum(((s8[0] + s8[1]) + (s8[2] + s8[3])) + ((s8[4] + s8[5]) + (s8[6] + s8[7]))) * (1.0f / C);
        float q8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            float v[8];
            unpack8_opaque<2>(raw[c], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q8[i] = fmaf(d, d, q8[i]); }
        }
        const float rstd = rsqrtf(warp_sum(((q8[0] + q8[1]) + (q8[2] + q8[3])) + ((q8[4] + q8[5]) + (q8[6] + q8[7]))) * (1.0f / C) + eps);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const float4 a0 = ln_prm[(c * 4 + 0) * 32 + lane], a1 = ln_prm[(c * 4 + 1) * 32 + lane];
            const float4 b0 = ln_prm[(c * 4 + 2) * 32 + lane], b1 = ln_prm[(c * 4 + 3) * 32 + lane];
            const float sc[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float sh[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float v[8];
            unpack8_opaque<3>(raw[c], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf((v[i] - mean) * rstd, sc[i], sh[i]);
            *reinterpret_cast<uint4*>(Y + row * ldy + (c * 32 + lane) * 8) = pack8(v);
        }
    }
}

// Backward: the row of x and of dy are staged in shared memory with cp.async (all 16-byte requests of a
// row in flight at once, no register pressure for C = 4096), then four passes read shared memory.
constexpr int LN_BWD_WARPS = 4;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ float sum8(const float (&a)[8]) {
    return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}
// CH = C / 256 (0: runtime C, rolled loops).  The passes over the staged row are fully unrolled for the widths the
// model uses and every reduction runs eight independent chains (round 2 ncu: 41 % issue utilisation, 18.75 % occupancy:
// the single-chain loops were latency-bound).
template <int CH>
__global__ void __launch_bounds__(LN_BWD_WARPS * 32, 3) ln_mod_bwd_kernel(
    __nv_bfloat16* __restrict__ dX, long long lddx, const __nv_bfloat16* __restrict__ dXr, long long ldr,
    const __nv_bfloat16* __restrict__ dY, long long lddy, const __nv_bfloat16* __restrict__ X, long long ldx,
    const void* __restrict__ scale, long long mod_ld, int params_bf16, float mul_base, long long rows, int C_rt, int tpf,
    float eps) {
    extern __shared__ uint4 ln_smem[];  // [LN_BWD_WARPS][2][C/8]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * LN_BWD_WARPS + warp;
    if (row >= rows) return;
    const int C = CH ? CH * 256 : C_rt;
    const int nvec = C / 8;
    constexpr int UNR = CH ? (CH < 4 ? CH : 4) : 1;
    uint4* xs = ln_smem + (size_t)warp * 2 * nvec;
    uint4* gs = xs + nvec;
#pragma unroll UNR
    for (int j = lane; j < nvec; j += 32) {
        cp_async16(xs + j, X + row * ldx + j * 8);
        cp_async16(gs + j, dY + row * lddy + j * 8);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    const float invC = 1.0f / (float)C;
    float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll UNR
    for (int j = lane; j < nvec; j += 32) {
        float v[8];
        unpack8(xs[j], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) a8[i] += v[i];
    }
    const float mean = warp_sum(sum8(a8)) * invC;
#pragma unroll
    for (int i = 0; i < 8; ++i) a8[i] = 0.f;
#pragma unroll UNR
    for (int j = lane; j < nvec; j += 32) {
        float v[8];
        unpack8(xs[j], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; a8[i] = fmaf(d, d, a8[i]); }
    }
    const float rstd = rsqrtf(warp_sum(sum8(a8)) * invC + eps);
    const long long f = row / tpf;
    // g = dy * (mul_base + scale); dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
    float g8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, x8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll UNR
    for (int j = lane; j < nvec; j += 32) {
        float sc[8], v[8], g[8];
        load_param8(scale, f * mod_ld + j * 8, params_bf16, sc);
        unpack8(xs[j], v);
        unpack8(gs[j], g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float gg = g[i] * (mul_base + sc[i]);
            g8[i] += gg;
            x8[i] = fmaf(gg, (v[i] - mean) * rstd, x8[i]);
        }
    }
    const float mg = warp_sum(sum8(g8)) * invC, mgx = warp_sum(sum8(x8)) * invC;
#pragma unroll UNR
    for (int j = lane; j < nvec; j += 32) {
        float sc[8], v[8], g[8], r[8];
        load_param8(scale, f * mod_ld + j * 8, params_bf16, sc);
        unpack8(xs[j], v);
        unpack8(gs[j], g);
        if (dXr != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(dXr + row * ldr + j * 8)), r);
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
            r[i] += rstd * (g[i] * (mul_base + sc[i]) - mg - (v[i] - mean) * rstd * mgx);
        *reinterpret_cast<uint4*>(dX + row * lddx + j * 8) = pack8(r);
    }
}

// column reductions over the rows of one frame: dscale[f,c] += sum dY*xhat, dshift[f,c] += sum dY.
// grid (blocks_per_frame, frames); a block owns PARAM_ROWS_PER_BLOCK rows of ONE frame, one warp per row
// (row in registers as in the kernels above), partial sums in shared memory, one global atomic per column.
constexpr int PARAM_ROWS_PER_BLOCK = 64;
template <int CH>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) ln_param_grad_kernel(
    const __nv_bfloat16* __restrict__ dY, long long lddy, const __nv_bfloat16* __restrict__ X, long long ldx,
    float* __restrict__ dscale, float* __restrict__ dshift, long long acc_ld, long long rows, int tpf, float eps) {
    constexpr int C = CH * 256;
    extern __shared__ float acc_s[];  // [2][C]
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) acc_s[i] = 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long f = blockIdx.y;
    const long long r0 = f * tpf + (long long)blockIdx.x * PARAM_ROWS_PER_BLOCK;
    const long long r1 = min(min(rows, (f + 1) * (long long)tpf), r0 + PARAM_ROWS_PER_BLOCK);
    for (long long row = r0 + warp; row < r1; row += WARPS_PER_BLOCK) {
        float v[CH][8];
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            unpack8(__ldg(reinterpret_cast<const uint4*>(X + row * ldx + (c * 32 + lane) * 8)), v[c]);
#pragma unroll
            for (int i = 0; i < 8; ++i) s += v[c][i];
        }
        const float mean = warp_sum(s) * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[c][i] -= mean; q += v[c][i] * v[c][i]; }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            float d[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(dY + row * lddy + (c * 32 + lane) * 8)), d);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                atomicAdd(&acc_s[(c * 32 + lane) * 8 + i], d[i] * v[c][i] * rstd);
                atomicAdd(&acc_s[C + (c * 32 + lane) * 8 + i], d[i]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        if (dscale) atomicAdd(dscale + f * acc_ld + i, acc_s[i]);
        if (dshift) atomicAdd(dshift + f * acc_ld + i, acc_s[C + i]);
    }
}

// ------------------------------------------------------------------------------------ RMSNorm(128) + RoPE
struct RopeCtx {
    float c0, s0, c1, s1;  // lane owns elements d = 4*lane .. 4*lane+3 = pairs (4l,4l+1), (4l+2,4l+3)
};
__device__ __forceinline__ RopeCtx rope_ctx(long long pos, int gh, int gw, float base, int lane) {
    constexpr int D = 128;
    constexpr int d_hw = 2 * (D / 6), d_t = D - 2 * d_hw;
    const int w = (int)(pos % gw);
    const int h = (int)((pos / gw) % gh);
    const int t = (int)(pos / ((long long)gw * gh));
    RopeCtx r;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int d = 4 * lane + 2 * k;
        int p, i2, dim;
        if (d < d_t) { p = t; i2 = d; dim = d_t; }
        else if (d < d_t + d_hw) { p = h; i2 = d - d_t; dim = d_hw; }
        else { p = w; i2 = d - d_t - d_hw; dim = d_hw; }
        const float inv = 1.0f / powf(base, (float)i2 / (float)dim);
        float sn, cs;
        sincosf((float)p * inv, &sn, &cs);
        if (k == 0) { r.c0 = cs; r.s0 = sn; } else { r.c1 = cs; r.s1 = sn; }
    }
    return r;
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) qk_norm_rope_fwd_kernel(
    __nv_bfloat16* __restrict__ Y, long long ldy, const __nv_bfloat16* __restrict__ X, long long ldx,
    const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk, int n_q, int n_k, long long rows,
    long long row_offset, int gh, int gw, int rope, float base, float eps) {
    const long long row = (long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    RopeCtx rc = {1.f, 0.f, 1.f, 0.f};
    if (rope) rc = rope_ctx(row + row_offset, gh, gw, base, lane);
    float wqf[4], wkf[4];
    {
        uint2 a = __ldg(reinterpret_cast<const uint2*>(wq + lane * 4));
        float2 t = unpack_bf16x2(a.x); wqf[0] = t.x; wqf[1] = t.y; t = unpack_bf16x2(a.y); wqf[2] = t.x; wqf[3] = t.y;
        if (n_k > 0) {
            uint2 b = __ldg(reinterpret_cast<const uint2*>(wk + lane * 4));
            t = unpack_bf16x2(b.x); wkf[0] = t.x; wkf[1] = t.y; t = unpack_bf16x2(b.y); wkf[2] = t.x; wkf[3] = t.y;
        } else { wkf[0] = wkf[1] = wkf[2] = wkf[3] = 0.f; }
    }
    const __nv_bfloat16* xr = X + row * ldx;
    __nv_bfloat16* yr = Y + row * ldy;
    // q heads then k heads as two loops: selecting the weight array through a pointer put both arrays in local memory
    // (32-byte stack frame, round-2 SASS); four heads in flight hide the load + shuffle-reduction latency of each
    auto head = [&](int s, const float (&w)[4]) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(xr + s * 128 + lane * 4));
        float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
        const float ss = warp_sum(a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y);
        const float rstd = rsqrtf(ss * (1.0f / 128.0f) + eps);
        const float y0 = a.x * rstd * w[0], y1 = a.y * rstd * w[1], y2 = b.x * rstd * w[2], y3 = b.y * rstd * w[3];
        uint2 o;
        o.x = pack_bf16x2(y0 * rc.c0 - y1 * rc.s0, y1 * rc.c0 + y0 * rc.s0);
        o.y = pack_bf16x2(y2 * rc.c1 - y3 * rc.s1, y3 * rc.c1 + y2 * rc.s1);
        *reinterpret_cast<uint2*>(yr + s * 128 + lane * 4) = o;
    };
#pragma unroll 4
    for (int s = 0; s < n_q; ++s) head(s, wqf);
#pragma unroll 4
    for (int s = n_q; s < n_q + n_k; ++s) head(s, wkf);
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) qk_norm_rope_bwd_kernel(
    __nv_bfloat16* __restrict__ dX, long long lddx, const __nv_bfloat16* __restrict__ dY, long long lddy,
    const __nv_bfloat16* __restrict__ X, long long ldx, const __nv_bfloat16* __restrict__ wq,
    const __nv_bfloat16* __restrict__ wk, float* __restrict__ dwq, float* __restrict__ dwk, int n_q, int n_k,
    long long rows, long long row_offset, int gh, int gw, int rope, float base, float eps) {
    __shared__ float red[2][WARPS_PER_BLOCK][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * WARPS_PER_BLOCK + warp;
    float aq[4] = {0.f, 0.f, 0.f, 0.f}, ak[4] = {0.f, 0.f, 0.f, 0.f};
    if (row < rows) {
        RopeCtx rc = {1.f, 0.f, 1.f, 0.f};
        if (rope) rc = rope_ctx(row + row_offset, gh, gw, base, lane);
        float wqf[4], wkf[4];
        uint2 a = __ldg(reinterpret_cast<const uint2*>(wq + lane * 4));
        float2 t = unpack_bf16x2(a.x); wqf[0] = t.x; wqf[1] = t.y; t = unpack_bf16x2(a.y); wqf[2] = t.x; wqf[3] = t.y;
        if (n_k > 0) {
            uint2 b = __ldg(reinterpret_cast<const uint2*>(wk + lane * 4));
            t = unpack_bf16x2(b.x); wkf[0] = t.x; wkf[1] = t.y; t = unpack_bf16x2(b.y); wkf[2] = t.x; wkf[3] = t.y;
        } else { wkf[0] = wkf[1] = wkf[2] = wkf[3] = 0.f; }
        auto head = [&](int s, const float (&w)[4], float (&acc)[4]) {
            const uint2 ux = __ldg(reinterpret_cast<const uint2*>(X + row * ldx + s * 128 + lane * 4));
            const uint2 ug = __ldg(reinterpret_cast<const uint2*>(dY + row * lddy + s * 128 + lane * 4));
            float2 xa = unpack_bf16x2(ux.x), xb = unpack_bf16x2(ux.y);
            float2 ga = unpack_bf16x2(ug.x), gb = unpack_bf16x2(ug.y);
            const float ss = warp_sum(xa.x * xa.x + xa.y * xa.y + xb.x * xb.x + xb.y * xb.y);
            const float rstd = rsqrtf(ss * (1.0f / 128.0f) + eps);
            // transpose of the rotation
            float d[4] = {ga.x * rc.c0 + ga.y * rc.s0, ga.y * rc.c0 - ga.x * rc.s0,
                          gb.x * rc.c1 + gb.y * rc.s1, gb.y * rc.c1 - gb.x * rc.s1};
            const float xh[4] = {xa.x * rstd, xa.y * rstd, xb.x * rstd, xb.y * rstd};
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[i] += d[i] * xh[i]; d[i] *= w[i]; dot += d[i] * xh[i]; }
            dot = warp_sum(dot) * (1.0f / 128.0f);
            uint2 o;
            o.x = pack_bf16x2(rstd * (d[0] - xh[0] * dot), rstd * (d[1] - xh[1] * dot));
            o.y = pack_bf16x2(rstd * (d[2] - xh[2] * dot), rstd * (d[3] - xh[3] * dot));
            *reinterpret_cast<uint2*>(dX + row * lddx + s * 128 + lane * 4) = o;
        };
        // q heads then k heads (compile-time arrays: no local-memory pointer select), four heads in flight
#pragma unroll 4
        for (int s = 0; s < n_q; ++s) head(s, wqf, aq);
#pragma unroll 4
        for (int s = n_q; s < n_q + n_k; ++s) head(s, wkf, ak);
    }
    if (dwq == nullptr && dwk == nullptr) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) { red[0][warp][lane * 4 + i] = aq[i]; red[1][warp][lane * 4 + i] = ak[i]; }
    __syncthreads();
    if (threadIdx.x < 128) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS_PER_BLOCK; ++w) { a += red[0][w][threadIdx.x]; b += red[1][w][threadIdx.x]; }
        if (dwq) atomicAdd(dwq + threadIdx.x, a);
        if (dwk && n_k > 0) atomicAdd(dwk + threadIdx.x, b);
    }
}

// ------------------------------------------------------------------------------------ gate multiply
__global__ void __launch_bounds__(256) gate_mul_kernel(__nv_bfloat16* __restrict__ dY, long long lddy,
                                                       const __nv_bfloat16* __restrict__ dX, long long lddx,
                                                       const float* __restrict__ gate, long long ldg, long long rows,
                                                       int C, int tpf) {
    const int vec_per_row = C / 8;
    const long long total = rows * vec_per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / vec_per_row;
        const int c = (int)(i - row * vec_per_row) * 8;
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dX + row * lddx + c)), v);
        if (gate != nullptr) {
            const float4* g = reinterpret_cast<const float4*>(gate + (row / tpf) * ldg + c);
            const float4 a = __ldg(g), b = __ldg(g + 1);
            v[0] *= a.x; v[1] *= a.y; v[2] *= a.z; v[3] *= a.w; v[4] *= b.x; v[5] *= b.y; v[6] *= b.z; v[7] *= b.w;
        }
        *reinterpret_cast<uint4*>(dY + row * lddy + c) = pack8(v);
    }
}

// dst[i, :] = src[idx[i], :] for rows of `vecs` 16-byte vectors (block-sparse attention: token order <-> block-major order)
__global__ void __launch_bounds__(256) gather_rows_kernel(uint4* __restrict__ dst, long long ldd_v, const uint4* __restrict__ src,
                                                          long long lds_v, const long long* __restrict__ idx, long long rows,
                                                          int vecs) {
    const long long total = rows * vecs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / vecs;
        const int c = (int)(i - r * vecs);
        dst[r * ldd_v + c] = __ldg(src + __ldg(idx + r) * lds_v + c);
    }
}

// out[c] += sum over a slab of rows of A[row, c]: 8 columns per thread (one 128-bit load per row), fp32 atomics per slab
__global__ void __launch_bounds__(256) colsum8_kernel(float* __restrict__ out, const __nv_bfloat16* __restrict__ A, long long lda,
                                                      long long rows, int C, int rows_per_block) {
    const int col = (blockIdx.x * 256 + threadIdx.x) * 8;
    if (col >= C) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(rows, r0 + rows_per_block);
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) {
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(A + r * lda + col)), v);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(out + col + i, s[i]);
}

// acc[f, c] += sum over rows of frame f of A[row, c] * B[row, c]  (B == nullptr -> sum of A)
__global__ void __launch_bounds__(256) colsum_prod_kernel(float* __restrict__ acc, long long acc_ld,
                                                          const __nv_bfloat16* __restrict__ A, long long lda,
                                                          const __nv_bfloat16* __restrict__ B, long long ldb,
                                                          long long rows, int C, int tpf, int rows_per_block) {
    // grid (C/256, ceil(rows / rows_per_block)); thread t owns column blockIdx.x*256 + t
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col >= C) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(rows, r0 + rows_per_block);
    float s = 0.f;
    long long cur_f = r0 / tpf;
    for (long long r = r0; r < r1; ++r) {
        const long long f = r / tpf;
        if (f != cur_f) { atomicAdd(acc + cur_f * acc_ld + col, s); s = 0.f; cur_f = f; }
        const float a = __bfloat162float(A[r * lda + col]);
        s += B ? a * __bfloat162float(B[r * ldb + col]) : a;
    }
    atomicAdd(acc + cur_f * acc_ld + col, s);
}

// ------------------------------------------------------------------------------------ noising / patchify / loss
// token n = (t, h2, w2); P column = c*4 + ph*2 + pw (Conv3d weight order); V column = (ph*2+pw)*16 + c
__global__ void __launch_bounds__(256) noise_patchify_kernel(__nv_bfloat16* __restrict__ P, float* __restrict__ V,
                                                             float* __restrict__ timestep,
                                                             const __nv_bfloat16* __restrict__ cond,
                                                             const __nv_bfloat16* __restrict__ target,
                                                             const __nv_bfloat16* __restrict__ noise,
                                                             const float* __restrict__ sigma_p, int t_cond, int t_tgt,
                                                             int H, int W, float nts) {
    const int H2 = H / 2, W2 = W / 2;
    const int T = t_cond + t_tgt;
    const long long n_tok = (long long)T * H2 * W2;
    const float sigma = sigma_p ? __ldg(sigma_p) : 0.f;
    if (blockIdx.x == 0 && threadIdx.x < T && timestep != nullptr)
        timestep[threadIdx.x] = threadIdx.x < t_cond ? 0.f : __bfloat162float(__float2bfloat16(sigma * nts));
    // one thread per (token, channel): reads a 2x2 patch (two bf16x2 loads), writes 4 consecutive P columns
    const long long total = n_tok * 16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i & 15);
        const long long n = i >> 4;
        const int w2 = (int)(n % W2);
        const int h2 = (int)((n / W2) % H2);
        const int t = (int)(n / ((long long)W2 * H2));
        float x[4];
        if (t < t_cond) {
            const __nv_bfloat16* src = cond + (((long long)c * t_cond + t) * H + 2 * h2) * W + 2 * w2;
            float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src));
            float2 b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + W));
            x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
        } else {
            const int tt = t - t_cond;
            const long long off = (((long long)c * t_tgt + tt) * H + 2 * h2) * W + 2 * w2;
            float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(target + off));
            float2 b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(target + off + W));
            float2 ea = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(noise + off));
            float2 eb = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(noise + off + W));
            const float x0[4] = {a.x, a.y, b.x, b.y}, e[4] = {ea.x, ea.y, eb.x, eb.y};
            const long long nt = n - (long long)t_cond * H2 * W2;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                x[k] = (1.0f - sigma) * x0[k] + sigma * e[k];
                if (V != nullptr) V[nt * 64 + k * 16 + c] = __bfloat162float(__float2bfloat16(e[k] - x0[k]));
            }
        }
        uint2 o;
        o.x = pack_bf16x2(x[0], x[1]);
        o.y = pack_bf16x2(x[2], x[3]);
        *reinterpret_cast<uint2*>(P + n * 64 + c * 4) = o;
    }
}

__global__ void __launch_bounds__(256) unpatchify_kernel(float* __restrict__ latent, const float* __restrict__ tok,
                                                         int T, int H, int W) {
    const int H2 = H / 2, W2 = W / 2;
    const long long total = (long long)16 * T * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        const int h = (int)((i / W) % H);
        const int t = (int)((i / ((long long)W * H)) % T);
        const int c = (int)(i / ((long long)W * H * T));
        const long long n = ((long long)t * H2 + h / 2) * W2 + w / 2;
        latent[i] = tok[n * 64 + ((h & 1) * 2 + (w & 1)) * 16 + c];
    }
}

__global__ void __launch_bounds__(256) mse_kernel(float* __restrict__ loss, __nv_bfloat16* __restrict__ dpred,
                                                  const float* __restrict__ pred, const float* __restrict__ V,
                                                  long long n, float inv_n, float loss_scale) {
    __shared__ float red[8];
    float s = 0.f;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        const float4 p = *reinterpret_cast<const float4*>(pred + i);
        const float4 v = *reinterpret_cast<const float4*>(V + i);
        const float d0 = p.x - v.x, d1 = p.y - v.y, d2 = p.z - v.z, d3 = p.w - v.w;
        s += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        if (dpred != nullptr) {
            const float k = 2.0f * inv_n * loss_scale;
            uint2 o;
            o.x = pack_bf16x2(d0 * k, d1 * k);
            o.y = pack_bf16x2(d2 * k, d3 * k);
            *reinterpret_cast<uint2*>(dpred + i) = o;
        }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        s = red[threadIdx.x];
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
        if (threadIdx.x == 0) atomicAdd(loss, s * inv_n);
    }
}

__global__ void sinusoid_kernel(float* __restrict__ F, const float* __restrict__ ts, int rows, int dim) {
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * half) return;
    const int r = i / half, k = i % half;
    const float freq = expf(-logf(10000.0f) * (float)k / (float)half);
    float sn, cs;
    sincosf(ts[r] * freq, &sn, &cs);
    F[r * dim + k] = cs;
    F[r * dim + half + k] = sn;
}

// ------------------------------------------------------------------------------------ small-M fp32 linear
// Y[R,out] = act(X[R,in]) W^T + b + addend.  X (activated) staged in shared memory; one warp per output column.
__global__ void __launch_bounds__(256) skinny_linear_kernel(float* __restrict__ Y, const float* __restrict__ X,
                                                            const void* __restrict__ W, const void* __restrict__ bias,
                                                            const float* __restrict__ addend, int w_bf16, int R, int in,
                                                            int out, int act) {
    extern __shared__ float xs[];  // [R][in]
    for (int i = threadIdx.x; i < R * in; i += blockDim.x) {
        float x = X[i];
        if (act == 1) x = x / (1.0f + __expf(-x));
        xs[i] = x;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = blockIdx.x * 8 + warp; j < out; j += gridDim.x * 8) {
        float accv[2] = {0.f, 0.f};
        for (int r0 = 0; r0 < R; r0 += 2) {
            float a0 = 0.f, a1 = 0.f;
            for (int k = lane; k < in; k += 32) {
                const float w = w_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(W)[(long long)j * in + k])
                                       : reinterpret_cast<const float*>(W)[(long long)j * in + k];
                a0 = fmaf(xs[r0 * in + k], w, a0);
                if (r0 + 1 < R) a1 = fmaf(xs[(r0 + 1) * in + k], w, a1);
            }
            accv[0] = warp_sum(a0);
            accv[1] = warp_sum(a1);
            if (lane == 0) {
                float b = 0.f;
                if (bias) b = w_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(bias)[j])
                                     : reinterpret_cast<const float*>(bias)[j];
                Y[(long long)r0 * out + j] = accv[0] + b + (addend ? addend[(long long)r0 * out + j] : 0.f);
                if (r0 + 1 < R)
                    Y[(long long)(r0 + 1) * out + j] = accv[1] + b + (addend ? addend[(long long)(r0 + 1) * out + j] : 0.f);
            }
        }
    }
}

// dX[R,in] (+)= (dY[R,out] W[out,in]) * act'(X).  grid over `in` columns: thread owns (r, k) pairs.
__global__ void __launch_bounds__(256) skinny_linear_bwd_kernel(float* __restrict__ dX, const float* __restrict__ dY,
                                                                const float* __restrict__ X, const void* __restrict__ W,
                                                                int w_bf16, int R, int in, int out, int act,
                                                                int accumulate, int out_chunk) {
    // grid (ceil(in/256), out_chunks); each block reduces a slice of `out` and atomically adds
    const int k = blockIdx.x * 256 + threadIdx.x;
    const int o0 = blockIdx.y * out_chunk, o1 = min(out, o0 + out_chunk);
    extern __shared__ float dys[];  // [R][out_chunk]
    for (int i = threadIdx.x; i < R * (o1 - o0); i += blockDim.x) {
        const int r = i / (o1 - o0), o = i % (o1 - o0);
        dys[r * out_chunk + o] = dY[(long long)r * out + o0 + o];
    }
    __syncthreads();
    if (k >= in) return;
    for (int r0 = 0; r0 < R; r0 += 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int o = o0; o < o1; ++o) {
            const float w = w_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(W)[(long long)o * in + k])
                                   : reinterpret_cast<const float*>(W)[(long long)o * in + k];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (r0 + i < R) acc[i] = fmaf(dys[(r0 + i) * out_chunk + (o - o0)], w, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (r0 + i >= R) break;
            float g = acc[i];
            if (act == 1) {
                const float x = X[(long long)(r0 + i) * in + k];
                const float sg = 1.0f / (1.0f + __expf(-x));
                g *= sg * (1.0f + x * (1.0f - sg));
            }
            atomicAdd(dX + (long long)(r0 + i) * in + k, g);
        }
    }
    (void)accumulate;
}


// ------------------------------------------------------------------------------------ standalone SwiGLU
__global__ void __launch_bounds__(256) swiglu_fwd_kernel(__nv_bfloat16* __restrict__ Hout, const __nv_bfloat16* __restrict__ H1,
                                                         const __nv_bfloat16* __restrict__ H3, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float a[8], b[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(H1) + i), a);
        unpack8(__ldg(reinterpret_cast<const uint4*>(H3) + i), b);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = a[k] / (1.0f + __expf(-a[k])) * b[k];
        reinterpret_cast<uint4*>(Hout)[i] = pack8(a);
    }
}
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(__nv_bfloat16* __restrict__ dH1, __nv_bfloat16* __restrict__ dH3,
                                                         const __nv_bfloat16* __restrict__ dH, const __nv_bfloat16* __restrict__ H1,
                                                         const __nv_bfloat16* __restrict__ H3, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float a[8], b[8], g[8], o1[8], o3[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(H1) + i), a);
        unpack8(__ldg(reinterpret_cast<const uint4*>(H3) + i), b);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dH) + i), g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float sig = 1.0f / (1.0f + __expf(-a[k]));
            o1[k] = g[k] * b[k] * sig * (1.0f + a[k] * (1.0f - sig));
            o3[k] = g[k] * a[k] * sig;
        }
        reinterpret_cast<uint4*>(dH1)[i] = pack8(o1);
        reinterpret_cast<uint4*>(dH3)[i] = pack8(o3);
    }
}
// latent f32 [16,T,H,W] -> tokens bf16 [n, 64] in final-layer column order (inverse of unpatchify; used for d(pred))
__global__ void __launch_bounds__(256) latent_to_tokens_kernel(__nv_bfloat16* __restrict__ tok, const float* __restrict__ latent,
                                                               int T, int H, int W, int t0) {
    const int H2 = H / 2, W2 = W / 2;
    const long long total = (long long)(T - t0) * H2 * W2 * 64;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i & 63);
        const long long n = i >> 6;
        const int c = col & 15, pw = (col >> 4) & 1, ph = col >> 5;
        const int w2 = (int)(n % W2), h2 = (int)((n / W2) % H2), t = (int)(n / ((long long)W2 * H2)) + t0;
        tok[i] = __float2bfloat16(latent[(((long long)c * T + t) * H + 2 * h2 + ph) * W + 2 * w2 + pw]);
    }
}

inline int grid_for(long long work_items, int per_block, int cap = 148 * 16) {
    long long g = (work_items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    return (int)(g > cap ? cap : g);
}

// ------------------------------------------------------------------------------------ VAE latent normalisation
// normalize_latents / denormalize_latents of the reference (delta_experiment/scripts/common.py:175-205) on a
// [B, C, T, H, W] latent: out = (x - mean[c]) * inv_std[c]   or   out = x / inv_std[c] + mean[c], with the reference's
// rounding points (every intermediate is rounded to the latent dtype: two roundings per element for bf16).
template <bool BF16IO, bool INVERSE>
__global__ void __launch_bounds__(256) latent_affine_kernel(void* __restrict__ out_, const void* __restrict__ in_,
                                                            const float* __restrict__ mean, const float* __restrict__ inv_std,
                                                            long long n, long long inner, int channels) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int c = (int)((i / inner) % channels);
        const float m = __ldg(mean + c), s = __ldg(inv_std + c);
        if (BF16IO) {
            const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(in_);
            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_);
            const float x = __bfloat162float(in[i]);
            const float t = __bfloat162float(__float2bfloat16_rn(INVERSE ? __fdiv_rn(x, s) : x - m));
            out[i] = __float2bfloat16_rn(INVERSE ? t + m : t * s);
        } else {
            const float x = reinterpret_cast<const float*>(in_)[i];
            reinterpret_cast<float*>(out_)[i] = INVERSE ? __fadd_rn(__fdiv_rn(x, s), m) : __fmul_rn(__fsub_rn(x, m), s);
        }
    }
}

}  // namespace
}  // namespace b200

using namespace b200;

#define LN_DISPATCH(CH_EXPR, KERNEL_CALL)                      \
    switch (CH_EXPR) {                                         \
        case 1: { constexpr int CH = 1; KERNEL_CALL; } break;  \
        case 2: { constexpr int CH = 2; KERNEL_CALL; } break;  \
        case 4: { constexpr int CH = 4; KERNEL_CALL; } break;  \
        case 8: { constexpr int CH = 8; KERNEL_CALL; } break;  \
        case 16: { constexpr int CH = 16; KERNEL_CALL; } break; \
        default: b200::set_last_error("ln_mod: C=%d unsupported (need C/256 in {1,2,4,8,16})", C); return B200TTA_EINVAL; \
    }

extern "C" int b200tta_ln_mod_fwd(void* Y, int64_t ldy, const void* X, int64_t ldx, const void* scale,
                                  const void* shift, int64_t mod_ld, int32_t params_bf16, float mul_base, int64_t rows,
                                  int32_t C, int32_t tokens_per_frame, float eps, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(rows > 0 && C % 256 == 0 && ldy % 8 == 0 && ldx % 8 == 0 && aligned16(Y) && aligned16(X) &&
                     aligned16(scale) && aligned16(shift) && mod_ld % 8 == 0 && tokens_per_frame > 0,
                 "ln_mod_fwd: bad shape/alignment (rows=%lld C=%d)", (long long)rows, C);
    // affine form (mod_ld == 0): one parameter set for every row -> the whole input is one "frame"
    const long long tpf = mod_ld == 0 ? rows : tokens_per_frame;
    const dim3 grid((unsigned)((tpf + LN_FWD_ROWS - 1) / LN_FWD_ROWS), (unsigned)((rows + tpf - 1) / tpf));
    const size_t smem = 2 * (size_t)C * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    LN_DISPATCH(C / 256, (ln_mod_fwd_kernel<CH><<<grid, WARPS_PER_BLOCK * 32, smem, st>>>(
                             (__nv_bfloat16*)Y, ldy, (const __nv_bfloat16*)X, ldx, scale, shift, mod_ld, params_bf16,
                             mul_base, rows, (int)tpf, eps)));
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_ln_mod_bwd(void* dX, int64_t lddx, const void* dX_resid, int64_t ldr, const void* dY,
                                  int64_t lddy, const void* X, int64_t ldx, const void* scale, int64_t mod_ld,
                                  int32_t params_bf16, float mul_base, float* dscale_acc, float* dshift_acc,
                                  int64_t acc_ld, int64_t rows, int32_t C, int32_t tokens_per_frame, float eps,
                                  b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(rows > 0 && C % 256 == 0 && lddx % 8 == 0 && lddy % 8 == 0 && ldx % 8 == 0 && aligned16(dX) &&
                     aligned16(dY) && aligned16(X) && aligned16(scale) && mod_ld % 8 == 0 && tokens_per_frame > 0 &&
                     (!dX_resid || (aligned16(dX_resid) && ldr % 8 == 0)),
                 "ln_mod_bwd: bad shape/alignment (rows=%lld C=%d)", (long long)rows, C);
    cudaStream_t st = (cudaStream_t)stream;
    {
        const size_t smem = (size_t)LN_BWD_WARPS * 2 * C * sizeof(__nv_bfloat16);
        static bool attr = false;
        if (!attr) {
            B200_CUDA(cudaFuncSetAttribute(ln_mod_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            B200_CUDA(cudaFuncSetAttribute(ln_mod_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            B200_CUDA(cudaFuncSetAttribute(ln_mod_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            attr = true;
        }
        B200_REQUIRE(smem <= 128 * 1024, "ln_mod_bwd: C=%d too wide", C);
        const int grid_b = (int)((rows + LN_BWD_WARPS - 1) / LN_BWD_WARPS);
#define LN_BWD_LAUNCH(CHV)                                                                                         \
    ln_mod_bwd_kernel<CHV><<<grid_b, LN_BWD_WARPS * 32, smem, st>>>(                                                \
        (__nv_bfloat16*)dX, lddx, (const __nv_bfloat16*)dX_resid, ldr, (const __nv_bfloat16*)dY, lddy,             \
        (const __nv_bfloat16*)X, ldx, scale, mod_ld, params_bf16, mul_base, rows, C, tokens_per_frame, eps)
        if (C == 4096) LN_BWD_LAUNCH(16);          // the 13.6 B width
        else if (C == 512) LN_BWD_LAUNCH(2);       // the tiny config
        else LN_BWD_LAUNCH(0);
#undef LN_BWD_LAUNCH
    }
    B200_LAUNCHED();
    if (dscale_acc || dshift_acc) {
        // affine form (mod_ld == 0): every row belongs to "frame" 0 -> treat the whole input as one frame
        const long long tpf = mod_ld == 0 ? rows : tokens_per_frame;
        const long long frames = (rows + tpf - 1) / tpf;
        dim3 g((unsigned)((tpf + PARAM_ROWS_PER_BLOCK - 1) / PARAM_ROWS_PER_BLOCK), (unsigned)frames);
        const size_t smem = 2 * (size_t)C * sizeof(float);
        LN_DISPATCH(C / 256, (ln_param_grad_kernel<CH><<<g, WARPS_PER_BLOCK * 32, smem, st>>>(
                                 (const __nv_bfloat16*)dY, lddy, (const __nv_bfloat16*)X, ldx, dscale_acc, dshift_acc,
                                 acc_ld, rows, (int)tpf, eps)));
        B200_LAUNCHED();
    }
    return B200TTA_OK;
}

extern "C" int b200tta_qk_rmsnorm_rope_fwd(void* Y, int64_t ldy, const void* X, int64_t ldx, const void* wq,
                                           const void* wk, int32_t n_q_slots, int32_t n_k_slots, int64_t rows,
                                           int64_t row_offset, int32_t grid_h, int32_t grid_w, int32_t rope,
                                           float rope_base, float eps, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(rows > 0 && n_q_slots >= 0 && n_k_slots >= 0 && n_q_slots + n_k_slots > 0 && ldy % 4 == 0 &&
                     ldx % 4 == 0 && wq && (n_k_slots == 0 || wk) && (!rope || (grid_h > 0 && grid_w > 0)),
                 "qk_rmsnorm_rope_fwd: bad arguments");
    const int grid = (int)((rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    qk_norm_rope_fwd_kernel<<<grid, WARPS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)Y, ldy, (const __nv_bfloat16*)X, ldx, (const __nv_bfloat16*)wq, (const __nv_bfloat16*)wk,
        n_q_slots, n_k_slots, rows, row_offset, grid_h, grid_w, rope, rope_base, eps);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_qk_rmsnorm_rope_bwd(void* dX, int64_t lddx, const void* dY, int64_t lddy, const void* X,
                                           int64_t ldx, const void* wq, const void* wk, float* dwq_acc, float* dwk_acc,
                                           int32_t n_q_slots, int32_t n_k_slots, int64_t rows, int64_t row_offset,
                                           int32_t grid_h, int32_t grid_w, int32_t rope, float rope_base, float eps,
                                           b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(rows > 0 && n_q_slots + n_k_slots > 0 && lddx % 4 == 0 && lddy % 4 == 0 && ldx % 4 == 0 && wq &&
                     (n_k_slots == 0 || wk),
                 "qk_rmsnorm_rope_bwd: bad arguments");
    const int grid = (int)((rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    qk_norm_rope_bwd_kernel<<<grid, WARPS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)dX, lddx, (const __nv_bfloat16*)dY, lddy, (const __nv_bfloat16*)X, ldx,
        (const __nv_bfloat16*)wq, (const __nv_bfloat16*)wk, dwq_acc, dwk_acc, n_q_slots, n_k_slots, rows, row_offset,
        grid_h, grid_w, rope, rope_base, eps);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_gate_mul(void* dY, int64_t lddy, const void* dX, int64_t lddx, const float* gate, int64_t ldg,
                                const void* branch, int64_t ldb, float* dgate_acc, int64_t acc_ld, int64_t rows,
                                int32_t C, int32_t tokens_per_frame, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(rows > 0 && C % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0 && aligned16(dY) && aligned16(dX) &&
                     (!gate || (aligned16(gate) && ldg % 4 == 0)) && tokens_per_frame > 0,
                 "gate_mul: bad shape/alignment");
    cudaStream_t st = (cudaStream_t)stream;
    gate_mul_kernel<<<grid_for(rows * (C / 8), 256), 256, 0, st>>>((__nv_bfloat16*)dY, lddy, (const __nv_bfloat16*)dX,
                                                                   lddx, gate, ldg, rows, C, tokens_per_frame);
    B200_LAUNCHED();
    if (dgate_acc) {
        B200_REQUIRE(branch != nullptr, "gate_mul: dgate_acc needs the branch output");
        const int rpb = 128;
        dim3 g((C + 255) / 256, (unsigned)((rows + rpb - 1) / rpb));
        colsum_prod_kernel<<<g, 256, 0, st>>>(dgate_acc, acc_ld, (const __nv_bfloat16*)dX, lddx,
                                              (const __nv_bfloat16*)branch, ldb, rows, C, tokens_per_frame, rpb);
        B200_LAUNCHED();
    }
    return B200TTA_OK;
}

extern "C" int b200tta_gather_rows(void* dst, int64_t ldd, const void* src, int64_t lds, const int64_t* idx, int64_t rows,
                                   int32_t row_elems, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(dst && src && idx && rows > 0 && row_elems > 0 && row_elems % 8 == 0 && ldd % 8 == 0 && lds % 8 == 0 &&
                     aligned16(dst) && aligned16(src),
                 "gather_rows: bf16 rows must be 16-byte aligned with a length that is a multiple of 8");
    gather_rows_kernel<<<grid_for(rows * (row_elems / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        (uint4*)dst, ldd / 8, (const uint4*)src, lds / 8, (const long long*)idx, rows, row_elems / 8);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_colsum(float* out, const void* A, int64_t lda, int64_t rows, int32_t C, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(out && A && rows > 0 && C > 0, "colsum: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    B200_CUDA(cudaMemsetAsync(out, 0, (size_t)C * sizeof(float), st));
    if (C % 8 == 0 && lda % 8 == 0 && aligned16(A)) {
        // enough row slabs to fill the chip whatever the width: ~8 blocks per SM
        const int xb = (C / 8 + 255) / 256;
        int rpb = (int)((rows * xb + 148 * 8 - 1) / (148 * 8));
        rpb = rpb < 32 ? 32 : rpb;
        dim3 g((unsigned)xb, (unsigned)((rows + rpb - 1) / rpb));
        colsum8_kernel<<<g, 256, 0, st>>>(out, (const __nv_bfloat16*)A, lda, rows, C, rpb);
    } else {
        const int rpb = 256;
        dim3 g((C + 255) / 256, (unsigned)((rows + rpb - 1) / rpb));
        colsum_prod_kernel<<<g, 256, 0, st>>>(out, 0, (const __nv_bfloat16*)A, lda, nullptr, 0, rows, C,
                                              (int)(rows > 2147483647ll ? 2147483647ll : rows), rpb);
    }
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_noise_patchify(void* P, float* V, float* timestep, const void* cond, const void* target,
                                      const void* noise, const float* sigma, int32_t t_cond, int32_t t_tgt, int32_t H,
                                      int32_t W, float num_train_timesteps, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(P && t_cond >= 0 && t_tgt >= 0 && t_cond + t_tgt > 0 && t_cond + t_tgt <= 256 && H % 2 == 0 &&
                     W % 2 == 0 && (t_cond == 0 || cond) && (t_tgt == 0 || (target && noise && sigma)),
                 "noise_patchify: bad arguments");
    const long long n_tok = (long long)(t_cond + t_tgt) * (H / 2) * (W / 2);
    noise_patchify_kernel<<<grid_for(n_tok * 16, 256), 256, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)P, V, timestep, (const __nv_bfloat16*)cond, (const __nv_bfloat16*)target,
        (const __nv_bfloat16*)noise, sigma, t_cond, t_tgt, H, W, num_train_timesteps);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_patchify(void* P, const void* latent, int32_t T, int32_t H, int32_t W, b200tta_stream_t stream) {
    return b200tta_noise_patchify(P, nullptr, nullptr, latent, nullptr, nullptr, nullptr, T, 0, H, W, 0.f, stream);
}

extern "C" int b200tta_unpatchify(float* latent, const float* tokens, int32_t T, int32_t H, int32_t W,
                                  b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(latent && tokens && T > 0 && H % 2 == 0 && W % 2 == 0, "unpatchify: bad arguments");
    unpatchify_kernel<<<grid_for((long long)16 * T * H * W, 256), 256, 0, (cudaStream_t)stream>>>(latent, tokens, T, H, W);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_mse_fwd_bwd(float* loss, void* dpred, const float* pred, const float* V, int64_t n,
                                   float loss_scale, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(loss && pred && V && n > 0 && n % 4 == 0 && aligned16(pred) && aligned16(V), "mse_fwd_bwd: bad arguments");
    mse_kernel<<<grid_for(n / 4, 256, 148 * 4), 256, 0, (cudaStream_t)stream>>>(loss, (__nv_bfloat16*)dpred, pred, V, n,
                                                                               1.0f / (float)n, loss_scale);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_timestep_sinusoid(float* F, const float* timestep, int32_t rows, int32_t dim,
                                         b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(F && timestep && rows > 0 && dim > 0 && dim % 2 == 0, "timestep_sinusoid: bad arguments");
    const int n = rows * dim / 2;
    sinusoid_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(F, timestep, rows, dim);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_skinny_linear(float* Y, const float* X, const void* W, const void* bias, const float* addend,
                                     int32_t w_bf16, int32_t R, int32_t in_features, int32_t out_features, int32_t act,
                                     b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(Y && X && W && R > 0 && R <= 64 && in_features > 0 && out_features > 0, "skinny_linear: bad arguments");
    const size_t smem = (size_t)R * in_features * sizeof(float);
    B200_REQUIRE(smem <= 200 * 1024, "skinny_linear: R*in too large for shared memory");
    static bool attr = false;
    if (!attr) {
        B200_CUDA(cudaFuncSetAttribute(skinny_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    int grid = (out_features + 7) / 8;
    if (grid > 148 * 4) grid = 148 * 4;
    skinny_linear_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(Y, X, W, bias, addend, w_bf16, R, in_features,
                                                                    out_features, act);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_skinny_linear_bwd(float* dX, const float* dY, const float* X, const void* W, int32_t w_bf16,
                                         int32_t R, int32_t in_features, int32_t out_features, int32_t act,
                                         int32_t accumulate, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(dX && dY && W && R > 0 && R <= 64 && (act == 0 || X), "skinny_linear_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) B200_CUDA(cudaMemsetAsync(dX, 0, (size_t)R * in_features * sizeof(float), st));
    const int out_chunk = 256;
    dim3 g((in_features + 255) / 256, (out_features + out_chunk - 1) / out_chunk);
    const size_t smem = (size_t)R * out_chunk * sizeof(float);
    static bool attr = false;
    if (!attr) {
        B200_CUDA(cudaFuncSetAttribute(skinny_linear_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr = true;
    }
    skinny_linear_bwd_kernel<<<g, 256, smem, st>>>(dX, dY, X, W, w_bf16, R, in_features, out_features, act, accumulate,
                                                   out_chunk);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_swiglu_fwd(void* Hout, const void* H1, const void* H3, int64_t n, b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(Hout && H1 && H3 && n > 0 && n % 8 == 0 && aligned16(Hout) && aligned16(H1) && aligned16(H3), "swiglu_fwd: bad arguments");
    swiglu_fwd_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)Hout, (const __nv_bfloat16*)H1,
                                                                              (const __nv_bfloat16*)H3, n / 8);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_swiglu_bwd(void* dH1, void* dH3, const void* dH, const void* H1, const void* H3, int64_t n,
                                  b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(dH1 && dH3 && dH && H1 && H3 && n > 0 && n % 8 == 0 && aligned16(dH1) && aligned16(dH3) && aligned16(dH) &&
                     aligned16(H1) && aligned16(H3), "swiglu_bwd: bad arguments");
    swiglu_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)dH1, (__nv_bfloat16*)dH3, (const __nv_bfloat16*)dH, (const __nv_bfloat16*)H1, (const __nv_bfloat16*)H3, n / 8);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_latent_to_tokens(void* tokens, const float* latent, int32_t T, int32_t H, int32_t W, int32_t t_begin,
                                        b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(tokens && latent && T > 0 && t_begin >= 0 && t_begin < T && H % 2 == 0 && W % 2 == 0, "latent_to_tokens: bad arguments");
    const long long total = (long long)(T - t_begin) * (H / 2) * (W / 2) * 64;
    latent_to_tokens_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)tokens, latent, T, H, W, t_begin);
    B200_LAUNCHED();
    return B200TTA_OK;
}

extern "C" int b200tta_latent_affine(void* out, const void* in, const float* mean, const float* inv_std, int64_t n,
                                     int64_t inner, int32_t channels, int32_t is_bf16, int32_t inverse,
                                     b200tta_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    B200_REQUIRE(out && in && mean && inv_std && n > 0 && inner > 0 && channels > 0 && n % (inner * channels) == 0,
                 "latent_affine: n=%lld must be batch x channels=%d x inner=%lld", (long long)n, channels, (long long)inner);
    const int grid = grid_for(n, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (is_bf16) {
        if (inverse) latent_affine_kernel<true, true><<<grid, 256, 0, st>>>(out, in, mean, inv_std, n, inner, channels);
        else latent_affine_kernel<true, false><<<grid, 256, 0, st>>>(out, in, mean, inv_std, n, inner, channels);
    } else {
        if (inverse) latent_affine_kernel<false, true><<<grid, 256, 0, st>>>(out, in, mean, inv_std, n, inner, channels);
        else latent_affine_kernel<false, false><<<grid, 256, 0, st>>>(out, in, mean, inv_std, n, inner, channels);
    }
    B200_LAUNCHED();
    return B200TTA_OK;
}
