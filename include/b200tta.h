/* b200tta -- C ABI of the B200-native TTA inner step (libb200tta.so).
 *
 * Drop-in boundary for ONE path of FifthEpoch/longcat-video-tta: the flow-matching
 * test-time-adaptation step through the LongCat-Video DiT.  The reference has no
 * FFI of its own (it is pure PyTorch calling cuBLAS / flash-attn / ATen); each entry
 * point below replaces the library call site named in its comment (file:line are
 * into the reference tree) and is what a maintainer would bind from Python
 * (ctypes stub in INTEGRATION.md).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch tensors);
 *     kernels never allocate, workspaces are caller-provided;
 *   - bf16 tensors are row-major with an explicit leading dimension in ELEMENTS;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return 0 or a negative code; b200tta_last_error() gives the thread-local text;
 *   - there is no CPU fallback: on a device that is not sm_100 every compute entry
 *     point returns B200TTA_EARCH.
 */
#ifndef B200TTA_H
#define B200TTA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200TTA_OK 0
#define B200TTA_EINVAL (-1) /* bad shape / alignment / argument */
#define B200TTA_EARCH (-2)  /* device is not sm_100 (B200) */
#define B200TTA_ECUDA (-3)  /* CUDA launch / runtime failure */

typedef void* b200tta_stream_t; /* cudaStream_t */

int b200tta_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py reports it as gpu_launches) */
int64_t b200tta_launch_count(void);
const char* b200tta_last_error(void);
/* Returns 0 on an sm_100 device after running a tiny tcgen05 GEMM against a CUDA-core
 * reference on device; B200TTA_EARCH elsewhere. */
int b200tta_selfcheck(void);

/* ------------------------------------------------------------------------------------
 * Multi-segment bf16 GEMM on tcgen05/TMEM/TMA -- the engine under the LoRA-fused linears.
 *     D[M,N] = epilogue( sum_s A_s[M,K_s] * op(B_s) )
 * A_s is row-major [M,K_s].  B_s is row-major [N,K_s] (b_mn_major = 0, "x @ W^T", forward)
 * or row-major [K_s,N] (b_mn_major = 1, "dy @ W", backward: no transposed weight copy).
 * b_hi (optional, b_mn_major = 0 only): the upper half of every 256-wide N tile is taken from
 * b_hi instead of b -- rows [n/2, n/2+128) of each -- used to co-tile w1/w3 for SwiGLU.
 * Replaces: cuBLASLt via nn.Linear inside the upstream blocks (module names evidenced at
 * lora_experiment/scripts/run_lora_tta.py:146-168).
 * ---------------------------------------------------------------------------------- */
typedef struct {
    const void* a;
    int64_t lda;
    const void* b;
    int64_t ldb;
    const void* b_hi;
    int64_t k;
    int32_t b_mn_major;
    int32_t a_mn_major; /* A stored [K, M] (M contiguous); needs b_mn_major.  The weight-gradient form dW = dY^T X with
                           dY [tokens, out], X [tokens, in] read in place (full-model TTA, run_full_tta.py:95-219) */
} b200tta_gemm_seg;

enum {
    B200TTA_EPI_STORE = 0,      /* d = bf16(acc + bias) */
    B200TTA_EPI_STORE_F32 = 1,  /* d = f32(acc + bias) */
    B200TTA_EPI_GELU = 2,       /* d = bf16(gelu_tanh(acc + bias)) */
    B200TTA_EPI_GATE_RESID = 3, /* d = bf16(resid + gate[row / tokens_per_frame] * (acc + bias)); gate NULL -> 1;
                                   d2 (optional) = bf16(acc + bias) */
    B200TTA_EPI_SWIGLU = 4,     /* N tile = [h1 | h3]: d[:, n/2] = bf16(silu(h1) * h3); d2, d3 (optional) = h1, h3 */
    B200TTA_EPI_SWIGLU_BWD = 5, /* acc = dh: d = bf16(dh * h3 * silu'(h1)), d2 = bf16(dh * silu(h1)); aux1 = h1, aux2 = h3 */
    B200TTA_EPI_MUL_GATE_IN = 6, /* d = bf16(acc + bias) (reserved) */
    B200TTA_EPI_GEGLU = 7       /* SWIGLU's tile layout with d = bf16(gelu_tanh(h1) * h3): the UMT5 feed-forward
                                   (gated-gelu: wi_0, wi_1; text encoder of common.py:228-255) */
};

typedef struct {
    int32_t mode;
    int32_t tokens_per_frame;
    void* d;
    int64_t ldd;
    void* d2;
    int64_t ldd2;
    void* d3;
    int64_t ldd3;
    const void* bias; /* [N], bf16 unless bias_is_f32 */
    int32_t bias_is_f32;
    int32_t reserved;
    const void* resid; /* bf16 [M, ldr] */
    int64_t ldr;
    const float* gate; /* f32 [frames, ldg] */
    int64_t ldg;
    const void* aux1;
    int64_t ldaux1;
    const void* aux2;
    int64_t ldaux2;
} b200tta_gemm_epi;

int b200tta_gemm(int64_t M, int64_t N, const b200tta_gemm_seg* segs, int32_t nseg, const b200tta_gemm_epi* epi,
                 b200tta_stream_t stream);

/* ------------------------------------------------------------------------------------
 * LoRA-fused linear.  Replaces LoRALinear.forward (run_lora_tta.py:255-260) and the builtin
 * hook (run_lora_tta.py:175-181): y = x W^T + b + scale * (x A^T) B^T with the frozen base GEMM
 * and the low-rank product sharing one accumulator (the rank-r product is one extra K block).
 *   X [n_tok, in] bf16, W [out, in] bf16, bias [out] bf16 or NULL,
 *   A [r, in] bf16 (lora_down.weight), B [out, r] bf16 (lora_up.weight), r % 8 == 0, r <= 64,
 *   XA [n_tok, r] bf16 workspace/output: bf16(scale * X A^T), kept for the backward.
 * r == 0 (A == NULL) degenerates to the plain frozen linear.  `epi` selects the fused epilogue
 * (its d/ldd are the output).
 * ---------------------------------------------------------------------------------- */
int b200tta_lora_linear_fwd(const void* X, int64_t ldx, const void* W, const void* W_hi, const void* A,
                            const void* B, void* XA, int64_t n_tok, int64_t in_features, int64_t out_features,
                            int32_t r, float scale, const b200tta_gemm_epi* epi, b200tta_stream_t stream);

/* Backward of the above for frozen W (autograd through nn.Linear with requires_grad=False
 * weights, run_lora_tta.py:512,814-815): dX = dY W + (scale dY B) A ; no dW is ever formed.
 *   dA_acc [in, r] f32 += X^T (scale dY B)   (the TRANSPOSE of dA: token reductions produce [features, r]);
 *   dB_acc [out, r] f32 += dY^T XA
 *   U [n_tok, r] bf16 workspace.  dX may be NULL (adapter grads only, e.g. kv_linear on text).
 *   `epi` is the epilogue of the dX GEMM (mode STORE, or SWIGLU_BWD for w2). */
int b200tta_lora_linear_bwd(const void* dY, int64_t lddy, const void* X, int64_t ldx, const void* W, const void* A,
                            const void* B, const void* XA, void* U, float* dA_acc, float* dB_acc, int64_t n_tok,
                            int64_t in_features, int64_t out_features, int32_t r, float scale,
                            const b200tta_gemm_epi* epi, b200tta_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Flash-style attention, D = 128, bf16, tcgen05.  Replaces flash_attn_func /
 * flash_attn_varlen_func (FlashAttention-2 enabled at delta_experiment/scripts/common.py:71-74).
 * q/k/v/o are [tokens, heads, 128] views with a token stride in elements (so q, k, v may alias
 * one fused qkv buffer).  Segments implement the clean-context / noised split of the upstream
 * Attention: segment s = queries [q_begin, q_end) attend keys [0, kv_len).
 * lse [heads, n_q] f32 (natural log) is written for the backward.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int32_t q_begin, q_end, kv_len;
} b200tta_attn_seg;

int b200tta_attn_fwd(void* O, int64_t ldo, float* LSE, const void* Q, int64_t ldq, const void* K, int64_t ldk,
                     const void* V, int64_t ldv, int32_t n_q, int32_t n_kv, int32_t heads, float softmax_scale,
                     const b200tta_attn_seg* segs, int32_t n_seg, b200tta_stream_t stream);

/* Backward (recomputes S = QK^T from Q, K, LSE; FlashAttention-2 backward replaced).
 * delta [heads, n_q] f32 workspace (rowsum(dO * O)).  dQ/dK/dV bf16 with their own strides. */
int b200tta_attn_bwd(void* dQ, int64_t lddq, void* dK, int64_t lddk, void* dV, int64_t lddv, const void* dO,
                     int64_t lddo, const void* O, int64_t ldo, const float* LSE, float* delta, const void* Q,
                     int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, int32_t n_q, int32_t n_kv,
                     int32_t heads, float softmax_scale, const b200tta_attn_seg* segs, int32_t n_seg,
                     b200tta_stream_t stream);

/* Same contract as b200tta_attn_bwd, computed by ONE kernel per (key block, head) that forms exactly the five necessary
 * products per (query tile, key block) pair (S^T, dP^T, dV, dK and the dQ partial product) instead of the seven of the
 * dq + dkv pair.  dQ partial tiles leave through fp32 reduce-adds into `dq_acc` -- caller-provided workspace,
 * [heads, n_q, 128] f32 (n_q * heads * 128 floats), contents irrelevant on entry (zeroed here) -- and are converted to bf16 `dQ` at the end.
 * The summation order of the partial tiles is not fixed run to run (differences at fp32 rounding level).
 * Replaces the same reference call as b200tta_attn_bwd: loss.backward() through flash-attn's backward
 * (lora_experiment/scripts/run_lora_tta.py:512). */
int b200tta_attn_bwd_fused(void* dQ, int64_t lddq, void* dK, int64_t lddk, void* dV, int64_t lddv, const void* dO,
                           int64_t lddo, const void* O, int64_t ldo, const float* LSE, float* delta, const void* Q,
                           int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, int32_t n_q, int32_t n_kv,
                           int32_t heads, float softmax_scale, const b200tta_attn_seg* segs, int32_t n_seg,
                           float* dq_acc, b200tta_stream_t stream);

/* Block-sparse self-attention for the 720p refinement stage (upstream `enable_bsa` / `bsa_params`, flags at
 * delta_experiment/scripts/common.py:71-74; the upstream kernel is not vendored -- SURVEY App. A.9 -- so the block
 * semantics below are OUR definition).  Tokens are in BLOCK-MAJOR order (the 128 tokens of one 3-D chunk are
 * consecutive; longcat_video_tta_b200.bsa builds the permutation and the lists), n_tok % 128 == 0, one segment.
 *   q_off [heads * n_blk + 1], q_idx : for (head h, query block i) the ascending key-block indices it attends are
 *                                      q_idx[q_off[h * n_blk + i] .. q_off[h * n_blk + i + 1])
 *   k_off / k_idx                    : the transpose (for each key block, the query blocks attending it), backward only.
 * Everything else as b200tta_attn_fwd / _bwd: the same kernels walk the lists instead of all blocks. */
int b200tta_attn_bsa_fwd(void* O, int64_t ldo, float* LSE, const void* Q, int64_t ldq, const void* K, int64_t ldk,
                         const void* V, int64_t ldv, int32_t n_tok, int32_t heads, float softmax_scale,
                         const int32_t* q_off, const int32_t* q_idx, b200tta_stream_t stream);
int b200tta_attn_bsa_bwd(void* dQ, int64_t lddq, void* dK, int64_t lddk, void* dV, int64_t lddv, const void* dO,
                         int64_t lddo, const void* O, int64_t ldo, const float* LSE, float* delta, const void* Q,
                         int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, int32_t n_tok,
                         int32_t heads, float softmax_scale, const int32_t* q_off, const int32_t* q_idx,
                         const int32_t* k_off, const int32_t* k_idx, b200tta_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Fused elementwise / row-reduction kernels (128-bit HBM access, warp-shuffle reductions).
 * Replace the unfused ATen chains of the upstream block.
 * ---------------------------------------------------------------------------------- */

/* Replaces: the LayerNorm + AdaLN-modulate chains of the upstream block, whose inputs the reference exposes at
 * delta_experiment/scripts/run_delta_a.py:196-207 (block call with the per-frame timestep embedding t), run_film_tta.py:129-144
 * (hook on adaLN_modulation's output) and run_norm_tune_tta.py:78-96 (pre_crs_attn_norm weight / bias as trainables).
 * y = LN(x) * (mul_base + scale[f]) + shift[f]   (eps 1e-6, fp32 math, bf16 in/out)
 * modulated form: scale/shift f32 rows of the adaLN output, f = row / tokens_per_frame, mul_base = 1
 * affine form (pre_crs_attn_norm): scale = weight, shift = bias (bf16, params_bf16 = 1), mod_ld = 0, mul_base = 0 */
int b200tta_ln_mod_fwd(void* Y, int64_t ldy, const void* X, int64_t ldx, const void* scale, const void* shift,
                       int64_t mod_ld, int32_t params_bf16, float mul_base, int64_t rows, int32_t C,
                       int32_t tokens_per_frame, float eps, b200tta_stream_t stream);
/* dX = (dX_resid or 0) + LN_bwd(dY * (mul_base + scale)); optional f32 accumulators:
 * dscale_acc[f, c] += sum_rows dY * xhat ; dshift_acc[f, c] += sum_rows dY  (f = 0 for the affine form) */
int b200tta_ln_mod_bwd(void* dX, int64_t lddx, const void* dX_resid, int64_t ldr, const void* dY, int64_t lddy,
                       const void* X, int64_t ldx, const void* scale, int64_t mod_ld, int32_t params_bf16,
                       float mul_base, float* dscale_acc, float* dshift_acc, int64_t acc_ld, int64_t rows, int32_t C,
                       int32_t tokens_per_frame, float eps, b200tta_stream_t stream);

/* Replaces: q_norm / k_norm (RMSNorm over head_dim; their weights are the qk_norm trainables of
 * run_norm_tune_tta.py:85-96) followed by the upstream 3-D rotary embedding.
 * per-head RMSNorm(128) * weight then 3-D RoPE (interleaved pairs; axis dims 44/42/42 for D=128).
 * X [rows, n_slots, 128] bf16 with row stride ldx; slots [0, n_q_slots) use wq, [n_q_slots, n_q_slots+n_k_slots) use wk.
 * rope = 0 skips the rotation (cross-attention).  grid (T, Hh, Ww): row -> (t, h, w) row-major with row_offset. */
int b200tta_qk_rmsnorm_rope_fwd(void* Y, int64_t ldy, const void* X, int64_t ldx, const void* wq, const void* wk,
                                int32_t n_q_slots, int32_t n_k_slots, int64_t rows, int64_t row_offset, int32_t grid_h,
                                int32_t grid_w, int32_t rope, float rope_base, float eps, b200tta_stream_t stream);
/* dX = RMSNorm_bwd(RoPE^T(dY)); optional f32 dwq_acc/dwk_acc [128] (norm-tune). */
int b200tta_qk_rmsnorm_rope_bwd(void* dX, int64_t lddx, const void* dY, int64_t lddy, const void* X, int64_t ldx,
                                const void* wq, const void* wk, float* dwq_acc, float* dwk_acc, int32_t n_q_slots,
                                int32_t n_k_slots, int64_t rows, int64_t row_offset, int32_t grid_h, int32_t grid_w,
                                int32_t rope, float rope_base, float eps, b200tta_stream_t stream);

/* Replaces: autograd through the gated residual x + gate * branch of the upstream block (the gate is a chunk of the
 * adaLN output that FiLM corrects, run_film_tta.py:129-144); the forward gate-residual is a GEMM epilogue.
 * dY = bf16(gate[f] * dX) (backward of x + gate * branch wrt branch); gate NULL -> copy.
 * optional dgate_acc[f, c] += sum_rows dX * branch (branch bf16, needed by FiLM / delta adapters). */
int b200tta_gate_mul(void* dY, int64_t lddy, const void* dX, int64_t lddx, const float* gate, int64_t ldg,
                     const void* branch, int64_t ldb, float* dgate_acc, int64_t acc_ld, int64_t rows, int32_t C,
                     int32_t tokens_per_frame, b200tta_stream_t stream);

/* Noising + concat + patchify (delta_experiment/scripts/common.py:458-470 fused with the im2col of PatchEmbed3D,
 * x_embedder call at run_delta_a.py:154):
 *   P[n, c*4 + ph*2 + pw] = bf16(cond) for context frames, bf16((1-sigma) x0 + sigma eps) for target frames
 *   V[n_tgt, (ph*2+pw)*16 + c] = f32(bf16(eps - x0))   velocity target in final-layer column order
 *   timestep[t] = 0 | f32(bf16(sigma * 1000))
 * cond [16,Tc,H,W], target/noise [16,Tt,H,W] bf16; sigma: device f32 scalar. */
int b200tta_noise_patchify(void* P, float* V, float* timestep, const void* cond, const void* target, const void* noise,
                           const float* sigma, int32_t t_cond, int32_t t_tgt, int32_t H, int32_t W,
                           float num_train_timesteps, b200tta_stream_t stream);
/* plain patchify of a [16,T,H,W] bf16 latent (forward-only API; x_embedder, run_delta_a.py:154) */
int b200tta_patchify(void* P, const void* latent, int32_t T, int32_t H, int32_t W, b200tta_stream_t stream);
/* tokens [T*H/2*W/2, 64] f32 (final-layer order) -> latent [16,T,H,W] f32   (dit.unpatchify, run_delta_a.py:214) */
int b200tta_unpatchify(float* latent, const float* tokens, int32_t T, int32_t H, int32_t W, b200tta_stream_t stream);

/* autograd through unpatchify (loss.backward(), run_lora_tta.py:512).
 * d(pred): latent-layout f32 gradient [16,T,H,W] -> token-layout bf16 [(T - t_begin)*H/2*W/2, 64] (frames >= t_begin) */
int b200tta_latent_to_tokens(void* tokens, const float* latent, int32_t T, int32_t H, int32_t W, int32_t t_begin,
                             b200tta_stream_t stream);

/* Replaces: silu(w1 x) * w3 x of the upstream FeedForward (ffn.w1 / w2 / w3, run_lora_tta.py:163-168).
 * standalone SwiGLU (used when w1/w3 carry LoRA adapters and cannot be co-tiled): h = silu(h1) * h3 over n bf16
 * elements, and its backward (dh1, dh3) from dh. */
int b200tta_swiglu_fwd(void* Hout, const void* H1, const void* H3, int64_t n, b200tta_stream_t stream);
int b200tta_swiglu_bwd(void* dH1, void* dH3, const void* dH, const void* H1, const void* H3, int64_t n,
                       b200tta_stream_t stream);

/* Replaces: F.mse_loss and its backward (delta_experiment/scripts/common.py:485-488; run_lora_tta.py:512).
 * loss += mean((pred - V)^2) over n elements; dpred = bf16(2 (pred - V) / n * loss_scale). loss must be zeroed. */
int b200tta_mse_fwd_bwd(float* loss, void* dpred, const float* pred, const float* V, int64_t n, float loss_scale,
                        b200tta_stream_t stream);

/* Replaces: t_embedder's frequency embedding (called at run_delta_a.py:163-165).
 * sinusoidal timestep features: F[t, :] = cat(cos(t w_i), sin(t w_i)), i < dim/2 (f32) */
int b200tta_timestep_sinusoid(float* F, const float* timestep, int32_t rows, int32_t dim, b200tta_stream_t stream);
/* small-M fp32 linear with bf16/f32 weights: Y[R,out] = act(X[R,in]) W^T + b (+ addend); act: 0 none, 1 SiLU on input.
 * Used for t_embedder and adaLN_modulation (fp32 islands); R <= 64. */
int b200tta_skinny_linear(float* Y, const float* X, const void* W, const void* bias, const float* addend,
                          int32_t w_bf16, int32_t R, int32_t in_features, int32_t out_features, int32_t act,
                          b200tta_stream_t stream);
/* backward wrt the input: dX[R,in] = (dY[R,out] W) * act'(X) ; accumulate = 1 adds into dX.  This is the path the delta-A / delta-B
 * gradient takes back to the timestep embedding (offset added at run_delta_a.py:168, run_delta_b.py:194-211). */
int b200tta_skinny_linear_bwd(float* dX, const float* dY, const float* X, const void* W, int32_t w_bf16, int32_t R,
                              int32_t in_features, int32_t out_features, int32_t act, int32_t accumulate,
                              b200tta_stream_t stream);

/* LoRA side products (token-dimension reductions), bf16 in, fp32 accumulate: lora_down(x) of LoRALinear.forward
 * (run_lora_tta.py:258) and the autograd products that give lora_down / lora_up their .grad (run_lora_tta.py:512):
 *   down:  T[n_tok, r] = bf16(scale * X[n_tok, k] * Wd^T)   Wd [r, k]      (wd_t = 0)
 *                      = bf16(scale * X[n_tok, k] * Wd)     Wd [k, r]      (wd_t = 1)
 *   grad:  G[m, r] (f32) += P[n_tok, m]^T * Q[n_tok, r] */
int b200tta_lora_down(void* T, int64_t ldt, const void* X, int64_t ldx, const void* Wd, int32_t wd_t, int64_t n_tok,
                      int64_t k, int32_t r, float scale, b200tta_stream_t stream);
int b200tta_lora_grad(float* G, const void* P, int64_t ldp, const void* Q, int64_t ldq, int64_t n_tok, int64_t m,
                      int32_t r, b200tta_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Multi-tensor clip + AdamW over the adapter parameters.  Replaces
 * torch.nn.utils.clip_grad_norm_ (run_lora_tta.py:513; per tensor at run_delta_b.py:386-388)
 * and torch.optim.AdamW.step (run_lora_tta.py:462-468,514).
 * A tensor list is an array of device-resident descriptors.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    void* param;      /* bf16 (is_bf16) or f32 */
    float* master;    /* optional f32 master copy of a bf16 param (NULL: update the param in place) */
    float* grad;      /* f32 gradient accumulator of this step */
    void* exp_avg;    /* f32 if master != NULL or the param is f32, else bf16 */
    void* exp_avg_sq; /* same dtype as exp_avg */
    int64_t numel;
    int32_t is_bf16;
    int32_t t_rows; /* > 0: param is [t_rows, t_cols] and grad is stored TRANSPOSED as [t_cols, t_rows] */
    int32_t t_cols; /*      (LoRA "down" gradients, see b200tta_lora_linear_bwd) */
    int32_t reserved;
} b200tta_tensor_desc;

/* sumsq[i] = sum(grad_i^2) for every tensor (f32 [n]); sumsq must be zeroed by the caller. */
int b200tta_mt_sumsq(const b200tta_tensor_desc* descs_dev, int32_t n, int64_t max_numel, float* sumsq,
                     b200tta_stream_t stream);
/* coef[i]: global mode (per_tensor = 0): coef[*] = min(1, max_norm / (grad_scale * sqrt(sum_i sumsq[i]) + 1e-6));
 * per-tensor mode: from sumsq[i] alone.  total_norm_out (optional) receives the global norm.
 * grad_scale is the factor mt_adamw applies to every gradient (1/world after an all-reduce sum). */
int b200tta_clip_coef(float* coef, float* total_norm_out, const float* sumsq, int32_t n, float max_norm,
                      int32_t per_tensor, float grad_scale, b200tta_stream_t stream);
/* decoupled-decay Adam; g = grad * grad_scale * coef[i]; bf16 tensors without a master copy are updated with the
 * op-by-op bf16 rounding of torch's foreach path when faithful_bf16 = 1, fp32 math otherwise. */
int b200tta_mt_adamw(const b200tta_tensor_desc* descs_dev, int32_t n, int64_t max_numel, const float* coef,
                     float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                     int32_t faithful_bf16, b200tta_stream_t stream);

/* torch.optim.SGD(momentum=0, weight_decay=wd) over the same descriptor table (full-model TTA default optimizer,
 * lora_experiment/scripts/run_full_tta.py:132-138): g = coef * grad_scale * grad + wd * p ; p -= lr * g. */
int b200tta_mt_sgd(const b200tta_tensor_desc* descs_dev, int32_t n, int64_t max_numel, const float* coef, float grad_scale,
                   float lr, float weight_decay, b200tta_stream_t stream);

/* dst[i, :] = src[idx[i], :]: bf16 rows of `row_elems` elements (multiple of 8), row strides ldd / lds in elements,
 * idx int64 [rows].  The block-sparse path's token-order <-> block-major permutations (our layout, see bsa.py). */
int b200tta_gather_rows(void* dst, int64_t ldd, const void* src, int64_t lds, const int64_t* idx, int64_t rows,
                        int32_t row_elems, b200tta_stream_t stream);

/* out[c] = sum over rows of A[row, c]  (bf16 in, f32 out; bias gradients of full-model TTA). */
int b200tta_colsum(float* out, const void* A, int64_t lda, int64_t rows, int32_t C, b200tta_stream_t stream);

/* VAE latent normalisation, the reference-owned part of SURVEY 8f row 4's VAE half: normalize_latents / denormalize_latents
 * (delta_experiment/scripts/common.py:175-205) on a dense [B, channels, T, H, W] latent (inner = T * H * W, n = total
 * elements).  inverse = 0: out = (x - mean[c]) * inv_std[c]; inverse = 1: out = x / inv_std[c] + mean[c]; bf16 (is_bf16)
 * or f32 in and out, every intermediate rounded to that dtype as torch does.  mean / inv_std f32 [channels], already
 * rounded to the latent dtype by the caller.  The VAE network itself (upstream AutoencoderKLWan) is not part of this
 * library. */
int b200tta_latent_affine(void* out, const void* in, const float* mean, const float* inv_std, int64_t n, int64_t inner,
                          int32_t channels, int32_t is_bf16, int32_t inverse, b200tta_stream_t stream);

/* ---- text encoder (text half): encode_prompt, delta_experiment/scripts/common.py:228-255, which calls
 * transformers' UMT5EncoderModel.  Linear layers go through b200tta_gemm (B200TTA_EPI_GEGLU for wi_0 | wi_1,
 * B200TTA_EPI_GATE_RESID with gate NULL for the residual adds), the embedding lookup through b200tta_gather_rows. */

/* UMT5LayerNorm: Y = bf16(w * bf16(X * rsqrt(mean(X^2) + eps))), no mean subtraction, no bias.  X, Y bf16 [rows, C]
 * (row strides ldx / ldy elements), w bf16 [C]; C a multiple of 8, at most 4096. */
int b200tta_t5_rmsnorm(void* Y, int64_t ldy, const void* X, int64_t ldx, const void* w, int64_t rows, int32_t C, float eps,
                       b200tta_stream_t stream);

/* UMT5Attention core: O = softmax(Q K^T + rel_bias[head, key - query] + key mask) V per (batch item, head); head_dim 64,
 * NO 1/sqrt(d) scaling.  Q/K/V/O bf16 [batch * n_tok, heads * 64] views (row strides in elements);
 * rel_bias f32 [heads, 2 * n_tok - 1] indexed by (key - query) + n_tok - 1 (the bucketed table of the layer expanded per
 * distance); key_valid int32 [batch, n_tok] (0 = padded key, gets finfo(float32).min like transformers' extended mask)
 * or NULL. */
int b200tta_t5_attn(void* O, int64_t ldo, const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V,
                    int64_t ldv, const float* rel_bias, const int32_t* key_valid, int32_t n_tok, int32_t heads,
                    int32_t batch, b200tta_stream_t stream);


#ifdef __cplusplus
}
#endif
#endif /* B200TTA_H */
