"""Thin torch -> C-ABI wrappers.  PyTorch is used for device memory and streams only; every function here
launches hand-written kernels from libb200tta.so on the current CUDA stream and raises if the library
reports an error (no fallback)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import os

import torch

from . import _lib
from ._lib import (AttnSeg, GemmEpi, GemmSeg, TensorDesc, EPI_GATE_RESID, EPI_GELU, EPI_STORE, EPI_STORE_F32,
                   EPI_SWIGLU, EPI_SWIGLU_BWD, EPI_GEGLU, check)

BF16, F32 = torch.bfloat16, torch.float32

LAUNCHES = 0  # number of kernel-launching C-ABI calls made through this module (bench.py reports it)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _ld(t: Optional[torch.Tensor]) -> int:
    if t is None:
        return 0
    assert t.dim() >= 2 and t.stride(-1) == 1, f"need a row-major 2-D view, got strides {t.stride()}"
    return t.stride(-2)


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.B200TTAError(f"{name}: tensor is on {t.device}; b200tta operators run on a B200 only (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")


def _call(name: str, *args):
    global LAUNCHES
    LAUNCHES += 1
    check(getattr(_lib.load(), name)(*args), name)


def kernel_launches() -> int:
    """CUDA kernels launched by libb200tta.so so far in this process."""
    return int(_lib.load().b200tta_launch_count())


def selfcheck():
    check(_lib.load().b200tta_selfcheck(), "b200tta_selfcheck")


# ------------------------------------------------------------------------------------------------ epilogues
def epi(mode: int, d: torch.Tensor, *, bias: Optional[torch.Tensor] = None, d2=None, d3=None, resid=None, gate=None,
        tokens_per_frame: int = 1, aux1=None, aux2=None) -> GemmEpi:
    e = GemmEpi()
    e.mode = mode
    e.tokens_per_frame = tokens_per_frame
    e.d, e.ldd = _p(d), _ld(d)
    e.d2, e.ldd2 = _p(d2), _ld(d2)
    e.d3, e.ldd3 = _p(d3), _ld(d3)
    e.bias = _p(bias)
    e.bias_is_f32 = int(bias is not None and bias.dtype == F32)
    e.resid, e.ldr = _p(resid), _ld(resid)
    e.gate, e.ldg = _p(gate), _ld(gate)
    e.aux1, e.ldaux1 = _p(aux1), _ld(aux1)
    e.aux2, e.ldaux2 = _p(aux2), _ld(aux2)
    e._keep = (d, bias, d2, d3, resid, gate, aux1, aux2)
    return e


def gemm(M: int, N: int, segs: Sequence[Tuple[torch.Tensor, torch.Tensor, int, bool, Optional[torch.Tensor]]],
         e: GemmEpi):
    """segs: (A [M,K], B ([N,K] or [K,N] if mn_major), K, mn_major, B_hi or None[, a_mn_major: A given as [K,M]])"""
    arr = (GemmSeg * len(segs))()
    for i, seg in enumerate(segs):
        a, b, k, mn, b_hi = seg[:5]
        arr[i].a_mn_major = int(len(seg) > 5 and bool(seg[5]))
        _req(a, BF16, "gemm A")
        _req(b, BF16, "gemm B")
        arr[i].a, arr[i].lda = _p(a), _ld(a)
        arr[i].b, arr[i].ldb = _p(b), _ld(b)
        arr[i].b_hi = _p(b_hi)
        arr[i].k = k
        arr[i].b_mn_major = int(mn)
    _call("b200tta_gemm", M, N, arr, len(segs), C.byref(e), _stream())


def lora_linear_fwd(x, W, e: GemmEpi, *, W_hi=None, A=None, B=None, XA=None, scale: float = 1.0):
    _req(x, BF16, "lora_linear_fwd x")
    _req(W, BF16, "lora_linear_fwd W")
    n_tok, in_f = x.shape
    out_f = W.shape[0]
    r = 0 if A is None else A.shape[0]
    _call("b200tta_lora_linear_fwd", _p(x), _ld(x), _p(W), _p(W_hi), _p(A), _p(B), _p(XA), n_tok, in_f, out_f, r,
          float(scale), C.byref(e), _stream())


def lora_linear_bwd(dy, W, e: Optional[GemmEpi], *, x=None, A=None, B=None, XA=None, U=None, dA_acc=None, dB_acc=None,
                    scale: float = 1.0):
    """dX = dY W (+ LoRA) through `e` (None: adapter grads only); dA_acc is [in, r] (transposed), dB_acc [out, r]."""
    _req(dy, BF16, "lora_linear_bwd dy")
    n_tok, out_f = dy.shape
    in_f = W.shape[1]
    r = 0 if A is None else A.shape[0]
    _call("b200tta_lora_linear_bwd", _p(dy), _ld(dy), _p(x), _ld(x) if x is not None else 0, _p(W), _p(A), _p(B),
          _p(XA), _p(U), _p(dA_acc), _p(dB_acc), n_tok, in_f, out_f, r, float(scale),
          C.byref(e) if e is not None else None, _stream())


# ------------------------------------------------------------------------------------------------ attention
def _segs(segs: Sequence[Tuple[int, int, int]]):
    arr = (AttnSeg * len(segs))()
    for i, (a, b, c) in enumerate(segs):
        arr[i].q_begin, arr[i].q_end, arr[i].kv_len = a, b, c
    return arr


def attn_fwd(q, k, v, o, lse, segs, softmax_scale: float):
    """q/k/v/o: [tokens, heads, 128] bf16 views (token stride arbitrary); lse [heads, n_q] f32."""
    n_q, H, D = q.shape
    assert D == 128 and q.stride(2) == 1 and q.stride(1) == 128
    _call("b200tta_attn_fwd", _p(o), o.stride(0), _p(lse), _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0),
          n_q, k.shape[0], H, float(softmax_scale), _segs(segs), len(segs), _stream())


_DQ_ACC = {}


def _dq_accumulator(n_q: int, width: int, device) -> torch.Tensor:
    """fp32 [heads, n_q, 128] reduce-add target of the fused backward (caller-provided workspace of the C ABI); one per
    device, grown on demand and kept (613 MB at the headline shape)."""
    key = (device.type, device.index)
    t = _DQ_ACC.get(key)
    if t is None or t.numel() < n_q * width:
        _DQ_ACC[key] = t = torch.empty(n_q * width, dtype=torch.float32, device=device)
    return t


def attn_bwd(dq, dk, dv, do, o, lse, delta, q, k, v, segs, softmax_scale: float, fused: Optional[bool] = None):
    """fused (default): ONE kernel forms dK, dV and the dQ partial products (five products per tile pair, dQ through
    asynchronous fp32 TMA reduce-adds into a [heads, n_q, 128] workspace kept per device);
    split (B200TTA_ATTN_BWD=split or fused=False): dq_kernel + dkv_kernel (seven products per tile pair, no atomics, no
    workspace).  Both are parity-tested at every size.  Stand-alone they tie (52.3 vs 53.5 ms at the headline shape, both
    bound by shared-memory bandwidth); inside the power-capped step the fused kernel's two fewer tensor-core
    products per tile pair buy clock for everything else: 5.32-5.33 s vs 5.37-5.38 s per step, alternating on one box
    (profiles/r2_bench_attn_bwd_split_vs_fused.json)."""
    n_q, H, D = q.shape
    if fused is None:
        fused = os.environ.get("B200TTA_ATTN_BWD", "fused") != "split"
    if fused:
        acc = _dq_accumulator(n_q, H * D, q.device)
        _call("b200tta_attn_bwd_fused", _p(dq), dq.stride(0), _p(dk), dk.stride(0), _p(dv), dv.stride(0), _p(do), do.stride(0),
              _p(o), o.stride(0), _p(lse), _p(delta), _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), n_q,
              k.shape[0], H, float(softmax_scale), _segs(segs), len(segs), _p(acc), _stream())
        return
    _call("b200tta_attn_bwd", _p(dq), dq.stride(0), _p(dk), dk.stride(0), _p(dv), dv.stride(0), _p(do), do.stride(0),
          _p(o), o.stride(0), _p(lse), _p(delta), _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), n_q,
          k.shape[0], H, float(softmax_scale), _segs(segs), len(segs), _stream())


def attn_bsa_fwd(q, k, v, o, lse, q_off, q_idx, softmax_scale: float):
    """block-sparse forward: tokens in block-major order, q_off/q_idx int32 CSR lists per (head, query block) -- see bsa.py"""
    n, H, D = q.shape
    _call("b200tta_attn_bsa_fwd", _p(o), o.stride(0), _p(lse), _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0),
          n, H, float(softmax_scale), _p(q_off), _p(q_idx), _stream())


def attn_bsa_bwd(dq, dk, dv, do, o, lse, delta, q, k, v, q_off, q_idx, k_off, k_idx, softmax_scale: float):
    n, H, D = q.shape
    _call("b200tta_attn_bsa_bwd", _p(dq), dq.stride(0), _p(dk), dk.stride(0), _p(dv), dv.stride(0), _p(do), do.stride(0),
          _p(o), o.stride(0), _p(lse), _p(delta), _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), n, H,
          float(softmax_scale), _p(q_off), _p(q_idx), _p(k_off), _p(k_idx), _stream())


# ------------------------------------------------------------------------------------------------ elementwise
def ln_mod_fwd(y, x, scale, shift, *, tokens_per_frame: int, affine: bool = False, eps: float = 1e-6):
    """modulated: scale/shift f32 [frames, C] views (row stride = adaLN row); affine: weight/bias [C]."""
    rows, Cdim = x.shape
    mod_ld = 0 if affine else scale.stride(0)
    _call("b200tta_ln_mod_fwd", _p(y), _ld(y), _p(x), _ld(x), _p(scale), _p(shift), mod_ld,
          int(scale.dtype == BF16), 0.0 if affine else 1.0, rows, Cdim, tokens_per_frame if not affine else rows,
          float(eps), _stream())


def ln_mod_bwd(dx, dy, x, scale, *, dx_resid=None, tokens_per_frame: int, affine: bool = False, dscale_acc=None,
               dshift_acc=None, eps: float = 1e-6):
    rows, Cdim = x.shape
    mod_ld = 0 if affine else scale.stride(0)
    acc = dscale_acc if dscale_acc is not None else dshift_acc
    acc_ld = 0 if acc is None else (acc.stride(0) if acc.dim() == 2 else 0)
    _call("b200tta_ln_mod_bwd", _p(dx), _ld(dx), _p(dx_resid), _ld(dx_resid), _p(dy), _ld(dy), _p(x), _ld(x),
          _p(scale), mod_ld, int(scale.dtype == BF16), 0.0 if affine else 1.0, _p(dscale_acc), _p(dshift_acc), acc_ld,
          rows, Cdim, tokens_per_frame if not affine else rows, float(eps), _stream())


def qk_rmsnorm_rope_fwd(y, x, wq, wk, n_q_slots: int, n_k_slots: int, *, grid_hw=(1, 1), row_offset: int = 0,
                        rope: bool = True, rope_base: float = 10000.0, eps: float = 1e-6):
    rows = x.shape[0]
    _call("b200tta_qk_rmsnorm_rope_fwd", _p(y), y.stride(0), _p(x), x.stride(0), _p(wq), _p(wk), n_q_slots, n_k_slots,
          rows, row_offset, grid_hw[0], grid_hw[1], int(rope), float(rope_base), float(eps), _stream())


def qk_rmsnorm_rope_bwd(dx, dy, x, wq, wk, n_q_slots: int, n_k_slots: int, *, grid_hw=(1, 1), row_offset: int = 0,
                        rope: bool = True, rope_base: float = 10000.0, eps: float = 1e-6, dwq_acc=None, dwk_acc=None):
    rows = x.shape[0]
    _call("b200tta_qk_rmsnorm_rope_bwd", _p(dx), dx.stride(0), _p(dy), dy.stride(0), _p(x), x.stride(0), _p(wq),
          _p(wk), _p(dwq_acc), _p(dwk_acc), n_q_slots, n_k_slots, rows, row_offset, grid_hw[0], grid_hw[1], int(rope),
          float(rope_base), float(eps), _stream())


def gate_mul(dy, dx, gate, *, tokens_per_frame: int, branch=None, dgate_acc=None):
    rows, Cdim = dx.shape
    _call("b200tta_gate_mul", _p(dy), _ld(dy), _p(dx), _ld(dx), _p(gate), _ld(gate), _p(branch), _ld(branch),
          _p(dgate_acc), _ld(dgate_acc), rows, Cdim, tokens_per_frame, _stream())


def noise_patchify(P, V, timestep, cond, target, noise, sigma, *, num_train_timesteps: float = 1000.0):
    t_cond = 0 if cond is None else cond.shape[1]
    t_tgt = 0 if target is None else target.shape[1]
    H, W = (cond if cond is not None else target).shape[2:]
    _call("b200tta_noise_patchify", _p(P), _p(V), _p(timestep), _p(cond), _p(target), _p(noise), _p(sigma), t_cond,
          t_tgt, H, W, float(num_train_timesteps), _stream())


def unpatchify(latent, tokens, T: int, H: int, W: int):
    _call("b200tta_unpatchify", _p(latent), _p(tokens), T, H, W, _stream())


def latent_to_tokens(tokens, latent, T: int, H: int, W: int, t_begin: int = 0):
    _call("b200tta_latent_to_tokens", _p(tokens), _p(latent), T, H, W, t_begin, _stream())


def swiglu_fwd(h, h1, h3):
    _call("b200tta_swiglu_fwd", _p(h), _p(h1), _p(h3), h.numel(), _stream())


def swiglu_bwd(dh1, dh3, dh, h1, h3):
    _call("b200tta_swiglu_bwd", _p(dh1), _p(dh3), _p(dh), _p(h1), _p(h3), dh.numel(), _stream())


def mse_fwd_bwd(loss, dpred, pred, V, *, loss_scale: float = 1.0):
    _call("b200tta_mse_fwd_bwd", _p(loss), _p(dpred), _p(pred), _p(V), pred.numel(), float(loss_scale), _stream())


def timestep_sinusoid(F, timestep):
    rows, dim = F.shape
    _call("b200tta_timestep_sinusoid", _p(F), _p(timestep), rows, dim, _stream())


def skinny_linear(y, x, W, bias=None, *, act: int = 0, addend=None):
    R, in_f = x.shape
    out_f = W.shape[0]
    _req(x, F32, "skinny_linear x")
    if bias is not None and bias.dtype != W.dtype:
        bias = bias.to(W.dtype)
    _call("b200tta_skinny_linear", _p(y), _p(x), _p(W), _p(bias), _p(addend), int(W.dtype == BF16), R, in_f, out_f, act,
          _stream())


def skinny_linear_bwd(dx, dy, x, W, *, act: int = 0, accumulate: bool = False):
    R, out_f = dy.shape
    in_f = W.shape[1]
    _call("b200tta_skinny_linear_bwd", _p(dx), _p(dy), _p(x), _p(W), int(W.dtype == BF16), R, in_f, out_f, act,
          int(accumulate), _stream())


def lora_down(T, x, Wd, *, transposed: bool = False, scale: float = 1.0):
    n_tok, k = x.shape
    r = Wd.shape[1] if transposed else Wd.shape[0]
    _call("b200tta_lora_down", _p(T), _ld(T), _p(x), _ld(x), _p(Wd), int(transposed), n_tok, k, r, float(scale), _stream())


def lora_grad(G, P, Q):
    n_tok, m = P.shape
    r = Q.shape[1]
    _call("b200tta_lora_grad", _p(G), _p(P), _ld(P), _p(Q), _ld(Q), n_tok, m, r, _stream())


def gather_rows(dst, src, idx):
    """dst[i] = src[idx[i]] for [rows, ...] bf16 tensors whose trailing dims are dense (row stride arbitrary); idx int64"""
    rows = dst.shape[0]
    row_elems = dst[0].numel()
    _req(dst, BF16, "gather_rows dst")
    _req(src, BF16, "gather_rows src")
    _call("b200tta_gather_rows", _p(dst), dst.stride(0), _p(src), src.stride(0), _p(idx), rows, row_elems, _stream())


def colsum(out, a):
    """out [C] f32 = column sums of a [rows, C] bf16 (row stride arbitrary)"""
    rows, Cdim = a.shape
    _call("b200tta_colsum", _p(out), _p(a), _ld(a), rows, Cdim, _stream())


# ------------------------------------------------------------------------------------------------ optimizer
class TensorList:
    """Device-resident descriptor table over the adapter tensors (one multi-tensor launch covers them all)."""

    def __init__(self, entries: List[dict], device):
        arr = (TensorDesc * len(entries))()
        self.max_numel = 0
        self._keep = entries
        for i, e in enumerate(entries):
            p = e["param"]
            arr[i].param = _p(p)
            arr[i].master = _p(e.get("master"))
            arr[i].grad = _p(e["grad"])
            arr[i].exp_avg = _p(e.get("exp_avg"))
            arr[i].exp_avg_sq = _p(e.get("exp_avg_sq"))
            arr[i].numel = p.numel()
            arr[i].is_bf16 = int(p.dtype == BF16)
            tr = e.get("grad_transposed", False)
            arr[i].t_rows = p.shape[0] if tr else 0
            arr[i].t_cols = p.shape[1] if tr else 0
            self.max_numel = max(self.max_numel, p.numel())
        raw = bytes(arr)
        self.n = len(entries)
        self.dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        self.sumsq = torch.zeros(self.n, dtype=F32, device=device)
        self.coef = torch.ones(self.n, dtype=F32, device=device)
        self.total_norm = torch.zeros(1, dtype=F32, device=device)

    def clip_coef(self, max_norm: float, per_tensor: bool = False, grad_scale: float = 1.0):
        self.sumsq.zero_()
        _call("b200tta_mt_sumsq", _p(self.dev), self.n, self.max_numel, _p(self.sumsq), _stream())
        _call("b200tta_clip_coef", _p(self.coef), _p(self.total_norm), _p(self.sumsq), self.n, float(max_norm),
              int(per_tensor), float(grad_scale), _stream())

    def sgd(self, *, lr, weight_decay=0.01, grad_scale: float = 1.0, use_coef: bool = True):
        _call("b200tta_mt_sgd", _p(self.dev), self.n, self.max_numel, _p(self.coef) if use_coef else None, float(grad_scale),
              float(lr), float(weight_decay), _stream())

    def adamw(self, *, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, step: int, grad_scale: float = 1.0,
              use_coef: bool = True, faithful_bf16: bool = False):
        _call("b200tta_mt_adamw", _p(self.dev), self.n, self.max_numel, _p(self.coef) if use_coef else None,
              float(grad_scale), float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step),
              int(faithful_bf16), _stream())


# ------------------------------------------------------------------------------------------------ text encoder
def t5_rmsnorm(y, x, w, eps: float):
    """UMT5LayerNorm on bf16 rows: y = bf16(w * bf16(x * rsqrt(mean(x^2) + eps)))"""
    _req(y, BF16, "t5_rmsnorm y")
    _req(x, BF16, "t5_rmsnorm x")
    _req(w, BF16, "t5_rmsnorm w")
    rows, Cdim = x.shape
    _call("b200tta_t5_rmsnorm", _p(y), _ld(y), _p(x), _ld(x), _p(w), rows, Cdim, float(eps), _stream())


def t5_attn(o, q, k, v, rel_bias, key_valid, n_tok: int, heads: int, batch: int):
    """o/q/k/v: bf16 [batch * n_tok, heads * 64] views; rel_bias f32 [heads, 2 n_tok - 1]; key_valid int32 [batch, n_tok]
    or None.  softmax(q k^T + bias + mask) v without the 1/sqrt(d) factor (UMT5Attention)."""
    for t, name in ((o, "o"), (q, "q"), (k, "k"), (v, "v")):
        _req(t, BF16, "t5_attn " + name)
        assert t.shape == (batch * n_tok, heads * 64), (name, tuple(t.shape))
    _req(rel_bias, F32, "t5_attn rel_bias")
    assert rel_bias.is_contiguous() and rel_bias.shape == (heads, 2 * n_tok - 1)
    if key_valid is not None:
        _req(key_valid, torch.int32, "t5_attn key_valid")
        assert key_valid.is_contiguous() and key_valid.shape == (batch, n_tok)
    _call("b200tta_t5_attn", _p(o), _ld(o), _p(q), _ld(q), _p(k), _ld(k), _p(v), _ld(v), _p(rel_bias), _p(key_valid),
          n_tok, heads, batch, _stream())


# ------------------------------------------------------------------------------------------------ VAE latents
def latent_affine(out, x, mean, inv_std, inverse: bool):
    """out = (x - mean[c]) * inv_std[c]  (inverse: x / inv_std[c] + mean[c]) on a dense [B, C, T, H, W] latent, bf16 or f32"""
    if x.dtype not in (BF16, F32) or out.dtype != x.dtype:
        raise TypeError(f"latent_affine: bf16 or f32 latents, got {x.dtype} -> {out.dtype}")
    _req(x, x.dtype, "latent_affine x")
    _req(mean, F32, "latent_affine mean")
    _req(inv_std, F32, "latent_affine inv_std")
    assert x.dim() == 5 and x.is_contiguous() and out.is_contiguous() and out.shape == x.shape
    Cn = x.shape[1]
    assert mean.numel() == Cn and inv_std.numel() == Cn
    _call("b200tta_latent_affine", _p(out), _p(x), _p(mean), _p(inv_std), x.numel(), x[0, 0].numel(), Cn,
          int(x.dtype == BF16), int(bool(inverse)), _stream())
