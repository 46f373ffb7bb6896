"""TTAEngine: runs the DiT forward, the adapter-only backward and the optimizer step with the sm_100a kernels.

Execution plan (per step, batch 1 as in every reference run -- common.py:458-463):
  noise+patchify -> patch-embed GEMM -> [48 x block forward, block inputs kept] -> final layer -> MSE (+ d pred)
  -> final-layer backward -> [48 x (re-run block forward into one reusable workspace, block backward)]
  -> adapter gradients in ONE flat fp32 buffer (the thing a data-parallel run all-reduces)
  -> multi-tensor clip + AdamW.
Frozen weights never get a gradient: every base GEMM backward is dX only (MN-major B operand, no transposed
copies).  Per-block recompute mirrors the reference's ``torch.utils.checkpoint`` policy
(lora_experiment/scripts/run_lora_tta.py:806-811) but lives in a single pre-allocated workspace.

What is saved where (HBM layout, N tokens, C hidden, F ffn):
  xs[L+1][N,C] bf16      block inputs (checkpoints) and the final hidden state
  o_all[L][N,C], lse_all[L][H,N]   self-attention output and log-sum-exp of every block (the recompute pass skips attention)
  ws.*                   one block's intermediates: xm1, qkv, qk(normed+roped), x1, xn, qc, qcn, kvc, kcn, oc,
                         lsec, x2, xm2, h1, h3, h and the gradient temporaries; re-used by every block
  stash (optional)       whatever HBM is left after the above (180 GB per B200) keeps, for as many blocks as fit,
                         x1 / x2 [N,C], qkv [N,3C] and h1,h3 [N,F] of the training forward, so the recompute pass skips
                         the GEMMs that produced them (same values, bit for bit); the w2 GEMM is never re-run
  grad_flat f32          all adapter gradients back to back (LoRA "down" gradients transposed, see b200tta.h)
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops

BF16, F32 = torch.bfloat16, torch.float32


# ---------------------------------------------------------------------------------------------------- geometry
@dataclass(frozen=True)
class Geometry:
    T: int        # latent frames (cond + target)
    Hl: int       # latent height
    Wl: int       # latent width
    n_cond: int   # clean-context latent frames
    M: int        # packed text tokens

    @property
    def gh(self): return self.Hl // 2
    @property
    def gw(self): return self.Wl // 2
    @property
    def tpf(self): return self.gh * self.gw
    @property
    def N(self): return self.T * self.tpf
    @property
    def Nc(self): return self.n_cond * self.tpf
    @property
    def Nn(self): return self.N - self.Nc

    def self_segments(self):
        if self.Nc == 0:
            return [(0, self.N, self.N)]
        return [(0, self.Nc, self.Nc), (self.Nc, self.N, self.N)]


# ---------------------------------------------------------------------------------------------------- adapter sites
def _hooked_lora(module: nn.Module):
    """The builtin-style ``LoRAModule`` attached to a plain linear: ours carry it as ``._b200_lora``; the reference's own
    ``inject_builtin_lora_into_dit`` keeps ``module.org_forward`` and replaces ``module.forward`` by a closure over the
    ``LoRAModule`` (run_lora_tta.py:137-140,173-182) -- found there, so that its injection works on a B200DiT unchanged."""
    lora = getattr(module, "_b200_lora", None)
    if lora is not None:
        return lora
    fwd = module.__dict__.get("forward")
    if fwd is None or not hasattr(module, "org_forward"):
        return None
    for cell in getattr(fwd, "__closure__", None) or ():
        try:
            obj = cell.cell_contents
        except ValueError:
            continue
        if all(hasattr(obj, a) for a in ("lora_down", "lora_up", "multiplier", "alpha_scale")):
            return obj
    return None


class LinearSite:
    """One linear of the block as the kernels see it: frozen W / bias plus an optional rank-r adapter.

    Recognised adapter forms (duck-typed):
      * a wrapper with ``.original`` / ``.lora_down`` / ``.lora_up`` / ``.scaling``  (reference ``LoRALinear``,
        run_lora_tta.py:224-260, or ours);
      * a plain ``nn.Linear`` carrying ``._b200_lora`` or the reference's hooked ``forward`` (builtin-style
        ``LoRAModule``: ``lora_down`` [n_sep*r, in], ``lora_up`` Linear or ``.blocks[i]``, ``multiplier * alpha_scale``;
        run_lora_tta.py:132-140,175-181).
    """

    def __init__(self, module: nn.Module, name: str):
        self.name = name
        self.params: List[nn.Parameter] = []
        self.blocks_up = None
        lin = module
        self.r = 0
        self.scale = 1.0
        down = up = None
        if hasattr(module, "original") and hasattr(module, "lora_down") and hasattr(module, "lora_up"):
            lin = module.original
            down, up = module.lora_down, module.lora_up
            self.scale = float(getattr(module, "scaling", 1.0))
            if getattr(module, "dropout", None) is not None and isinstance(module.dropout, nn.Dropout) and module.dropout.p > 0:
                raise NotImplementedError("LoRA dropout > 0 is not supported by the fused kernels (reference default is 0.0)")
        elif _hooked_lora(module) is not None:
            lora = _hooked_lora(module)
            if getattr(lora, "use_lora", True):
                down, up = lora.lora_down, lora.lora_up
                self.scale = float(lora.multiplier) * float(lora.alpha_scale)
        if not isinstance(lin, nn.Linear):
            raise TypeError(f"{name}: expected nn.Linear (or a LoRA wrapper around one), got {type(lin).__name__}")
        self.W, self.bias = lin.weight, lin.bias
        self.out_features, self.in_features = lin.weight.shape
        if down is not None:
            self.A_param = down.weight                       # [r_total, in]
            self.r_total = down.weight.shape[0]
            if hasattr(up, "blocks"):                        # n_separate > 1: block-diagonal up projection
                self.blocks_up = [b.weight for b in up.blocks]
                self.params = [self.A_param, *self.blocks_up]
            else:
                self.B_param = up.weight                     # [out, r]
                self.params = [self.A_param, self.B_param]
            self.r = (self.r_total + 7) // 8 * 8
            if self.r > 64:
                raise NotImplementedError(f"{name}: effective LoRA rank {self.r_total} > 64 is not supported")
            self.staged = self.blocks_up is not None or self.r != self.r_total
        # filled by the engine
        self.A = self.B = None          # bf16 tensors the kernels read ([r, in], [out, r])
        self.dA_acc = self.dB_acc = None  # f32 views into grad_flat ([in, r], [out, r])

    @property
    def has_lora(self):
        return self.r > 0

    def refresh(self):
        """(Re)build the kernel-facing A/B when they are staged copies (rank padding / block-diagonal up)."""
        if not self.has_lora:
            return
        if not self.staged:
            self.A, self.B = self.A_param.detach(), self.B_param.detach()
            return
        dev = self.A_param.device
        if self.A is None:
            self.A = torch.zeros(self.r, self.in_features, dtype=BF16, device=dev)
            self.B = torch.zeros(self.out_features, self.r, dtype=BF16, device=dev)
        self.A[: self.r_total].copy_(self.A_param.detach())
        if self.blocks_up is not None:
            n = len(self.blocks_up)
            rr, oo = self.r_total // n, self.out_features // n
            for i, w in enumerate(self.blocks_up):
                self.B[i * oo:(i + 1) * oo, i * rr:(i + 1) * rr].copy_(w.detach())
        else:
            self.B[:, : self.r_total].copy_(self.B_param.detach())

    def param_grads(self) -> List[torch.Tensor]:
        """fp32 gradients in the parameters' own layout (views / small copies of the flat accumulators)."""
        gA = self.dA_acc.t()[: self.r_total]
        if self.blocks_up is not None:
            n = len(self.blocks_up)
            rr, oo = self.r_total // n, self.out_features // n
            return [gA] + [self.dB_acc[i * oo:(i + 1) * oo, i * rr:(i + 1) * rr] for i in range(n)]
        return [gA, self.dB_acc[:, : self.r_total]]


SITE_NAMES = ("qkv", "proj", "q_linear", "kv_linear", "cproj", "w1", "w2", "w3")


def _block_sites(blk, i: int) -> Dict[str, LinearSite]:
    return {
        "qkv": LinearSite(blk.attn.qkv, f"blocks.{i}.attn.qkv"),
        "proj": LinearSite(blk.attn.proj, f"blocks.{i}.attn.proj"),
        "q_linear": LinearSite(blk.cross_attn.q_linear, f"blocks.{i}.cross_attn.q_linear"),
        "kv_linear": LinearSite(blk.cross_attn.kv_linear, f"blocks.{i}.cross_attn.kv_linear"),
        "cproj": LinearSite(blk.cross_attn.proj, f"blocks.{i}.cross_attn.proj"),
        "w1": LinearSite(blk.ffn.w1, f"blocks.{i}.ffn.w1"),
        "w2": LinearSite(blk.ffn.w2, f"blocks.{i}.ffn.w2"),
        "w3": LinearSite(blk.ffn.w3, f"blocks.{i}.ffn.w3"),
    }


# ---------------------------------------------------------------------------------------------------- extras (delta / norm / FiLM)
class Extras:
    """Optional non-LoRA trainables of one step (delta / norm-tune / FiLM methods).  All default to "absent".

    t_offset[b]      f32 [C_t]   added to the timestep embedding seen by block b (delta-A: same vector for every
                                 block and the final layer; delta-B timestep: per group)         run_delta_a.py:168
    t_offset_final   f32 [C_t]   offset seen by the final layer (delta-A only)
    film[b]          f32 [6C]    added to block b's adaLN output (FiLM, already expanded)          run_film_tta.py:129-144
    hidden[b]        f32 [C]     added to block b's output (delta-B hidden), hidden_final likewise  run_delta_b.py:318-324
    out_bias         f32 [16]    added to the prediction per channel (delta-C)                     run_delta_c.py:164-166
    norm_grads       bool        accumulate gradients of pre_crs_attn_norm / q,k RMSNorm weights   run_norm_tune_tta.py:74-98
    need_dt / need_dmod          which modulation-side gradients the backward must produce
    """

    def __init__(self, depth: int):
        self.t_offset: List[Optional[torch.Tensor]] = [None] * depth
        self.t_offset_final: Optional[torch.Tensor] = None
        self.film: List[Optional[torch.Tensor]] = [None] * depth
        self.hidden: List[Optional[torch.Tensor]] = [None] * depth
        self.hidden_final: Optional[torch.Tensor] = None
        self.out_bias: Optional[torch.Tensor] = None
        self.norm_grads = False
        self.need_dt = False
        self.need_dmod = False
        # outputs of the backward
        self.d_t: List[Optional[torch.Tensor]] = [None] * depth      # f32 [T, C_t] per block
        self.d_t_final: Optional[torch.Tensor] = None
        self.d_mod: List[Optional[torch.Tensor]] = [None] * depth    # f32 [T, 6C] per block
        self.d_hidden: List[Optional[torch.Tensor]] = [None] * depth  # f32 [C]
        self.d_hidden_final: Optional[torch.Tensor] = None
        self.d_out_bias: Optional[torch.Tensor] = None
        self.d_norm: Dict[str, torch.Tensor] = {}


class _WS:
    pass


class _Range:
    """NVTX range around a phase (B200TTA_NVTX=1): forward / backward of every block show up by name in nsys / ncu
    timelines; a no-op otherwise."""
    on = bool(os.environ.get("B200TTA_NVTX"))

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _Range.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _Range.on:
            torch.cuda.nvtx.range_pop()
        return False


# ---------------------------------------------------------------------------------------------------- full-model gradients
class FullGrads:
    """fp32 gradients of EVERY DiT parameter in one flat buffer (full-model TTA, lora_experiment/scripts/run_full_tta.py:95-219
    lets autograd produce them; here the engine's backward additionally forms dW = dY^T X / db = colsum(dY) for every
    linear it walks through, plus the norm / adaLN / embedder gradients).  One flat buffer = one all-reduce payload."""

    def __init__(self, dit):
        self.params: List[nn.Parameter] = [p for p in dit.parameters()]
        self.names = [n for n, _ in dit.named_parameters()]
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=F32, device=dev)
        self._by_id: Dict[int, torch.Tensor] = {}
        off = 0
        for p in self.params:
            self._by_id[id(p)] = self.flat[off: off + p.numel()].view(p.shape)
            off += p.numel()

        # per-block [lo, hi) ranges of the flat buffer (module order keeps a block's parameters contiguous): the unit of
        # the overlapped gradient all-reduce (stepper.TTAStepper, world > 1)
        self.block_ranges: List[Tuple[int, int]] = []
        off = 0
        spans: Dict[int, List[int]] = {}
        for n, p in zip(self.names, self.params):
            if n.startswith("blocks."):
                b = int(n.split(".")[1])
                sp = spans.setdefault(b, [off, off])
                if sp[1] != off:
                    raise RuntimeError(f"parameters of block {b} are not contiguous in module order")
                sp[1] = off + p.numel()
            off += p.numel()
        for b in sorted(spans):
            self.block_ranges.append((spans[b][0], spans[b][1]))

    def g(self, p) -> torch.Tensor:
        return self._by_id[id(p)]

    def named(self):
        return {n: self._by_id[id(p)] for n, p in zip(self.names, self.params)}


def _silu(x):
    return x * torch.sigmoid(x)


# ---------------------------------------------------------------------------------------------------- engine
class TTAEngine:
    def __init__(self, dit):
        self.dit = dit
        cfg = dit.config
        self.C, self.F, self.H, self.D = cfg.hidden_size, cfg.ffn_dim, cfg.num_heads, cfg.head_dim
        self.L, self.Ct = cfg.depth, cfg.adaln_tembed_dim
        self.softmax_scale = self.D ** -0.5
        self.geo: Optional[Geometry] = None
        self.ws = None
        self.sites: List[Dict[str, LinearSite]] = []
        self._site_sig = None
        self.grad_flat: Optional[torch.Tensor] = None
        self._text_cache = None
        cfgb = getattr(cfg, "bsa_params", None) or {}
        # block-sparse self-attention (720p refinement stage, configs[4]): see bsa.py for the definition
        self.bsa = dict(sparsity=float(cfgb.get("sparsity", 0.9375)), chunk=tuple(cfgb.get("chunk", (4, 4, 8)))) \
            if getattr(cfg, "enable_bsa", False) else None
        self.full: Optional[FullGrads] = None   # full-model TTA: gradients for every parameter (enable_full_grads)
        # full-model TTA on several ranks: called with b right after block b's gradients are final, and once (None)
        # before the late pieces (norm weights, embedders) are written into the flat buffer
        self.on_block_grads = None
        self._stash = None          # activation stash of the training geometry (see _ensure_stash)
        self._stash_on = False      # the forward in flight writes / the recompute reads the stash
        self.device = dit.x_embedder.proj.weight.device
        if dit.x_embedder.proj.weight.dtype != BF16:
            raise TypeError("B200DiT parameters must be bf16 (the kernels read them in place)")

    # ------------------------------------------------------------------ adapters
    def _signature(self):
        # the module objects themselves (compared by identity): holding them means a re-injected adapter can never be
        # mistaken for a freed one whose id() the allocator handed out again
        sig = []
        for blk in self.dit.blocks:
            for m in (blk.attn.qkv, blk.attn.proj, blk.cross_attn.q_linear, blk.cross_attn.kv_linear, blk.cross_attn.proj,
                      blk.ffn.w1, blk.ffn.w2, blk.ffn.w3):
                sig.append(m)
                sig.append(_hooked_lora(m))
        return sig

    def resolve_sites(self, force: bool = False):
        sig = self._signature()
        old = self._site_sig
        if not force and old is not None and len(old) == len(sig) and all(a is b for a, b in zip(old, sig)):
            return
        self._site_sig = sig
        self.sites = [_block_sites(blk, i) for i, blk in enumerate(self.dit.blocks)]
        total = 0
        for bs in self.sites:
            for s in bs.values():
                if s.has_lora:
                    total += (s.in_features + s.out_features) * s.r
        self.grad_flat = torch.zeros(max(total, 1), dtype=F32, device=self.device)
        off = 0
        for bs in self.sites:
            for nm in SITE_NAMES:
                s = bs[nm]
                if s.has_lora:
                    s.dA_acc = self.grad_flat[off: off + s.in_features * s.r].view(s.in_features, s.r)
                    off += s.in_features * s.r
                    s.dB_acc = self.grad_flat[off: off + s.out_features * s.r].view(s.out_features, s.r)
                    off += s.out_features * s.r
        self.max_r = max([s.r for bs in self.sites for s in bs.values()] + [0])
        if self.ws is not None:
            self._alloc_lora_ws()

    def lora_sites(self) -> List[LinearSite]:
        return [bs[nm] for bs in self.sites for nm in SITE_NAMES if bs[nm].has_lora]

    def adapter_parameters(self) -> List[nn.Parameter]:
        self.resolve_sites()
        return [p for s in self.lora_sites() for p in s.params]

    # ------------------------------------------------------------------ full-model gradients
    def enable_full_grads(self) -> FullGrads:
        if self.full is None:
            self._release_stash()
            self.full = FullGrads(self.dit)
        return self.full

    def disable_full_grads(self):
        self.full = None

    def _wgrad(self, W: torch.Tensor, bias: Optional[torch.Tensor], dy: torch.Tensor, x: torch.Tensor):
        """dW [out, in] = dY^T X and db = colsum(dY), fp32, with dY [tokens, out] and X [tokens, in] read in place."""
        out_f, in_f = W.shape[0], W.numel() // W.shape[0]
        gw = self.full.g(W).view(out_f, in_f)
        ops.gemm(out_f, in_f, [(dy, x, dy.shape[0], True, None, True)], ops.epi(ops.EPI_STORE_F32, gw))
        if bias is not None:
            ops.colsum(self.full.g(bias), dy)

    def _modulation_grads(self, lin: nn.Linear, t_in: torch.Tensor, dmod: torch.Tensor):
        """adaLN linear  mod = silu(t) W^T + b  with T (= latent frames) rows: tiny fp32 products"""
        self.full.g(lin.weight).copy_(dmod.t() @ _silu(t_in))
        self.full.g(lin.bias).copy_(dmod.sum(0))

    # ------------------------------------------------------------------ workspace
    def plan(self, geo: Geometry):
        """Allocate (once per geometry; train and early-stopping geometries are both kept) the step workspace."""
        if self.geo == geo and self.ws is not None:
            return
        self._ws_holds = None
        cache = self.__dict__.setdefault("_ws_cache", {})
        if geo in cache:
            self.geo, self.ws = geo, cache[geo]
            self._alloc_lora_ws()
            return
        if len(cache) >= 2:
            cache.clear()
        try:
            self._plan_new(geo, cache)
        except torch.OutOfMemoryError:
            # the activation stash of another geometry took the memory: it is an optimisation, give it back
            self._release_stash()
            self.ws = None
            torch.cuda.empty_cache()
            self._plan_new(geo, cache)

    def _plan_new(self, geo: Geometry, cache):
        self.geo = geo
        dev, C, F, H = self.device, self.C, self.F, self.H
        N, Nn, M, T = geo.N, geo.Nn, geo.M, geo.T
        e = lambda *s, dt=BF16: torch.empty(*s, dtype=dt, device=dev)
        ws = _WS()
        ws.xs = e(self.L + 1, N, C)
        ws.P = e(N, 64)
        ws.V = e(max(Nn, 1), 64, dt=F32)
        ws.timestep = e(T, dt=F32)
        ws.tfeat, ws.th, ws.t = e(T, self.dit.config.frequency_embedding_size, dt=F32), e(T, self.Ct, dt=F32), e(T, self.Ct, dt=F32)
        ws.t_blk = e(T, self.Ct, dt=F32)
        ws.y1, ws.y = e(M, C), e(M, C)
        # adaLN modulation of every block (2.4 MB each at T = 24): the recompute pass re-uses it instead of re-running
        # the [T, 512] x [6C, 512]^T product (a 135 us latency-bound launch per block)
        ws.mod_all, ws.modf = e(self.L, T, 6 * C, dt=F32), e(T, 2 * C, dt=F32)
        ws.mod = ws.mod_all[0]
        ws.xm1, ws.xm2 = e(N, C), e(N, C)
        ws.qkv_tmp, ws.qk = e(N, 3 * C), e(N, 2 * C)
        # self-attention output + log-sum-exp are kept for EVERY block (312 MB per block at 37k tokens) so that the
        # per-block recompute in the backward never re-runs the attention forward (the most expensive op to redo)
        ws.o_all, ws.lse_all = e(self.L, N, C), e(self.L, H, N, dt=F32)
        ws.delta = e(H, N, dt=F32)
        ws.x1_tmp, ws.x2_tmp = e(N, C), e(N, C)
        ws.xn, ws.qc, ws.qcn, ws.oc = e(Nn, C), e(Nn, C), e(Nn, C), e(Nn, C)
        ws.kvc, ws.kcn = e(M, 2 * C), e(M, C)
        ws.lsec, ws.deltac = e(H, Nn, dt=F32), e(H, Nn, dt=F32)
        ws.h, ws.h1_tmp, ws.h3_tmp = e(N, F), e(N, F), e(N, F)
        ws.qkv, ws.x1, ws.x2, ws.h1, ws.h3 = ws.qkv_tmp, ws.x1_tmp, ws.x2_tmp, ws.h1_tmp, ws.h3_tmp
        ws.xf, ws.pred = e(N, C), e(N, 64, dt=F32)
        # backward temporaries
        ws.dx = e(N, C)
        ws.g1, ws.g2 = e(N, C), e(N, C)
        ws.dh1, ws.dh3 = e(N, F), e(N, F)
        ws.dqkv, ws.dqk = e(N, 3 * C), e(N, 2 * C)
        ws.dqc, ws.dkvc = e(Nn, C), e(M, 2 * C)
        ws.dpred = e(max(Nn, 1), 64)
        ws.loss = torch.zeros(1, dtype=F32, device=dev)
        ws.dmod, ws.dmodf, ws.dt = e(T, 6 * C, dt=F32), e(T, 2 * C, dt=F32), e(T, self.Ct, dt=F32)
        ws.dt_total = e(T, self.Ct, dt=F32)      # full-model TTA: d loss / d(timestep embedding), summed over its users
        ws.d_y = e(M, C)                         # full-model TTA: d loss / d(text rows), summed over the blocks
        ws.branch_a = ws.branch_m = None  # branch outputs, allocated on demand (FiLM / delta gate gradients)
        if self.bsa is not None:
            from . import bsa as _bsa
            ct = self.bsa["chunk"][0]
            if geo.n_cond % ct:
                raise ValueError(f"block-sparse attention needs the context frames ({geo.n_cond}) to be whole chunks of {ct}")
            ws.bsa_perm, ws.bsa_inv = _bsa.block_permutation(geo.T, geo.gh, geo.gw, self.bsa["chunk"], device=dev)
            ws.bsa_nctx = geo.Nc // _bsa.BLOCK
            # block-major copies of q, k, v, o and of their gradients; per-block lists are kept for the backward
            ws.bq, ws.bk, ws.bv, ws.bo = e(N, C), e(N, C), e(N, C), e(N, C)
            ws.bdo, ws.bdq, ws.bdk, ws.bdv = e(N, C), e(N, C), e(N, C), e(N, C)
            ws.bsa_lists = [None] * self.L
        self.ws = ws
        cache[geo] = ws
        self._alloc_lora_ws()

    # ------------------------------------------------------------------ activation stash
    def _release_stash(self):
        self._stash = None
        self._stash_on = False

    def _ensure_stash(self, geo: Geometry):
        """Spend the HBM that is still free on per-block activations of the training forward (x1, x2, qkv, h1/h3, in
        that order, each for as many leading blocks as fit).  B200TTA_STASH_GB caps it (0 disables); 12 GB stay free."""
        if self._stash is not None and self._stash["geo"] == geo:
            return
        self._release_stash()
        torch.cuda.empty_cache()
        free, _ = torch.cuda.mem_get_info(self.device)
        budget = free - (12 << 30)
        cap = os.environ.get("B200TTA_STASH_GB")
        if cap is not None:
            budget = min(budget, int(float(cap) * (1 << 30)))
        N, C, F, L = geo.N, self.C, self.F, self.L
        st = {"geo": geo}
        for name, cols, parts in (("x1", C, 1), ("x2", C, 1), ("qkv", 3 * C, 1), ("h", F, 2)):
            per_block = N * cols * 2 * parts
            k = int(max(0, min(L, budget // per_block)))
            budget -= k * per_block
            st["k_" + name] = k
            st[name] = torch.empty(k, parts, N, cols, dtype=BF16, device=self.device) if k else None
        self._stash = st

    def _bind(self, b: int):
        """Point the per-block workspace names at block b's slots."""
        ws, st = self.ws, self._stash if self._stash_on else None
        ws.o, ws.lse = ws.o_all[b], ws.lse_all[b]
        ws.mod = ws.mod_all[b]
        ws.x1 = st["x1"][b, 0] if st and b < st["k_x1"] else ws.x1_tmp
        ws.x2 = st["x2"][b, 0] if st and b < st["k_x2"] else ws.x2_tmp
        ws.qkv = st["qkv"][b, 0] if st and b < st["k_qkv"] else ws.qkv_tmp
        if st and b < st["k_h"]:
            ws.h1, ws.h3 = st["h"][b, 0], st["h"][b, 1]
        else:
            ws.h1, ws.h3 = ws.h1_tmp, ws.h3_tmp

    def _stashed(self, b: int, name: str) -> bool:
        return self._stash_on and self._stash is not None and b < self._stash["k_" + name]

    def _xa_only(self, s, x, nm):
        """recompute pass, output already stashed: only the LoRA down-projection x A^T is needed (for dB)."""
        if s.has_lora:
            ops.lora_down(self._xa(s, nm, x.shape[0]), x, s.A, scale=s.scale)

    def _alloc_lora_ws(self):
        ws, geo = self.ws, self.geo
        r = max(getattr(self, "max_r", 0), 8)
        if getattr(ws, "xa_r", 0) >= r:
            return
        ws.xa_r = r
        e = lambda *s: torch.empty(*s, dtype=BF16, device=self.device)
        rows = max(geo.N, geo.M)
        # XA of every LoRA'd linear of one block is kept between the block's forward and backward
        ws.xa = {nm: e(rows, r) for nm in SITE_NAMES}
        ws.u = e(rows, r)

    # ------------------------------------------------------------------ small pieces
    def _xa(self, s, nm, rows):
        return self.ws.xa[nm].view(-1)[: rows * s.r].view(rows, s.r)

    def _linear_bwd(self, s: LinearSite, dy, x, e, nm):
        if self.full is not None:
            self._wgrad(s.W, s.bias, dy, x)
        if s.has_lora:
            rows = dy.shape[0]
            u = self.ws.u.view(-1)[: rows * s.r].view(rows, s.r)
            ops.lora_linear_bwd(dy, s.W, e, x=x, A=s.A, B=s.B, XA=self._xa(s, nm, rows), U=u, dA_acc=s.dA_acc,
                                dB_acc=s.dB_acc, scale=s.scale)
        elif e is not None:
            ops.lora_linear_bwd(dy, s.W, e)

    def _linear_fwd_xa(self, s, x, e, nm):
        if s.has_lora:
            ops.lora_linear_fwd(x, s.W, e, A=s.A, B=s.B, XA=self._xa(s, nm, x.shape[0]), scale=s.scale)
        else:
            ops.lora_linear_fwd(x, s.W, e)

    # ------------------------------------------------------------------ text / time embeddings
    def pack_text(self, prompt_embeds: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
        """[B=1,1,Ltxt,Cc] (+ mask [1,Ltxt]) -> text rows [M, Cc] bf16 fed to y_embedder (run_delta_a.py:170-192).

        text_tokens_zero_pad = True  : every token stays a key/value; rows of padded tokens are ZEROED after
                                       y_embedder (the reference multiplies by the mask, then resets the mask to ones),
                                       so M = Ltxt and ``self._text_keep`` holds the 0/1 row mask;
        text_tokens_zero_pad = False : only valid tokens are packed (y_embedder is row-wise, so packing before or
                                       after it is the same arithmetic), M = mask.sum().
        """
        # one-entry cache keyed on the SOURCE tensors themselves (identity + in-place version): holding them keeps their
        # addresses from being recycled by the caching allocator for the next video's embeddings (a data_ptr key
        # would then hit and silently train on the previous prompt)
        c = self._text_cache
        if (c is not None and c[0] is prompt_embeds and c[1] == prompt_embeds._version and c[2] is mask
                and (mask is None or c[3] == mask._version)):
            self._text_keep = c[5]
            return c[4]
        if prompt_embeds.shape[0] != 1:
            raise NotImplementedError("batch size 1 only (as in every reference run)")
        pe = prompt_embeds.reshape(-1, prompt_embeds.shape[-1])
        keep = None
        if mask is not None:
            if self.dit.text_tokens_zero_pad:
                keep = (mask.reshape(-1) != 0).to(device=self.device, dtype=BF16)[:, None].contiguous()
            else:
                idx = mask.reshape(-1).nonzero(as_tuple=False).flatten()
                pe = pe.index_select(0, idx.to(pe.device))
        pe = pe.to(device=self.device, dtype=BF16).contiguous()
        self._text_cache = (prompt_embeds, prompt_embeds._version, mask, None if mask is None else mask._version, pe, keep)
        self._text_keep = keep
        return pe

    def _embed_text(self, text_valid):
        ws, ye = self.ws, self.dit.y_embedder.y_proj
        M = text_valid.shape[0]
        ops.gemm(M, self.C, [(text_valid, ye[0].weight, text_valid.shape[1], False, None)],
                 ops.epi(ops.EPI_GELU, ws.y1, bias=ye[0].bias))
        ops.gemm(M, self.C, [(ws.y1, ye[2].weight, self.C, False, None)], ops.epi(ops.EPI_STORE, ws.y, bias=ye[2].bias))
        if getattr(self, "_text_keep", None) is not None:
            ws.y.mul_(self._text_keep)   # zero the padded tokens (they remain keys/values)

    def _embed_time(self):
        ws, te = self.ws, self.dit.t_embedder
        ops.timestep_sinusoid(ws.tfeat, ws.timestep)
        ops.skinny_linear(ws.th, ws.tfeat, te.mlp[0].weight, te.mlp[0].bias)
        ops.skinny_linear(ws.t, ws.th, te.mlp[2].weight, te.mlp[2].bias, act=1)
        if te._forward_hooks:  # generation-time hooks (e.g. DeltaAWrapper.apply_to_dit); forward only
            t = ws.t
            for hook in te._forward_hooks.values():
                out = hook(te, (ws.timestep,), t)
                if out is not None:
                    t = out
            ws.t.copy_(t)

    def _t_for_block(self, b, ex: Optional[Extras]):
        if ex is not None and ex.t_offset[b] is not None:
            torch.add(self.ws.t, ex.t_offset[b].to(F32)[None, :], out=self.ws.t_blk)
            return self.ws.t_blk
        return self.ws.t

    # ------------------------------------------------------------------ block forward
    def _block_fwd(self, b: int, x_in, x_out, ex: Optional[Extras], recompute: bool = False):
        ws, geo, C, H, D = self.ws, self.geo, self.C, self.H, self.D
        self._bind(b)
        blk, st = self.dit.blocks[b], self.sites[b]
        N, Nc, Nn, M, tpf = geo.N, geo.Nc, geo.Nn, geo.M, geo.tpf
        ada = blk.adaLN_modulation[1]
        film = None
        if ex is not None and ex.film[b] is not None:
            film = ex.film[b].to(F32)[None, :].expand(geo.T, -1).contiguous()
        elif blk.adaLN_modulation._forward_hooks:
            # dit_forward_autograd turns such hooks into an adapter (HookedModulationAdapter); reaching this line means
            # the engine was driven directly (TTAStepper) with hooks installed but no adapter describing them
            raise NotImplementedError(
                "forward hooks on blocks[%d].adaLN_modulation are only honoured through B200DiT.forward(); for the fused "
                "stepper use longcat_video_tta_b200.adapters.FiLMAdapterWrapper (same constructor as the reference's)" % b)
        if not recompute:      # the recompute pass finds this block's modulation where the forward left it
            ops.skinny_linear(ws.mod, self._t_for_block(b, ex), ada.weight, ada.bias, act=1, addend=film)
        mod = ws.mod
        shift_msa, scale_msa, gate_msa = mod[:, 0:C], mod[:, C:2 * C], mod[:, 2 * C:3 * C]
        shift_mlp, scale_mlp, gate_mlp = mod[:, 3 * C:4 * C], mod[:, 4 * C:5 * C], mod[:, 5 * C:6 * C]
        keep_branch = ex is not None and ex.need_dmod

        # ---- self attention
        ops.ln_mod_fwd(ws.xm1, x_in, scale_msa, shift_msa, tokens_per_frame=tpf)
        s = st["qkv"]
        if recompute and self._stashed(b, "qkv"):
            self._xa_only(s, ws.xm1, "qkv")
        else:
            self._linear_fwd_xa(s, ws.xm1, ops.epi(ops.EPI_STORE, ws.qkv, bias=s.bias), "qkv")
        ops.qk_rmsnorm_rope_fwd(ws.qk, ws.qkv, blk.attn.q_norm.weight, blk.attn.k_norm.weight, H, H,
                                grid_hw=(geo.gh, geo.gw), rope_base=self.dit.config.rope_base)
        q = ws.qk.view(N, 2 * H, D)[:, :H]
        k = ws.qk.view(N, 2 * H, D)[:, H:]
        v = ws.qkv.view(N, 3 * H, D)[:, 2 * H:]
        if getattr(self, "_ctx_fill", False):   # keep the context rows' K / V for the noise-rows-only forwards that follow
            self._ctx["k"][b].view(Nc, H, D).copy_(k[:Nc])
            self._ctx["v"][b].view(Nc, H, D).copy_(v[:Nc])
        if not recompute:
            if self.bsa is None:
                ops.attn_fwd(q, k, v, ws.o.view(N, H, D), ws.lse, geo.self_segments(), self.softmax_scale)
            else:
                self._bsa_fwd(b, q, k, v)
        s = st["proj"]
        if recompute and self._stashed(b, "x1") and not keep_branch:
            self._xa_only(s, ws.o, "proj")
        else:
            self._linear_fwd_xa(s, ws.o, ops.epi(ops.EPI_GATE_RESID, ws.x1, bias=s.bias, resid=x_in, gate=gate_msa,
                                                 tokens_per_frame=tpf, d2=ws.branch_a if keep_branch else None), "proj")
        # ---- cross attention (noise tokens only; context rows pass through)
        if Nn > 0:
            nrm = blk.pre_crs_attn_norm
            ops.ln_mod_fwd(ws.xn, ws.x1[Nc:], nrm.weight, nrm.bias, tokens_per_frame=tpf, affine=True)
            s = st["q_linear"]
            self._linear_fwd_xa(s, ws.xn, ops.epi(ops.EPI_STORE, ws.qc, bias=s.bias), "q_linear")
            s = st["kv_linear"]
            self._linear_fwd_xa(s, ws.y, ops.epi(ops.EPI_STORE, ws.kvc, bias=s.bias), "kv_linear")
            ops.qk_rmsnorm_rope_fwd(ws.qcn, ws.qc, blk.cross_attn.q_norm.weight, None, H, 0, rope=False)
            ops.qk_rmsnorm_rope_fwd(ws.kcn, ws.kvc[:, :C], blk.cross_attn.k_norm.weight, None, H, 0, rope=False)
            vc = ws.kvc.view(M, 2 * H, D)[:, H:]
            ops.attn_fwd(ws.qcn.view(Nn, H, D), ws.kcn.view(M, H, D), vc, ws.oc.view(Nn, H, D), ws.lsec,
                         [(0, Nn, M)], self.softmax_scale)
            s = st["cproj"]
            if recompute and self._stashed(b, "x2"):
                self._xa_only(s, ws.oc, "cproj")
            else:
                if Nc > 0:
                    ws.x2[:Nc].copy_(ws.x1[:Nc])
                self._linear_fwd_xa(s, ws.oc, ops.epi(ops.EPI_GATE_RESID, ws.x2[Nc:], bias=s.bias, resid=ws.x1[Nc:]), "cproj")
        elif not (recompute and self._stashed(b, "x2")):
            ws.x2.copy_(ws.x1)
        # ---- FFN
        s1, s3, s2 = st["w1"], st["w3"], st["w2"]
        # the recompute pass never needs the block output (xs[b+1] is kept): w2 only re-runs for its own LoRA / FiLM needs
        need_w2 = (not recompute) or keep_branch or s2.has_lora
        need_h = need_w2 or self.full is not None      # full-model TTA: dW of w2 needs its input h = silu(h1) h3
        h_stashed = recompute and self._stashed(b, "h")
        if not h_stashed or s1.has_lora or s3.has_lora or self.full is not None:
            ops.ln_mod_fwd(ws.xm2, ws.x2, scale_mlp, shift_mlp, tokens_per_frame=tpf)
        if s1.has_lora or s3.has_lora:
            if h_stashed:
                self._xa_only(s1, ws.xm2, "w1")
                self._xa_only(s3, ws.xm2, "w3")
            else:
                self._linear_fwd_xa(s1, ws.xm2, ops.epi(ops.EPI_STORE, ws.h1), "w1")
                self._linear_fwd_xa(s3, ws.xm2, ops.epi(ops.EPI_STORE, ws.h3), "w3")
            if need_h:
                ops.swiglu_fwd(ws.h, ws.h1, ws.h3)
        elif not h_stashed:
            ops.lora_linear_fwd(ws.xm2, s1.W, ops.epi(ops.EPI_SWIGLU, ws.h, d2=ws.h1, d3=ws.h3), W_hi=s3.W)
        elif need_h:
            ops.swiglu_fwd(ws.h, ws.h1, ws.h3)
        if need_w2:
            self._linear_fwd_xa(s2, ws.h, ops.epi(ops.EPI_GATE_RESID, x_out, resid=ws.x2, gate=gate_mlp, tokens_per_frame=tpf,
                                                  d2=ws.branch_m if keep_branch else None), "w2")
            if ex is not None and ex.hidden[b] is not None:
                x_out.add_(ex.hidden[b].to(BF16)[None, :])
        self._ws_holds = b
        if getattr(self, "debug", None) is not None:
            for nm in ("mod", "xm1", "qkv", "qk", "o", "x1", "xn", "qc", "qcn", "kvc", "kcn", "oc", "x2", "xm2", "h1", "h3", "h", "y"):
                self._tap(b, "f_" + nm, getattr(ws, nm))
            self._tap(b, "f_xin", x_in)
            self._tap(b, "f_xout", x_out)

    # ------------------------------------------------------------------ block-sparse self-attention (configs[4])
    def _bsa_gather(self, q, k, v):
        ws, N, H, D = self.ws, self.geo.N, self.H, self.D
        bq, bk, bv = (t.view(N, H, D) for t in (ws.bq, ws.bk, ws.bv))
        ops.gather_rows(bq, q, ws.bsa_perm)               # (t, h, w) row-major -> block-major token order
        ops.gather_rows(bk, k, ws.bsa_perm)
        ops.gather_rows(bv, v, ws.bsa_perm)
        return bq, bk, bv

    def _bsa_fwd(self, b: int, q, k, v):
        """self-attention of block b through the list-driven kernels; O comes back in row-major order, LSE stays block-major
        (only the block-sparse backward reads it) and the block lists are kept for that backward."""
        from . import bsa as _bsa
        ws, N, H, D = self.ws, self.geo.N, self.H, self.D
        bq, bk, bv = self._bsa_gather(q, k, v)
        lists = _bsa.select_blocks(bq, bk, self.bsa["sparsity"], ws.bsa_nctx)
        ws.bsa_lists[b] = lists
        ops.attn_bsa_fwd(bq, bk, bv, ws.bo.view(N, H, D), ws.lse, lists.q_off, lists.q_idx, self.softmax_scale)
        ops.gather_rows(ws.o.view(N, H, D), ws.bo.view(N, H, D), ws.bsa_inv)

    def _bsa_bwd(self, b: int, dq, dk, dv, do, q, k, v):
        ws, N, H, D = self.ws, self.geo.N, self.H, self.D
        lists = ws.bsa_lists[b]
        bq, bk, bv = self._bsa_gather(q, k, v)
        bo, bdo = ws.bo.view(N, H, D), ws.bdo.view(N, H, D)
        ops.gather_rows(bo, ws.o.view(N, H, D), ws.bsa_perm)
        ops.gather_rows(bdo, do, ws.bsa_perm)
        bdq, bdk, bdv = (t.view(N, H, D) for t in (ws.bdq, ws.bdk, ws.bdv))
        ops.attn_bsa_bwd(bdq, bdk, bdv, bdo, bo, ws.lse, ws.delta, bq, bk, bv, lists.q_off, lists.q_idx, lists.k_off,
                         lists.k_idx, self.softmax_scale)
        ops.gather_rows(dq, bdq, ws.bsa_inv)              # straight into the strided dq / dk / dv views
        ops.gather_rows(dk, bdk, ws.bsa_inv)
        ops.gather_rows(dv, bdv, ws.bsa_inv)

    # ------------------------------------------------------------------ forward with cached context K/V
    # The context (clean conditioning) frames carry timestep 0 and only attend to themselves, so for a fixed video, text
    # and adapter state their rows are the SAME in every forward whatever sigma / noise the other frames get.  The
    # anchored early stopper evaluates sigmas x draws = 6 forwards on [cond | val] per check (early_stopping.py:293-317
    # -> common.py:492-559): the first one stores every block's context K (normed + RoPE'd) and V, the other five run
    # the noised rows only against [cached context K/V | their own K/V].  Same idea as upstream's use_kv_cache in
    # generate_vc (common.py:598-611).  Forward only.
    def _ensure_ctx_cache(self, geo: Geometry):
        c = getattr(self, "_ctx", None)
        if c is None or c["geo"] != geo:
            self._ctx = None
            e = lambda: torch.empty(self.L, geo.Nc, self.C, dtype=BF16, device=self.device)
            try:
                k, v = e(), e()
            except torch.OutOfMemoryError:      # the training stash is an optimisation: give its memory back
                self._release_stash()
                torch.cuda.empty_cache()
                k, v = e(), e()
            self._ctx = c = {"geo": geo, "k": k, "v": v, "valid": False}
        return c

    def _block_fwd_noise_rows(self, b: int, x_in, x_out, ex: Optional[Extras]):
        """_block_fwd restricted to the noised rows [Nc:], self-attention keys/values = cached context + own rows."""
        ws, geo, C, H, D = self.ws, self.geo, self.C, self.H, self.D
        self._bind(b)
        blk, st, cache = self.dit.blocks[b], self.sites[b], self._ctx
        N, Nc, Nn, M, tpf, Tc = geo.N, geo.Nc, geo.Nn, geo.M, geo.tpf, geo.n_cond
        ada = blk.adaLN_modulation[1]
        film = None
        if ex is not None and ex.film[b] is not None:
            film = ex.film[b].to(F32)[None, :].expand(geo.T, -1).contiguous()
        ops.skinny_linear(ws.mod, self._t_for_block(b, ex), ada.weight, ada.bias, act=1, addend=film)
        mod = ws.mod[Tc:]                                   # modulation rows of the noised frames
        shift_msa, scale_msa, gate_msa = mod[:, 0:C], mod[:, C:2 * C], mod[:, 2 * C:3 * C]
        shift_mlp, scale_mlp, gate_mlp = mod[:, 3 * C:4 * C], mod[:, 4 * C:5 * C], mod[:, 5 * C:6 * C]
        xin = x_in[Nc:]
        # ---- self attention: queries = noised rows, keys / values = all rows
        ops.ln_mod_fwd(ws.xm1[Nc:], xin, scale_msa, shift_msa, tokens_per_frame=tpf)
        s = st["qkv"]
        self._linear_fwd_xa(s, ws.xm1[Nc:], ops.epi(ops.EPI_STORE, ws.qkv[Nc:], bias=s.bias), "qkv")
        ops.qk_rmsnorm_rope_fwd(ws.qk[Nc:], ws.qkv[Nc:], blk.attn.q_norm.weight, blk.attn.k_norm.weight, H, H,
                                grid_hw=(geo.gh, geo.gw), row_offset=Nc, rope_base=self.dit.config.rope_base)
        q = ws.qk.view(N, 2 * H, D)[:, :H]
        k = ws.qk.view(N, 2 * H, D)[:, H:]
        v = ws.qkv.view(N, 3 * H, D)[:, 2 * H:]
        k[:Nc].copy_(cache["k"][b].view(Nc, H, D))
        v[:Nc].copy_(cache["v"][b].view(Nc, H, D))
        ops.attn_fwd(q, k, v, ws.o.view(N, H, D), ws.lse, [(Nc, N, N)], self.softmax_scale)
        s = st["proj"]
        self._linear_fwd_xa(s, ws.o[Nc:], ops.epi(ops.EPI_GATE_RESID, ws.x1[Nc:], bias=s.bias, resid=xin, gate=gate_msa,
                                                  tokens_per_frame=tpf), "proj")
        # ---- cross attention (the noised rows are the only ones that ever had it)
        nrm = blk.pre_crs_attn_norm
        ops.ln_mod_fwd(ws.xn, ws.x1[Nc:], nrm.weight, nrm.bias, tokens_per_frame=tpf, affine=True)
        s = st["q_linear"]
        self._linear_fwd_xa(s, ws.xn, ops.epi(ops.EPI_STORE, ws.qc, bias=s.bias), "q_linear")
        s = st["kv_linear"]
        self._linear_fwd_xa(s, ws.y, ops.epi(ops.EPI_STORE, ws.kvc, bias=s.bias), "kv_linear")
        ops.qk_rmsnorm_rope_fwd(ws.qcn, ws.qc, blk.cross_attn.q_norm.weight, None, H, 0, rope=False)
        ops.qk_rmsnorm_rope_fwd(ws.kcn, ws.kvc[:, :C], blk.cross_attn.k_norm.weight, None, H, 0, rope=False)
        vc = ws.kvc.view(M, 2 * H, D)[:, H:]
        ops.attn_fwd(ws.qcn.view(Nn, H, D), ws.kcn.view(M, H, D), vc, ws.oc.view(Nn, H, D), ws.lsec,
                     [(0, Nn, M)], self.softmax_scale)
        s = st["cproj"]
        self._linear_fwd_xa(s, ws.oc, ops.epi(ops.EPI_GATE_RESID, ws.x2[Nc:], bias=s.bias, resid=ws.x1[Nc:]), "cproj")
        # ---- FFN
        ops.ln_mod_fwd(ws.xm2[Nc:], ws.x2[Nc:], scale_mlp, shift_mlp, tokens_per_frame=tpf)
        s1, s3, s2 = st["w1"], st["w3"], st["w2"]
        if s1.has_lora or s3.has_lora:
            self._linear_fwd_xa(s1, ws.xm2[Nc:], ops.epi(ops.EPI_STORE, ws.h1[Nc:]), "w1")
            self._linear_fwd_xa(s3, ws.xm2[Nc:], ops.epi(ops.EPI_STORE, ws.h3[Nc:]), "w3")
            ops.swiglu_fwd(ws.h[Nc:], ws.h1[Nc:], ws.h3[Nc:])
        else:
            ops.lora_linear_fwd(ws.xm2[Nc:], s1.W, ops.epi(ops.EPI_SWIGLU, ws.h[Nc:], d2=ws.h1[Nc:], d3=ws.h3[Nc:]), W_hi=s3.W)
        self._linear_fwd_xa(s2, ws.h[Nc:], ops.epi(ops.EPI_GATE_RESID, x_out[Nc:], resid=ws.x2[Nc:], gate=gate_mlp,
                                                  tokens_per_frame=tpf), "w2")
        if ex is not None and ex.hidden[b] is not None:
            x_out[Nc:].add_(ex.hidden[b].to(BF16)[None, :])
        self._ws_holds = None

    # ------------------------------------------------------------------ block backward (dx in ws.dx, in place)
    def _block_bwd(self, b: int, x_in, ex: Optional[Extras]):
        ws, geo, C, H, D = self.ws, self.geo, self.C, self.H, self.D
        self._bind(b)
        blk, st = self.dit.blocks[b], self.sites[b]
        N, Nc, Nn, M, tpf = geo.N, geo.Nc, geo.Nn, geo.M, geo.tpf
        mod = ws.mod
        scale_msa, gate_msa = mod[:, C:2 * C], mod[:, 2 * C:3 * C]
        scale_mlp, gate_mlp = mod[:, 4 * C:5 * C], mod[:, 5 * C:6 * C]
        dx = ws.dx
        tap = self._tap
        tap(b, "dx_out", dx)
        want_mod = ex is not None and ex.need_dmod
        dmod = ws.dmod if want_mod else None
        if want_mod:
            dmod.zero_()
        if ex is not None and ex.hidden[b] is not None:
            ex.d_hidden[b] = dx.float().sum(0)

        # ---- FFN
        ops.gate_mul(ws.g1, dx, gate_mlp, tokens_per_frame=tpf, branch=ws.branch_m if want_mod else None,
                     dgate_acc=dmod[:, 5 * C:6 * C] if want_mod else None)
        tap(b, "ffn_gout", ws.g1)
        s1, s3, s2 = st["w1"], st["w3"], st["w2"]
        if s1.has_lora or s3.has_lora:
            self._linear_bwd(s2, ws.g1, ws.h, ops.epi(ops.EPI_STORE, ws.h), "w2")      # dh overwrites h (h no longer needed)
            ops.swiglu_bwd(ws.dh1, ws.dh3, ws.h, ws.h1, ws.h3)
            self._linear_bwd(s1, ws.dh1, ws.xm2, ops.epi(ops.EPI_STORE, ws.g2), "w1")
            self._linear_bwd(s3, ws.dh3, ws.xm2, ops.epi(ops.EPI_STORE, ws.g1), "w3")
            ws.g2.add_(ws.g1)
        else:
            self._linear_bwd(s2, ws.g1, ws.h, ops.epi(ops.EPI_SWIGLU_BWD, ws.dh1, d2=ws.dh3, aux1=ws.h1, aux2=ws.h3), "w2")
            ops.gemm(N, C, [(ws.dh1, s1.W, self.F, True, None), (ws.dh3, s3.W, self.F, True, None)],
                     ops.epi(ops.EPI_STORE, ws.g2))
            if self.full is not None:
                self._wgrad(s1.W, s1.bias, ws.dh1, ws.xm2)
                self._wgrad(s3.W, s3.bias, ws.dh3, ws.xm2)
        tap(b, "ffn_gin", ws.g2)
        ops.ln_mod_bwd(dx, ws.g2, ws.x2, scale_mlp, dx_resid=dx, tokens_per_frame=tpf,
                       dscale_acc=dmod[:, 4 * C:5 * C] if want_mod else None,
                       dshift_acc=dmod[:, 3 * C:4 * C] if want_mod else None)
        tap(b, "dx_after_ffn", dx)
        # ---- cross attention
        if Nn > 0:
            dxn = dx[Nc:]
            s = st["cproj"]
            self._linear_bwd(s, dxn, ws.oc, ops.epi(ops.EPI_STORE, ws.g1[:Nn]), "cproj")          # dOc
            vc = ws.kvc.view(M, 2 * H, D)[:, H:]
            dkc = ws.dkvc.view(M, 2 * H, D)[:, :H]
            dvc = ws.dkvc.view(M, 2 * H, D)[:, H:]
            ops.attn_bwd(ws.dqc.view(Nn, H, D), dkc, dvc, ws.g1[:Nn].view(Nn, H, D), ws.oc.view(Nn, H, D), ws.lsec,
                         ws.deltac, ws.qcn.view(Nn, H, D), ws.kcn.view(M, H, D), vc, [(0, Nn, M)], self.softmax_scale)
            ng = ex is not None and ex.norm_grads
            cq, ck = blk.cross_attn.q_norm.weight, blk.cross_attn.k_norm.weight
            ops.qk_rmsnorm_rope_bwd(ws.dqc, ws.dqc, ws.qc, cq, None, H, 0, rope=False,
                                    dwq_acc=self._ngrad(ex, b, "cross_attn.q_norm.weight", cq) if ng else None)
            ops.qk_rmsnorm_rope_bwd(ws.dkvc[:, :C], ws.dkvc[:, :C], ws.kvc[:, :C], ck, None, H, 0, rope=False,
                                    dwq_acc=self._ngrad(ex, b, "cross_attn.k_norm.weight", ck) if ng else None)
            if self.full is not None:     # the text embedder trains too: d y accumulates over the blocks (resid = running sum)
                self._linear_bwd(st["kv_linear"], ws.dkvc, ws.y, ops.epi(ops.EPI_GATE_RESID, ws.d_y, resid=ws.d_y), "kv_linear")
            else:
                self._linear_bwd(st["kv_linear"], ws.dkvc, ws.y, None, "kv_linear")                # adapter grads only
            s = st["q_linear"]
            self._linear_bwd(s, ws.dqc, ws.xn, ops.epi(ops.EPI_STORE, ws.g2[:Nn]), "q_linear")    # d xn
            tap(b, "cross_gin", ws.g2[:Nn])
            nrm = blk.pre_crs_attn_norm
            ops.ln_mod_bwd(dxn, ws.g2[:Nn], ws.x1[Nc:], nrm.weight, dx_resid=dxn, tokens_per_frame=tpf, affine=True,
                           dscale_acc=self._ngrad(ex, b, "pre_crs_attn_norm.weight", nrm.weight) if ng else None,
                           dshift_acc=self._ngrad(ex, b, "pre_crs_attn_norm.bias", nrm.bias) if ng else None)
        tap(b, "dx_after_cross", dx)
        # ---- self attention
        ops.gate_mul(ws.g1, dx, gate_msa, tokens_per_frame=tpf, branch=ws.branch_a if want_mod else None,
                     dgate_acc=dmod[:, 2 * C:3 * C] if want_mod else None)
        tap(b, "attn_gout", ws.g1)
        s = st["proj"]
        self._linear_bwd(s, ws.g1, ws.o, ops.epi(ops.EPI_STORE, ws.g2), "proj")                   # dO
        q = ws.qk.view(N, 2 * H, D)[:, :H]
        k = ws.qk.view(N, 2 * H, D)[:, H:]
        v = ws.qkv.view(N, 3 * H, D)[:, 2 * H:]
        dq = ws.dqk.view(N, 2 * H, D)[:, :H]
        dk = ws.dqk.view(N, 2 * H, D)[:, H:]
        dv = ws.dqkv.view(N, 3 * H, D)[:, 2 * H:]
        if self.bsa is None:
            ops.attn_bwd(dq, dk, dv, ws.g2.view(N, H, D), ws.o.view(N, H, D), ws.lse, ws.delta, q, k, v,
                         geo.self_segments(), self.softmax_scale)
        else:
            self._bsa_bwd(b, dq, dk, dv, ws.g2.view(N, H, D), q, k, v)
        ng = ex is not None and ex.norm_grads
        ops.qk_rmsnorm_rope_bwd(ws.dqkv, ws.dqk, ws.qkv, blk.attn.q_norm.weight, blk.attn.k_norm.weight, H, H,
                                grid_hw=(geo.gh, geo.gw), rope_base=self.dit.config.rope_base,
                                dwq_acc=self._ngrad(ex, b, "attn.q_norm.weight", blk.attn.q_norm.weight) if ng else None,
                                dwk_acc=self._ngrad(ex, b, "attn.k_norm.weight", blk.attn.k_norm.weight) if ng else None)
        s = st["qkv"]
        self._linear_bwd(s, ws.dqkv, ws.xm1, ops.epi(ops.EPI_STORE, ws.g1), "qkv")                # d xm1
        tap(b, "attn_gin", ws.g1)
        ops.ln_mod_bwd(dx, ws.g1, x_in, scale_msa, dx_resid=dx, tokens_per_frame=tpf,
                       dscale_acc=dmod[:, C:2 * C] if want_mod else None,
                       dshift_acc=dmod[:, 0:C] if want_mod else None)
        tap(b, "dx_in", dx)
        if want_mod:
            ex.d_mod[b] = dmod.clone()
            if ex.need_dt:
                ada = blk.adaLN_modulation[1]
                ops.skinny_linear_bwd(ws.dt, dmod, self._t_for_block(b, ex), ada.weight, act=1)
                ex.d_t[b] = ws.dt.clone()
            if self.full is not None:
                self._modulation_grads(blk.adaLN_modulation[1], self._t_for_block(b, ex), dmod)
                ws.dt_total.add_(ws.dt)

    def _tap(self, b: int, name: str, t: torch.Tensor):
        """debug hook: engine.debug = {} collects clones of backward intermediates (tests only)"""
        dbg = getattr(self, "debug", None)
        if dbg is not None:
            dbg[(b, name)] = t.detach().float().clone()

    @staticmethod
    def _ngrad(ex: Extras, b: int, name: str, like: torch.Tensor):
        key = f"blocks.{b}.{name}"
        if key not in ex.d_norm:
            ex.d_norm[key] = torch.zeros(like.numel(), dtype=F32, device=like.device)
        return ex.d_norm[key]

    # ------------------------------------------------------------------ whole network
    def _prepare(self, geo: Geometry, ex: Optional[Extras]):
        self.resolve_sites()
        self.plan(geo)
        for s in self.lora_sites():
            s.refresh()
        if ex is not None and ex.need_dmod and self.ws.branch_a is None:
            self.ws.branch_a = torch.empty(geo.N, self.C, dtype=BF16, device=self.device)
            self.ws.branch_m = torch.empty(geo.N, self.C, dtype=BF16, device=self.device)

    def forward_tokens(self, text_valid: torch.Tensor, ex: Optional[Extras] = None, stash: bool = False,
                       ctx: Optional[str] = None) -> torch.Tensor:
        """ws.P / ws.timestep must hold the patchified input and the per-frame timestep.  Returns ws.pred [N,64] f32
        (final-layer token layout) and leaves the block inputs in ws.xs for the backward.  ``stash``: a backward
        follows -- keep per-block activations in spare HBM (see _ensure_stash).  ``ctx``: "fill" stores every block's
        context K/V on the way, "use" runs the noised rows only against that cache (forward only, same video / text /
        adapter state as the fill pass -- the caller guarantees it; see _block_fwd_noise_rows)."""
        ws, geo, C = self.ws, self.geo, self.C
        self._stash_on = False
        if stash:
            self._ensure_stash(geo)
            self._stash_on = True
        if geo.Nc == 0 or geo.Nn == 0:
            ctx = None
        self._ctx_fill = False
        if ctx is not None and self.bsa is not None:
            ctx = None   # the context K/V cache is implemented for the dense attention path only
        if ctx is not None:
            if stash:
                raise ValueError("the context cache is for forward-only passes")
            cache = self._ensure_ctx_cache(geo)
            if ctx == "use" and not cache["valid"]:
                raise RuntimeError("context cache used before it was filled for this geometry")
            self._ctx_fill = ctx == "fill"
        pe = self.dit.x_embedder.proj
        ops.gemm(geo.N, C, [(ws.P, pe.weight.view(C, 64), 64, False, None)], ops.epi(ops.EPI_STORE, ws.xs[0], bias=pe.bias))
        self._embed_time()
        self._embed_text(text_valid)
        for b in range(self.L):
            with _Range(f"fwd.block{b}"):
                if ctx == "use":
                    self._block_fwd_noise_rows(b, ws.xs[b], ws.xs[b + 1], ex)
                else:
                    self._block_fwd(b, ws.xs[b], ws.xs[b + 1], ex)
        if ctx is not None:
            self._ctx["valid"] = True
            self._ctx_fill = False
        x_last = ws.xs[self.L]
        if ex is not None and ex.hidden_final is not None:
            x_last.add_(ex.hidden_final.to(BF16)[None, :])
        fl = self.dit.final_layer
        t_f = ws.t
        if ex is not None and ex.t_offset_final is not None:
            t_f = torch.add(ws.t, ex.t_offset_final.to(F32)[None, :], out=ws.t_blk)
        ops.skinny_linear(ws.modf, t_f, fl.adaLN_modulation[1].weight, fl.adaLN_modulation[1].bias, act=1)
        ops.ln_mod_fwd(ws.xf, x_last, ws.modf[:, C:], ws.modf[:, :C], tokens_per_frame=geo.tpf)
        ops.gemm(geo.N, 64, [(ws.xf, fl.linear.weight, C, False, None)], ops.epi(ops.EPI_STORE_F32, ws.pred, bias=fl.linear.bias))
        if ex is not None and ex.out_bias is not None:
            ws.pred.view(geo.N, 4, 16).add_(ex.out_bias.to(F32)[None, None, :])
        return ws.pred

    def backward_tokens(self, ex: Optional[Extras] = None, only_out_bias: bool = False):
        """ws.dpred [Nn, 64] bf16 holds d loss / d pred for the noise rows (context rows carry no loss)."""
        ws, geo, C = self.ws, self.geo, self.C
        fl = self.dit.final_layer
        if ex is not None and ex.out_bias is not None:
            ex.d_out_bias = ws.dpred.float().view(-1, 4, 16).sum((0, 1))
            if only_out_bias:
                return
        self.grad_flat.zero_()
        if ex is not None:
            for v in ex.d_norm.values():
                v.zero_()
        full = self.full
        if full is not None:
            if ex is None or not (ex.need_dmod and ex.need_dt and ex.norm_grads):
                raise RuntimeError("full-model gradients need Extras(need_dmod, need_dt, norm_grads) -- use TTAStepper(full=True)")
            full.flat.zero_()
            ws.dt_total.zero_()
            ws.d_y.zero_()
            # final linear: dW [64, C] = dpred^T xf (noised rows), db = colsum(dpred)
            self._wgrad(fl.linear.weight, fl.linear.bias, ws.dpred, ws.xf[geo.Nc:])
        # final layer: d xf = dpred W_lin ; context rows are zero
        ws.g1[: geo.Nc].zero_()
        ops.gemm(geo.Nn, C, [(ws.dpred, fl.linear.weight, 64, True, None)], ops.epi(ops.EPI_STORE, ws.g1[geo.Nc:]))
        want_mod = ex is not None and ex.need_dmod
        if want_mod:
            ws.dmodf.zero_()
        ops.ln_mod_bwd(ws.dx, ws.g1, ws.xs[self.L], ws.modf[:, C:], tokens_per_frame=geo.tpf,
                       dscale_acc=ws.dmodf[:, C:] if want_mod else None, dshift_acc=ws.dmodf[:, :C] if want_mod else None)
        if want_mod and ex.need_dt:
            t_f = ws.t
            if ex.t_offset_final is not None:
                t_f = torch.add(ws.t, ex.t_offset_final.to(F32)[None, :], out=ws.t_blk)
            ops.skinny_linear_bwd(ws.dt, ws.dmodf, t_f, fl.adaLN_modulation[1].weight, act=1)
            ex.d_t_final = ws.dt.clone()
            if full is not None:
                self._modulation_grads(fl.adaLN_modulation[1], t_f, ws.dmodf)
                ws.dt_total.add_(ws.dt)
        if ex is not None and ex.hidden_final is not None:
            ex.d_hidden_final = ws.dx.float().sum(0)
        for b in reversed(range(self.L)):
            if getattr(self, "_ws_holds", None) != b:  # the last block's intermediates are still in the workspace
                with _Range(f"recompute.block{b}"):
                    self._block_fwd(b, ws.xs[b], ws.g2, ex, recompute=True)   # block output discarded into g2
            with _Range(f"bwd.block{b}"):
                self._block_bwd(b, ws.xs[b], ex)
            if full is not None and self.on_block_grads is not None:
                self.on_block_grads(b)
        if full is not None:
            if self.on_block_grads is not None:
                self.on_block_grads(None)
            self._embedder_grads(ex)
        self._ws_holds = None
        self._stash_on = False

    def _embedder_grads(self, ex: Extras):
        """full-model TTA: what is left after the block loop -- norm weights, patch / timestep / text embedders.
        ws.dx holds d loss / d(patch-embedder output); the small fp32 pieces (T or M rows) are plain tensor algebra."""
        ws, geo, C, full, dit = self.ws, self.geo, self.C, self.full, self.dit
        named = full.named()
        for key, g in ex.d_norm.items():                       # pre_crs_attn_norm (w, b), q / k RMSNorm weights
            named[key].copy_(g.view(named[key].shape))
        # patch embedder: x0 = P W^T + b with W [C, 64] (Conv3d weight flattened in P's column order)
        pe = dit.x_embedder.proj
        gw = full.g(pe.weight).view(C, 64)
        ops.gemm(C, 64, [(ws.dx, ws.P, geo.N, True, None, True)], ops.epi(ops.EPI_STORE_F32, gw))
        ops.colsum(full.g(pe.bias), ws.dx)
        # timestep embedder: t = W2 silu(W0 f + b0) + b2, T rows, fp32
        te, dt = dit.t_embedder, ws.dt_total
        w0, w2 = te.mlp[0], te.mlp[2]
        full.g(w2.weight).copy_(dt.t() @ _silu(ws.th))
        full.g(w2.bias).copy_(dt.sum(0))
        sg = torch.sigmoid(ws.th)
        d_th = (dt @ w2.weight.float()) * (sg * (1.0 + ws.th * (1.0 - sg)))
        full.g(w0.weight).copy_(d_th.t() @ ws.tfeat)
        full.g(w0.bias).copy_(d_th.sum(0))
        # text embedder: y = (gelu_tanh(text W0^T + b0) W2^T + b2) * keep
        ye = dit.y_embedder.y_proj
        text = self._text_cache[4]
        d_y = ws.d_y
        if getattr(self, "_text_keep", None) is not None:
            d_y.mul_(self._text_keep)
        M = text.shape[0]
        self._wgrad(ye[2].weight, ye[2].bias, d_y, ws.y1)
        d_y1 = torch.empty(M, C, dtype=BF16, device=self.device)
        ops.lora_linear_bwd(d_y, ye[2].weight, ops.epi(ops.EPI_STORE, d_y1))
        pre = torch.empty(M, C, dtype=BF16, device=self.device)
        ops.gemm(M, C, [(text, ye[0].weight, text.shape[1], False, None)], ops.epi(ops.EPI_STORE, pre, bias=ye[0].bias))
        u = pre.float()
        k0, k1 = 0.7978845608028654, 0.044715
        th = torch.tanh(k0 * (u + k1 * u ** 3))
        dgelu = 0.5 * (1.0 + th) + 0.5 * u * (1.0 - th * th) * k0 * (1.0 + 3.0 * k1 * u * u)
        d_pre = (d_y1.float() * dgelu).to(BF16)
        self._wgrad(ye[0].weight, ye[0].bias, d_pre, text)

    # ------------------------------------------------------------------ step-level helpers
    def set_inputs(self, cond, target, noise, sigma):
        """cond [16,Tc,H,W], target/noise [16,Tt,H,W] bf16, sigma f32 [1] (device).  Fills ws.P / ws.V / ws.timestep."""
        ops.noise_patchify(self.ws.P, self.ws.V, self.ws.timestep, cond, target, noise, sigma)

    def loss_and_dpred(self, want_grad: bool = True, loss_scale: float = 1.0) -> torch.Tensor:
        ws, geo = self.ws, self.geo
        ws.loss.zero_()
        ops.mse_fwd_bwd(ws.loss, ws.dpred if want_grad else None, ws.pred[geo.Nc:], ws.V, loss_scale=loss_scale)
        return ws.loss


# ---------------------------------------------------------------------------------------------------- autograd entry
def _geometry_for(dit, hidden_states, n_cond, M) -> Geometry:
    B, Cin, T, Hl, Wl = hidden_states.shape
    if B != 1:
        raise NotImplementedError("batch size 1 only (as in every reference run: common.py:448, run_sweep.sbatch:8)")
    return Geometry(T=T, Hl=Hl, Wl=Wl, n_cond=int(n_cond), M=M)


class _DiTFunction(torch.autograd.Function):
    """pred = dit(hidden, timestep, text) with gradients for the adapter parameters only (LoRA tensors and, when an
    ``adapters`` wrapper is given, its delta / norm / FiLM trainables)."""

    @staticmethod
    def forward(ctx, dit, adapter, hidden_states, timestep, text_valid, n_cond, n_lora, *params):
        eng = dit.engine
        geo = _geometry_for(dit, hidden_states, n_cond, text_valid.shape[0])
        ex = adapter.build_extras() if adapter is not None else None
        eng._prepare(geo, ex)
        ws = eng.ws
        ops.noise_patchify(ws.P, None, None, hidden_states[0].to(BF16).contiguous(), None, None, None)
        ws.timestep.copy_(timestep.reshape(-1).to(BF16).float())  # the DiT re-casts the timestep to its dtype first
        eng.forward_tokens(text_valid, ex, stash=any(getattr(ctx, "needs_input_grad", ())))
        out = torch.empty(1, 16, geo.T, geo.Hl, geo.Wl, dtype=F32, device=eng.device)
        ops.unpatchify(out[0], ws.pred, geo.T, geo.Hl, geo.Wl)
        ctx.dit, ctx.geo, ctx.adapter, ctx.ex = dit, geo, adapter, ex
        ctx.n_lora, ctx.n_params = n_lora, len(params)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        eng, geo, ex = ctx.dit.engine, ctx.geo, ctx.ex
        if eng.geo != geo:
            raise RuntimeError("the engine workspace was re-planned between forward and backward")
        ws = eng.ws
        # d pred on context frames is dropped: the loss never touches them (common.py:485) -- checked here
        g = grad_out[0].contiguous().float()
        if geo.n_cond > 0 and bool((g[:, : geo.n_cond] != 0).any()):
            raise NotImplementedError("gradient flowing into context-frame predictions is not supported")
        ops.latent_to_tokens(ws.dpred, g, geo.T, geo.Hl, geo.Wl, geo.n_cond)
        eng.backward_tokens(ex)
        grads = []
        for s in eng.lora_sites():
            for p, gr in zip(s.params, s.param_grads()):
                grads.append(gr.to(p.dtype).contiguous() if p.requires_grad else None)
        assert len(grads) == ctx.n_lora
        if ctx.adapter is not None:
            for p, gr in zip(ctx.adapter.trainable(), ctx.adapter.grads_from(ex)):
                grads.append(gr.reshape(p.shape).to(p.dtype))
        assert len(grads) == ctx.n_params
        return (None, None, None, None, None, None, None, *grads)


class HookedModulationAdapter:
    """The adapter protocol (``trainable`` / ``build_extras`` / ``grads_from``) synthesised from forward hooks found on
    ``blocks[i].adaLN_modulation`` -- the seam the reference's ``FiLMAdapterWrapper.apply_to_dit`` uses
    (delta_experiment/scripts/run_film_tta.py:146-163: ``output + expand(correction)``), so that a B200DiT carrying those
    hooks behaves like upstream's when it is called (generation after TTA, or a loss.backward() through ``dit(...)``).

    Only ADDITIVE hooks can be folded into the fused adaLN kernel (the term is evaluated once per forward on a zero
    output, with an autograd graph to whatever parameters the hook closes over; the engine's d loss / d(adaLN output)
    is pushed back through that graph).  Anything else is refused loudly."""

    def __init__(self, dit):
        self.dit = dit
        self.hooks = [list(blk.adaLN_modulation._forward_hooks.values()) for blk in dit.blocks]
        self.params: List[nn.Parameter] = []
        for hs in self.hooks:
            for h in hs:
                for cell in getattr(h, "__closure__", None) or ():
                    try:
                        obj = cell.cell_contents
                    except ValueError:
                        continue
                    if isinstance(obj, nn.Parameter) and obj.requires_grad and not any(obj is p for p in self.params):
                        self.params.append(obj)
        self._terms: List[Optional[torch.Tensor]] = []

    @staticmethod
    def present(dit) -> bool:
        return any(blk.adaLN_modulation._forward_hooks for blk in dit.blocks)

    def trainable(self) -> List[nn.Parameter]:
        return self.params

    def _run(self, b: int, out: torch.Tensor) -> torch.Tensor:
        mod = self.dit.blocks[b].adaLN_modulation
        for h in self.hooks[b]:
            r = h(mod, (None,), out)
            if r is not None:
                out = r
        return out

    def build_extras(self) -> Extras:
        cfg = self.dit.config
        dev = self.dit.x_embedder.proj.weight.device
        ex = Extras(len(self.dit.blocks))
        self._terms = []
        zero = torch.zeros(1, 1, 6 * cfg.hidden_size, dtype=F32, device=dev)
        probe = torch.linspace(-1.0, 1.0, 6 * cfg.hidden_size, dtype=F32, device=dev).view(1, 1, -1)
        with torch.enable_grad():
            for b, hs in enumerate(self.hooks):
                if not hs:
                    self._terms.append(None)
                    continue
                term = self._run(b, zero)
                with torch.no_grad():
                    if term.shape != zero.shape or not torch.allclose(self._run(b, probe) - probe, term, atol=1e-6):
                        raise NotImplementedError(
                            f"blocks[{b}].adaLN_modulation carries a forward hook that is not `output + constant`: only "
                            "additive (FiLM-style) hooks can be folded into the fused adaLN kernel")
                ex.film[b] = term.detach().reshape(-1).float()
                self._terms.append(term if term.requires_grad else None)
        ex.need_dmod = any(t is not None for t in self._terms)
        return ex

    def grads_from(self, ex: Extras) -> List[torch.Tensor]:
        gs = [torch.zeros(p.shape, dtype=F32, device=p.device) for p in self.params]
        for b, term in enumerate(self._terms):
            if term is None:
                continue
            g = ex.d_mod[b].sum(0).view_as(term).to(term.dtype)
            for acc, gp in zip(gs, torch.autograd.grad(term, self.params, g, allow_unused=True)):
                if gp is not None:
                    acc += gp.float()
        return gs


def dit_forward_autograd(dit, hidden_states, timestep, encoder_hidden_states, encoder_attention_mask, num_cond_latents,
                         adapter=None):
    if not hidden_states.is_cuda:
        from ._lib import B200TTAError
        raise B200TTAError("B200DiT runs on a B200 only: inputs are on %s and there is no CPU fallback" % hidden_states.device)
    eng = dit.engine
    if adapter is None and HookedModulationAdapter.present(dit):
        adapter = HookedModulationAdapter(dit)
    text_valid = eng.pack_text(encoder_hidden_states, encoder_attention_mask)
    if timestep.dim() == 1:
        timestep = timestep.unsqueeze(1).expand(-1, hidden_states.shape[2])
    lora_params = eng.adapter_parameters()
    params = lora_params + (list(adapter.trainable()) if adapter is not None else [])
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return _DiTFunction.apply(dit, adapter, hidden_states, timestep, text_valid, num_cond_latents, len(lora_params), *params)
    with torch.no_grad():
        return _DiTFunction.forward(_NoCtx(), dit, adapter, hidden_states, timestep, text_valid, num_cond_latents,
                                    len(lora_params), *params)


class _NoCtx:
    pass
