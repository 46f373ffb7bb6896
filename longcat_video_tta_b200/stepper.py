"""TTAStepper: one fused flow-matching TTA update (noise -> forward -> adapter backward -> clip -> AdamW).

This is the hot loop body of ``finetune_lora_on_conditioning`` (lora_experiment/scripts/run_lora_tta.py:490-516)
and of the ``optimize_*`` loops of the delta / norm / FiLM methods, executed without autograd: the random draws stay
in PyTorch (bit-exact RNG stream, SURVEY 7 "RNG parity"), everything else is hand-written kernels.

Data-parallel mode (SURVEY 8e -- our extension, the reference is single-GPU): every rank holds the frozen backbone
and a replica of the adapters, processes its own (sigma, eps) draw, and the flat fp32 adapter-gradient buffer is
all-reduced once per step with NCCL; the 1/world scale is folded into the clip / AdamW kernels.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from . import ops
from .engine import Geometry, TTAEngine

BF16, F32 = torch.bfloat16, torch.float32


class ParamGroup:
    """A set of trainable tensors with one clip policy and its own device descriptor table."""

    def __init__(self, entries: List[dict], device, per_tensor_clip: bool = False):
        self.entries = entries
        self.per_tensor_clip = per_tensor_clip
        self.tl = ops.TensorList(entries, device) if entries else None


def _state_like(p: torch.Tensor, fp32: bool):
    return torch.zeros(p.shape, dtype=F32 if fp32 else p.dtype, device=p.device)


class TTAStepper:
    def __init__(self, dit, *, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 max_grad_norm: float = 1.0, per_tensor_clip: bool = False, master_weights: bool = True,
                 faithful_bf16: bool = False, adapter=None, train_lora: bool = True, build_optimizer: bool = True,
                 process_group=None, cuda_graph: Optional[bool] = None, full: bool = False, optimizer: str = "adamw"):
        """
        master_weights : keep an fp32 master copy (and fp32 Adam moments) of every bf16 adapter tensor.  The
                         reference keeps params and moments in bf16 (run_lora_tta.py:332) where updates below half an
                         ulp are lost; ``faithful_bf16=True, master_weights=False`` reproduces
                         that op-by-op rounding instead.
        full           : full-model TTA (lora_experiment/scripts/run_full_tta.py:95-219): EVERY parameter of the DiT trains.
                         The engine's backward then also forms the weight / bias / norm / embedder gradients (fp32, one
                         flat buffer = the all-reduce payload); parameters stay bf16 without fp32 masters (13.6 B masters
                         would not fit next to the gradients), the update is computed in fp32 and rounded once.
        optimizer      : "adamw" | "sgd" (torch.optim.SGD(momentum=0) semantics; the reference's default for full TTA)
        adapter        : one of adapters.{DeltaA,DeltaB,DeltaC,NormTuneForward,FiLMAdapter}Wrapper: supplies the non-LoRA
                         trainables (``trainable()``), where they enter the network (``build_extras()``) and their
                         fp32 gradients after the backward (``grads_from(extras)``).
        """
        self.dit = dit
        self.eng: TTAEngine = dit.engine
        self.cuda_graph = bool(os.environ.get("B200TTA_CUDA_GRAPH")) if cuda_graph is None else bool(cuda_graph)
        self.betas, self.eps, self.wd, self.max_norm = betas, eps, weight_decay, max_grad_norm
        self.faithful = faithful_bf16
        self.adapter = adapter
        self.extras = None
        self.extra_params = list(adapter.trainable()) if adapter is not None else []
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.step_count = 0
        dev = self.eng.device
        self.eng.resolve_sites()
        entries, self._staged = [], []
        if train_lora:
            for s in self.eng.lora_sites():
                grads = [s.dA_acc, s.dB_acc] if not s.staged else [None] * len(s.params)
                for i, p in enumerate(s.params):
                    use_master = master_weights and p.dtype == BF16
                    fp32_state = use_master or p.dtype == F32
                    e = dict(param=p.data, exp_avg=_state_like(p, fp32_state), exp_avg_sq=_state_like(p, fp32_state))
                    if use_master:
                        e["master"] = p.data.float()
                    if s.staged:
                        e["grad"] = torch.zeros(p.shape, dtype=F32, device=dev)
                    else:
                        e["grad"] = grads[i]
                        e["grad_transposed"] = i == 0  # dA is accumulated as [in, r]
                    entries.append(e)
                if s.staged:
                    self._staged.append((s, entries[-len(s.params):]))
        self.full = bool(full)
        self.optimizer = optimizer
        if optimizer not in ("adamw", "sgd"):
            raise ValueError(f"unknown optimizer {optimizer!r}")
        self._full_extras = None
        if self.full:
            if adapter is not None or self.eng.lora_sites():
                raise ValueError("full-model TTA runs on a plain DiT (no LoRA injection, no adapter wrapper)")
            fg = self.eng.enable_full_grads()
            entries = []
            for p in fg.params:
                e = dict(param=p.data, grad=fg.g(p))
                if optimizer == "adamw":   # states in the parameter dtype, as torch.optim.AdamW keeps them
                    e["exp_avg"], e["exp_avg_sq"] = torch.zeros_like(p.data), torch.zeros_like(p.data)
                else:
                    e["exp_avg"] = e["exp_avg_sq"] = None
                entries.append(e)
            # several ranks: all-reduce each block's 1.1 GB of fp32 gradients as soon as the block's backward is enqueued,
            # on NCCL's stream, under the remaining blocks' compute (B200TTA_OVERLAP_ALLREDUCE=0: one blocking
            # all-reduce of the whole 54 GB buffer after the backward)
            self._pending, self._blocks_reduced = [], False
            if self.world > 1 and os.environ.get("B200TTA_OVERLAP_ALLREDUCE", "1") != "0":
                self.eng.on_block_grads = self._reduce_block_grads
            else:
                self.eng.on_block_grads = None
        self._n_lora_entries = len(entries)
        self._extra_entries = []
        for p in self.extra_params:
            use_master = master_weights and p.dtype == BF16
            fp32_state = use_master or p.dtype == F32
            e = dict(param=p.data, grad=torch.zeros(p.shape, dtype=F32, device=dev),
                     exp_avg=_state_like(p, fp32_state), exp_avg_sq=_state_like(p, fp32_state))
            if use_master:
                e["master"] = p.data.float()
            self._extra_entries.append(e)
        if not build_optimizer:
            entries, self._extra_entries = [], []
        self.group = ParamGroup(entries + self._extra_entries, dev, per_tensor_clip)
        self.n_params = sum(e["param"].numel() for e in self.group.entries)

    # ------------------------------------------------------------------
    def _geometry(self, cond, target, text_valid) -> Geometry:
        B, _, Tc, Hl, Wl = cond.shape
        if B != 1:
            raise NotImplementedError("batch size 1 only (every reference run; common.py:448)")
        if cond.shape[1] != 16 or target.shape[:2] != cond.shape[:2] or target.shape[3:] != cond.shape[3:]:
            raise ValueError(f"cond {tuple(cond.shape)} / target {tuple(target.shape)}: expected [1,16,T,H,W] with equal H, W")
        if cond.device != self.eng.device or target.device != self.eng.device:
            raise ValueError(f"latents must live on {self.eng.device}")
        return Geometry(T=Tc + target.shape[2], Hl=Hl, Wl=Wl, n_cond=Tc, M=text_valid.shape[0])

    @staticmethod
    def _dense(x: torch.Tensor) -> torch.Tensor:
        """[16, T, H, W] bf16, densely laid out: the noising kernel reads latents by raw pointer (a frame slice of a
        longer latent, or randn_like of one, is a strided view -- `.to(bf16)` alone would hand its storage over as is).
        fp32 latents are rounded to bf16 here, before the (1 - s) x + s eps mix (the reference mixes in fp32 and rounds
        the result, common.py:463-466: a difference of one bf16 rounding on x0 and eps)."""
        return x[0].to(BF16).contiguous()

    # ------------------------------------------------------------------ CUDA graph of forward + backward
    # The ~6 800 kernel launches of one headline step (140 per block) go through ctypes one by one; with a fixed
    # geometry every pointer is static (engine._WS, the stash, grad_flat), so the whole forward + adapter backward can be
    # captured once and replayed: cuda_graph=True / B200TTA_CUDA_GRAPH=1.  The optimizer stays outside the graph (the
    # learning rate is a host argument of the AdamW kernel and changes during warm-up).  The first call with a new
    # (shapes, adapter set) runs eagerly (workspace planning, stash sizing and the library's one-time attribute calls are
    # not capturable), the second captures, later ones replay; inputs are copied into static buffers.
    def _graph_key(self, cond, target, prompt_embeds, prompt_mask):
        return (tuple(cond.shape), tuple(target.shape), tuple(prompt_embeds.shape),
                None if prompt_mask is None else tuple(prompt_mask.shape), len(self.eng.lora_sites()))

    def _forward_backward_graphed(self, cond, target, prompt_embeds, prompt_mask, sigma, noise) -> torch.Tensor:
        key = self._graph_key(cond, target, prompt_embeds, prompt_mask)
        st = self.__dict__.setdefault("_graph_state", {"key": None})
        if st["key"] != key:
            st.clear()
            st.update(key=key, calls=0)
        st["calls"] += 1
        if st["calls"] == 1:
            return self._forward_backward_eager(cond, target, prompt_embeds, prompt_mask, sigma, noise)
        if st["calls"] == 2:
            st["in"] = [t.clone() if t is not None else None for t in (cond, target, prompt_embeds, prompt_mask, sigma, noise)]
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                st["loss"] = self._forward_backward_eager(*st["in"])
            st["graph"] = g
        else:
            for dst, src in zip(st["in"], (cond, target, prompt_embeds, prompt_mask, sigma, noise)):
                if dst is not None:
                    dst.copy_(src)
        st["graph"].replay()
        return st["loss"]

    def forward_backward(self, cond, target, prompt_embeds, prompt_mask, sigma, noise) -> torch.Tensor:
        """Loss (device scalar, f32) and adapter gradients for explicit (sigma, eps)."""
        if self.cuda_graph and self.adapter is None and self.eng.bsa is None and not self.full:
            return self._forward_backward_graphed(cond, target, prompt_embeds, prompt_mask, sigma, noise)
        return self._forward_backward_eager(cond, target, prompt_embeds, prompt_mask, sigma, noise)

    @torch.no_grad()      # the engine IS the backward: nothing here may be recorded by autograd (parameters may require grad)
    def _forward_backward_eager(self, cond, target, prompt_embeds, prompt_mask, sigma, noise) -> torch.Tensor:
        eng = self.eng
        text_valid = eng.pack_text(prompt_embeds, prompt_mask)
        geo = self._geometry(cond, target, text_valid)
        ex = self.extras = self.adapter.build_extras() if self.adapter is not None else None
        if self.full:
            from .engine import Extras
            ex = self.extras = Extras(eng.L)
            ex.need_dmod = ex.need_dt = ex.norm_grads = True
        eng._prepare(geo, ex)
        if noise.shape != target.shape:
            raise ValueError(f"noise {tuple(noise.shape)} must match the target latents {tuple(target.shape)}")
        eng.set_inputs(self._dense(cond), self._dense(target), self._dense(noise), sigma.to(F32))
        eng.forward_tokens(text_valid, ex, stash=True)
        loss = eng.loss_and_dpred(True)
        only_bias = ex is not None and ex.out_bias is not None and not self._needs_dit_backward()
        eng.backward_tokens(ex, only_out_bias=only_bias)
        return loss

    def _needs_dit_backward(self) -> bool:
        if self._n_lora_entries > 0:
            return True
        ex = self.extras
        if ex is None:
            return False
        return (ex.need_dmod or ex.norm_grads or any(h is not None for h in ex.hidden) or ex.hidden_final is not None)

    def _reduce_block_grads(self, b: Optional[int]):
        import torch.distributed as dist
        if b is None:      # the late pieces are about to be written into ranges that are being reduced: order them behind
            for w in self._pending:
                w.wait()   # stream-level wait, the host does not block
            self._pending = []
            self._blocks_reduced = True
            return
        self._blocks_reduced = False       # (a backward whose gradients were never applied leaves nothing behind)
        lo, hi = self.eng.full.block_ranges[b]
        self._pending.append(dist.all_reduce(self.eng.full.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def _sync_full_grads_overlapped(self):
        """what the per-block all-reduces have not covered: the ranges before / after the blocks (embedders, final layer)
        and the norm weights inside the block ranges, which the engine writes after the block loop"""
        import torch.distributed as dist
        fg = self.eng.full
        flat, ranges = fg.flat, fg.block_ranges
        for lo, hi in ((0, ranges[0][0]), (ranges[-1][1], flat.numel())):
            if hi > lo:
                dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
        named = fg.named()
        late = [named[k] for k in self.extras.d_norm if k.startswith("blocks.")]
        if late:
            buf = torch.cat([t.reshape(-1) for t in late])
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg)
            off = 0
            for t in late:
                t.copy_(buf[off: off + t.numel()].view(t.shape))
                off += t.numel()
        self._blocks_reduced = False

    def _sync_grads(self):
        if self.world > 1:
            from .dist import all_reduce_grads
            if self.full and getattr(self, "_blocks_reduced", False):
                return self._sync_full_grads_overlapped()
            if self.full:
                bufs = [self.eng.full.flat]
            else:
                bufs = ([self.eng.grad_flat] if self._n_lora_entries else []) + [e["grad"] for e in self._extra_entries]
            all_reduce_grads(bufs, self.pg)

    def optimizer_step(self, lr: float):
        self.step_count += 1
        if self.adapter is not None and self._extra_entries:
            for e, g in zip(self._extra_entries, self.adapter.grads_from(self.extras)):
                e["grad"].copy_(g.reshape(e["grad"].shape))
        self._sync_grads()
        for s, ents in self._staged:
            for e, g in zip(ents, s.param_grads()):
                e["grad"].copy_(g)
        tl = self.group.tl
        gs = 1.0 / self.world
        if self.max_norm is not None and self.max_norm > 0:
            tl.clip_coef(self.max_norm, per_tensor=self.group.per_tensor_clip, grad_scale=gs)
        use_coef = self.max_norm is not None and self.max_norm > 0
        if self.optimizer == "sgd":
            tl.sgd(lr=lr, weight_decay=self.wd, grad_scale=gs, use_coef=use_coef)
        else:
            tl.adamw(lr=lr, betas=self.betas, eps=self.eps, weight_decay=self.wd, step=self.step_count, grad_scale=gs,
                     use_coef=use_coef, faithful_bf16=self.faithful)

    def step(self, cond, target, prompt_embeds, prompt_mask, sigma, noise, lr: float) -> torch.Tensor:
        loss = self.forward_backward(cond, target, prompt_embeds, prompt_mask, sigma, noise)
        out = loss.clone()
        self.optimizer_step(lr)
        return out

    # ------------------------------------------------------------------ forward-only (anchor loss, common.py:492-559)
    @torch.no_grad()
    def eval_loss(self, cond, target, prompt_embeds, prompt_mask, sigma, noise, ctx: Optional[str] = None) -> torch.Tensor:
        """ctx: None | "fill" | "use" -- context K/V cache across forwards that share video, text and adapter state
        (engine._block_fwd_noise_rows)."""
        eng = self.eng
        text_valid = eng.pack_text(prompt_embeds, prompt_mask)
        geo = self._geometry(cond, target, text_valid)
        ex = self.adapter.build_extras() if self.adapter is not None else None
        eng._prepare(geo, ex)
        eng.set_inputs(self._dense(cond), self._dense(target), self._dense(noise), sigma.reshape(-1)[:1].to(F32))
        eng.forward_tokens(text_valid, ex, ctx=ctx)
        eng._ws_holds = None
        return eng.loss_and_dpred(False).clone()

    @torch.no_grad()
    def predict_velocity(self, cond, x_t, prompt_embeds, prompt_mask, sigma, ctx: Optional[str] = None) -> torch.Tensor:
        """Forward only: the predicted velocity of the non-conditioning frames, [1,16,T_gen,H,W] fp32, for an already
        noised latent ``x_t`` at noise level ``sigma`` (denoise loop; common.py:566-611 via upstream generate_vc)."""
        eng = self.eng
        text_valid = eng.pack_text(prompt_embeds, prompt_mask)
        geo = self._geometry(cond, x_t, text_valid)
        ex = self.adapter.build_extras() if self.adapter is not None else None
        eng._prepare(geo, ex)
        xt = self._dense(x_t)
        eng.set_inputs(self._dense(cond), xt, xt, sigma.reshape(-1)[:1].to(F32))   # (1 - s) x + s x = x: no re-noising
        eng.forward_tokens(text_valid, ex, ctx=ctx)
        eng._ws_holds = None
        out = torch.empty(16, geo.T, geo.Hl, geo.Wl, dtype=F32, device=eng.device)
        ops.unpatchify(out, eng.ws.pred, geo.T, geo.Hl, geo.Wl)
        return out[:, geo.n_cond:].unsqueeze(0).clone()

    def total_grad_norm(self) -> torch.Tensor:
        return self.group.tl.total_norm
