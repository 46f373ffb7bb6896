"""delta / norm-tune / FiLM adapters ("AdaSteer" family) -- same class names, constructor arguments, parameter
attributes and optimisation-loop signatures as the reference:

  DeltaAWrapper, optimize_delta_a            delta_experiment/scripts/run_delta_a.py:88-305
  DeltaBWrapper, optimize_delta_b            delta_experiment/scripts/run_delta_b.py:99-421
  DeltaCWrapper, optimize_delta_c            delta_experiment/scripts/run_delta_c.py:82-245
  collect_norm_params, NormTuneForward,
  optimize_norm_params                       delta_experiment/scripts/run_norm_tune_tta.py:74-283
  FiLMAdapterWrapper, optimize_film_adapter  delta_experiment/scripts/run_film_tta.py:78-340

The reference re-implements the DiT forward inside every wrapper and lets autograd find the gradients.  Here a wrapper
only describes WHERE its trainables enter the network (``build_extras``) and how the engine's modulation-side
gradients map back onto them (``grads_from``): d loss / d(adaLN output) per block comes out of the fused LayerNorm /
gate kernels as token reductions, d loss / d(timestep embedding) is one small fp32 GEMM further.
``wrapper(...)`` stays differentiable (torch.autograd over the engine) so the reference's own loops also work on it;
``optimize_*`` run the fused stepper instead.
"""
from __future__ import annotations

import copy
import math
import time
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .engine import Extras
from .lora import _parse_target_blocks
from .stepper import TTAStepper

BF16, F32 = torch.bfloat16, torch.float32


class _AdapterBase(nn.Module):
    per_tensor_clip = False

    def __init__(self, dit: nn.Module):
        super().__init__()
        self.dit = dit
        for p in self.dit.parameters():
            p.requires_grad = False

    @property
    def config(self):
        return self.dit.config

    # -- to be provided by subclasses
    def build_extras(self) -> Extras:
        raise NotImplementedError

    def trainable(self) -> List[nn.Parameter]:
        raise NotImplementedError

    def grads_from(self, ex: Extras) -> List[torch.Tensor]:
        raise NotImplementedError

    def forward(self, hidden_states, timestep, encoder_hidden_states, encoder_attention_mask=None, num_cond_latents=0,
                **kwargs):
        from .engine import dit_forward_autograd
        return dit_forward_autograd(self.dit, hidden_states, timestep, encoder_hidden_states, encoder_attention_mask,
                                    num_cond_latents or 0, adapter=self)


def _sum_t(t: Optional[torch.Tensor]) -> torch.Tensor:
    return t.sum(0)


# ---------------------------------------------------------------------------------------------------- delta-A
class DeltaAWrapper(_AdapterBase):
    """One fp32 vector added to the timestep embedding seen by every block and the final layer."""

    def __init__(self, dit: nn.Module, adaln_tembed_dim: int = 512):
        super().__init__(dit)
        self.delta = nn.Parameter(torch.zeros(adaln_tembed_dim, device=next(dit.parameters()).device))
        self._gen_hook = None

    def apply_to_dit(self):
        delta = self.delta

        def _hook(_m, _i, output):
            return output + delta.unsqueeze(0).to(output.dtype)
        self._gen_hook = self.dit.t_embedder.register_forward_hook(_hook)

    def remove_from_dit(self):
        if self._gen_hook is not None:
            self._gen_hook.remove()
            self._gen_hook = None

    def trainable(self):
        return [self.delta]

    def build_extras(self):
        ex = Extras(len(self.dit.blocks))
        d = self.delta.detach()
        ex.t_offset = [d] * len(self.dit.blocks)
        ex.t_offset_final = d
        ex.need_dmod = ex.need_dt = True
        return ex

    def grads_from(self, ex):
        g = _sum_t(ex.d_t_final)
        for dt in ex.d_t:
            g = g + _sum_t(dt)
        return [g]


# ---------------------------------------------------------------------------------------------------- delta-B
class DeltaBWrapper(_AdapterBase):
    """Per-group vectors: ``delta_target='timestep'`` offsets the timestep embedding of the group's blocks,
    ``'hidden'`` adds a residual after them (+ ``delta_final`` before the final layer).  Clipped per tensor."""
    per_tensor_clip = True

    def __init__(self, dit, num_groups: int = 4, adaln_tembed_dim: int = 512, hidden_size: int = 4096,
                 delta_target: str = "timestep", delta_dim: Optional[int] = None, target_blocks: str = "all"):
        super().__init__(dit)
        dev = next(dit.parameters()).device
        self.num_groups, self.num_blocks, self.delta_target = num_groups, len(dit.blocks), delta_target
        self.target_block_indices = _parse_target_blocks(target_blocks, self.num_blocks)
        self._full_dim = adaln_tembed_dim if delta_target == "timestep" else hidden_size
        self._partial_dim = delta_dim if delta_dim is not None else self._full_dim
        self.deltas = nn.ParameterList([nn.Parameter(torch.zeros(self._partial_dim, device=dev)) for _ in range(num_groups)])
        self.delta_final = nn.Parameter(torch.zeros(delta_dim, device=dev)) if delta_target == "hidden" else None
        per = math.ceil(self.num_blocks / num_groups)
        self.block_to_group = [min(i // per, num_groups - 1) for i in range(self.num_blocks)]
        self._gen_hooks: list = []

    def _pad_delta(self, dv):
        return dv if dv.shape[0] >= self._full_dim else F.pad(dv, (0, self._full_dim - dv.shape[0]))

    def _active(self, i):
        return self.target_block_indices is None or i in self.target_block_indices

    def apply_to_dit(self):
        raise NotImplementedError("generation-time hooks for delta-B are applied through build_extras(); "
                                  "call the wrapper, not the bare DiT")

    def remove_from_dit(self):
        self._gen_hooks = []

    def trainable(self):
        ps = list(self.deltas.parameters())
        if self.delta_final is not None:
            ps.append(self.delta_final)
        return ps

    def build_extras(self):
        ex = Extras(self.num_blocks)
        for i in range(self.num_blocks):
            if not self._active(i):
                continue
            d = self._pad_delta(self.deltas[self.block_to_group[i]].detach())
            if self.delta_target == "timestep":
                ex.t_offset[i] = d
            else:
                ex.hidden[i] = d
        if self.delta_target == "timestep":
            ex.need_dmod = ex.need_dt = True
        elif self.delta_final is not None:
            ex.hidden_final = self._pad_delta(self.delta_final.detach())
        return ex

    def grads_from(self, ex):
        gs = [torch.zeros(self._partial_dim, dtype=F32, device=self.deltas[0].device) for _ in range(self.num_groups)]
        for i in range(self.num_blocks):
            if not self._active(i):
                continue
            src = _sum_t(ex.d_t[i]) if self.delta_target == "timestep" else ex.d_hidden[i]
            gs[self.block_to_group[i]] += src[: self._partial_dim]
        if self.delta_final is not None:
            gs.append(ex.d_hidden_final[: self.delta_final.shape[0]])
        return gs


# ---------------------------------------------------------------------------------------------------- delta-C
class DeltaCWrapper(_AdapterBase):
    """Per-channel bias on the prediction; its gradient is a reduction of d loss / d pred (no DiT backward)."""

    def __init__(self, dit, mode: str = "per_channel", out_channels: int = 16):
        super().__init__(dit)
        if mode != "per_channel":
            raise ValueError(f"Unknown mode: {mode}. Use 'per_channel'.")
        self.mode = mode
        self.delta_out = nn.Parameter(torch.zeros(out_channels, device=next(dit.parameters()).device))
        self._gen_hook = None

    def apply_to_dit(self):
        d = self.delta_out

        def _hook(_m, _i, output):
            return output + d.view(1, -1, 1, 1, 1).to(output.dtype)
        self._gen_hook = self.dit.register_forward_hook(_hook)

    def remove_from_dit(self):
        if self._gen_hook is not None:
            self._gen_hook.remove()
            self._gen_hook = None

    def trainable(self):
        return [self.delta_out]

    def build_extras(self):
        ex = Extras(len(self.dit.blocks))
        ex.out_bias = self.delta_out.detach()
        return ex

    def grads_from(self, ex):
        return [ex.d_out_bias]


# ---------------------------------------------------------------------------------------------------- norm tuning
_NORM_SITES = ("attn.q_norm", "attn.k_norm", "cross_attn.q_norm", "cross_attn.k_norm")


def collect_norm_params(dit: nn.Module, norm_target: str) -> List[nn.Parameter]:
    """run_norm_tune_tta.py:74-98 -- cross_attn_norm | qk_norm | all_norm."""
    params = []
    for blk in dit.blocks:
        if norm_target in ("cross_attn_norm", "all_norm"):
            n = blk.pre_crs_attn_norm
            if getattr(n, "weight", None) is not None:
                params.append(n.weight)
            if getattr(n, "bias", None) is not None:
                params.append(n.bias)
        if norm_target in ("qk_norm", "all_norm"):
            for name in _NORM_SITES:
                mod = blk
                for part in name.split("."):
                    mod = getattr(mod, part, None)
                    if mod is None:
                        break
                if mod is not None and getattr(mod, "weight", None) is not None:
                    params.append(mod.weight)
    return params


def snapshot_params(params):
    return [p.data.clone() for p in params]


def restore_params(params, snapshot):
    for p, s in zip(params, snapshot):
        p.data.copy_(s)


class NormTuneForward(_AdapterBase):
    """The DiT's own norm affine parameters are the trainables (whichever have requires_grad)."""

    def __init__(self, dit: nn.Module, also_tune_delta: bool = False, adaln_tembed_dim: int = 512):
        """``also_tune_delta``: the reference's ``--also-tune-delta`` (run_norm_tune_tta.py:380-390): one fp32 delta-A vector,
        added to the timestep embedding, joins the norm parameters as the LAST entry of the optimizer's list."""
        nn.Module.__init__(self)
        self.dit = dit   # NOTE: does not re-freeze: the caller has just unfrozen the norm parameters
        self.delta = nn.Parameter(torch.zeros(adaln_tembed_dim, device=next(dit.parameters()).device)) \
            if also_tune_delta else None

    def _named(self):
        out = []
        for b, blk in enumerate(self.dit.blocks):
            cands = [("pre_crs_attn_norm.weight", blk.pre_crs_attn_norm.weight), ("pre_crs_attn_norm.bias", blk.pre_crs_attn_norm.bias)]
            cands += [(f"{n}.weight", getattr(getattr(blk, n.split(".")[0]), n.split(".")[1]).weight) for n in _NORM_SITES]
            out += [(f"blocks.{b}.{k}", p) for k, p in cands if p is not None and p.requires_grad]
        return out

    def trainable(self):
        return [p for _, p in self._named()] + ([self.delta] if self.delta is not None else [])

    def build_extras(self):
        ex = Extras(len(self.dit.blocks))
        ex.norm_grads = True
        if self.delta is not None:      # as DeltaAWrapper.build_extras: every block and the final layer see t + delta
            d = self.delta.detach()
            ex.t_offset = [d] * len(self.dit.blocks)
            ex.t_offset_final = d
            ex.need_dmod = ex.need_dt = True
        return ex

    def grads_from(self, ex):
        grads = [ex.d_norm[k] for k, _ in self._named()]
        if self.delta is not None:
            g = _sum_t(ex.d_t_final)
            for dt in ex.d_t:
                g = g + _sum_t(dt)
            grads.append(g)
        return grads


# ---------------------------------------------------------------------------------------------------- FiLM
class FiLMAdapterWrapper(_AdapterBase):
    """Per-group additive corrections to the adaLN output [shift_msa|scale_msa|gate_msa|shift_mlp|scale_mlp|gate_mlp]."""

    def __init__(self, dit, num_groups: int = 4, hidden_size: int = 4096, film_mode: str = "full"):
        super().__init__(dit)
        dims = {"full": 6, "shift_scale": 4, "scale_only": 2}
        if film_mode not in dims:
            raise ValueError(f"Unknown film_mode: {film_mode}")
        self.num_groups, self.num_blocks = num_groups, len(dit.blocks)
        self.hidden_size, self.film_mode = hidden_size, film_mode
        self.correction_dim = dims[film_mode] * hidden_size
        dev = next(dit.parameters()).device
        self.corrections = nn.ParameterList([nn.Parameter(torch.zeros(self.correction_dim, device=dev)) for _ in range(num_groups)])
        self._hooks = []

    def _get_group_idx(self, block_idx: int) -> int:
        return block_idx * self.num_groups // self.num_blocks

    _SLOTS = {"full": [0, 1, 2, 3, 4, 5], "shift_scale": [0, 1, 3, 4], "scale_only": [1, 4]}

    def _expand_correction(self, corr):
        C = self.hidden_size
        if self.film_mode == "full":
            return corr
        full = torch.zeros(6 * C, device=corr.device, dtype=corr.dtype)
        for j, slot in enumerate(self._SLOTS[self.film_mode]):
            full[slot * C:(slot + 1) * C] = corr[j * C:(j + 1) * C]
        return full

    def _contract(self, g6):
        C = self.hidden_size
        return torch.cat([g6[slot * C:(slot + 1) * C] for slot in self._SLOTS[self.film_mode]])

    def apply_to_dit(self):
        """The reference installs forward hooks on every adaLN_modulation; here the corrections reach the engine through
        build_extras() whenever the WRAPPER is called, so this only records that the adapter is live."""
        self._hooks = [True]

    def remove_from_dit(self):
        self._hooks = []

    def reset_corrections(self):
        for c in self.corrections:
            c.data.zero_()

    def trainable(self):
        return list(self.corrections.parameters())

    def build_extras(self):
        ex = Extras(self.num_blocks)
        for b in range(self.num_blocks):
            ex.film[b] = self._expand_correction(self.corrections[self._get_group_idx(b)].detach())
        ex.need_dmod = True
        return ex

    def grads_from(self, ex):
        gs = [torch.zeros(self.correction_dim, dtype=F32, device=self.corrections[0].device) for _ in range(self.num_groups)]
        for b in range(self.num_blocks):
            gs[self._get_group_idx(b)] += self._contract(_sum_t(ex.d_mod[b]))
        return gs


# ---------------------------------------------------------------------------------------------------- loops
def stepper_for_eval(model) -> Optional[TTAStepper]:
    """A forward-only stepper for a B200DiT or one of the wrappers above (cached on the object)."""
    from .dit import B200DiT
    st = getattr(model, "_b200_eval_stepper", None)
    if st is not None:
        return st
    if isinstance(model, B200DiT):
        st = TTAStepper(model, train_lora=False, build_optimizer=False)
    elif isinstance(model, _AdapterBase) and isinstance(model.dit, B200DiT):
        st = TTAStepper(model.dit, adapter=model, train_lora=False, build_optimizer=False)
    else:
        return None
    object.__setattr__(model, "_b200_eval_stepper", st)
    return st


def _optimize(wrapper, params, save_fn, restore_fn, cond_latents, train_latents, prompt_embeds, prompt_mask, num_steps, lr,
              device, early_stopper, train_latents_variants, track_es_time=True):
    """Common body of optimize_delta_a/b/c, optimize_norm_params, optimize_film_adapter:
    AdamW(lr, betas=(0.9,0.999), eps=1e-15) (default weight decay 0.01), no warm-up, clip 1.0 (per tensor for delta-B);
    RNG order per step: CPU randint (variant pick), sigma, eps."""
    stepper = TTAStepper(wrapper.dit, adapter=wrapper, train_lora=False, eps=1e-15, weight_decay=0.01, max_grad_norm=1.0,
                         per_tensor_clip=wrapper.per_tensor_clip)
    if train_latents_variants is None:
        train_latents_variants = [{"latents": train_latents, "name": "orig"}]
    wrapper.train()
    losses, es_time = [], 0.0
    for step in range(num_steps):
        vi = torch.randint(0, len(train_latents_variants), (1,)).item()
        step_train = train_latents_variants[vi]["latents"]
        sigma = torch.rand(step_train.shape[0], device=device, dtype=F32) * (1.0 - 0.001) + 0.001
        noise = torch.randn_like(step_train)
        losses.append(stepper.step(cond_latents, step_train, prompt_embeds, prompt_mask, sigma, noise, lr))
        if early_stopper is not None:
            t0 = time.time()
            should_stop, info = early_stopper.step(step + 1, save_fn=save_fn)
            es_time += time.time() - t0
            if should_stop:
                print(f"  Early stopping at step {step + 1}: {info}")
                break
    losses = [float(v) for v in torch.cat(losses).tolist()] if losses else []
    es_state = None
    if early_stopper is not None:
        early_stopper.restore(restore_fn=restore_fn)
        es_state = early_stopper.state
    return losses, es_time, es_state


def optimize_delta_a(wrapper: DeltaAWrapper, cond_latents, train_latents, prompt_embeds, prompt_mask, num_steps: int = 20,
                     lr: float = 1e-3, device: str = "cuda", dtype: torch.dtype = BF16, early_stopper=None,
                     train_latents_variants: Optional[List[Dict]] = None) -> Dict:
    losses, es_t, es = _optimize(wrapper, [wrapper.delta], lambda: copy.deepcopy(wrapper.delta.data),
                                 lambda s: wrapper.delta.data.copy_(s), cond_latents, train_latents, prompt_embeds,
                                 prompt_mask, num_steps, lr, device, early_stopper, train_latents_variants)
    return {"losses": losses, "delta_norm": wrapper.delta.detach().norm().item(), "es_check_time": es_t,
            "early_stopping_info": es}


def _optimize_delta_a_batch(wrapper: DeltaAWrapper, batch_data: List[Dict], num_steps: int = 20, lr: float = 1e-3,
                            device: str = "cuda", dtype: torch.dtype = BF16) -> Dict:
    """run_delta_a.py:308-362: ONE shared delta vector trained round-robin (``step % K``) over K pre-encoded videos that
    stay on the host; no variants, no early stopping.  Per step the reference draws sigma, then eps, and nothing else."""
    stepper = TTAStepper(wrapper.dit, adapter=wrapper, train_lora=False, eps=1e-15, weight_decay=0.01, max_grad_norm=1.0,
                         per_tensor_clip=wrapper.per_tensor_clip)
    wrapper.train()
    losses = []
    for step in range(num_steps):
        bd = batch_data[step % len(batch_data)]
        cond, train = bd["cond_latents"].to(device, non_blocking=True), bd["train_latents"].to(device, non_blocking=True)
        pe = bd["prompt_embeds"].to(device, non_blocking=True)
        pm = bd["prompt_mask"].to(device, non_blocking=True) if bd["prompt_mask"] is not None else None
        sigma = torch.rand(train.shape[0], device=device, dtype=F32) * (1.0 - 0.001) + 0.001
        noise = torch.randn_like(train)
        losses.append(stepper.step(cond, train, pe, pm, sigma, noise, lr))
    losses = [float(v) for v in torch.cat(losses).tolist()] if losses else []
    return {"losses": losses, "delta_norm": wrapper.delta.detach().norm().item(), "es_check_time": 0.0,
            "early_stopping_info": None}


def optimize_delta_b(wrapper: DeltaBWrapper, cond_latents, train_latents, prompt_embeds, prompt_mask, num_steps: int = 20,
                     lr: float = 1e-3, device: str = "cuda", dtype: torch.dtype = BF16, early_stopper=None,
                     train_latents_variants: Optional[List[Dict]] = None) -> Dict:
    params = wrapper.trainable()

    def _restore(saved):
        for p, s in zip(params, saved):
            p.data.copy_(s)
    losses, es_t, es = _optimize(wrapper, params, lambda: [copy.deepcopy(p.data) for p in params], _restore, cond_latents,
                                 train_latents, prompt_embeds, prompt_mask, num_steps, lr, device, early_stopper,
                                 train_latents_variants)
    return {"losses": losses, "delta_norms": [p.detach().norm().item() for p in params], "es_check_time": es_t,
            "early_stopping_info": es}


def optimize_delta_c(wrapper: DeltaCWrapper, cond_latents, train_latents, prompt_embeds, prompt_mask, num_steps: int = 20,
                     lr: float = 1e-3, device: str = "cuda", dtype: torch.dtype = BF16, early_stopper=None,
                     train_latents_variants: Optional[List[Dict]] = None) -> Dict:
    losses, es_t, es = _optimize(wrapper, [wrapper.delta_out], lambda: copy.deepcopy(wrapper.delta_out.data),
                                 lambda s: wrapper.delta_out.data.copy_(s), cond_latents, train_latents, prompt_embeds,
                                 prompt_mask, num_steps, lr, device, early_stopper, train_latents_variants)
    return {"losses": losses, "delta_out_norm": wrapper.delta_out.detach().norm().item(),
            "delta_out_values": wrapper.delta_out.detach().cpu().tolist(), "es_check_time": es_t, "early_stopping_info": es}


def optimize_norm_params(wrapper: NormTuneForward, norm_params: List[nn.Parameter], cond_latents, train_latents,
                         prompt_embeds, prompt_mask, num_steps: int = 20, lr: float = 1e-3, device: str = "cuda",
                         dtype: torch.dtype = BF16, early_stopper=None,
                         train_latents_variants: Optional[List[Dict]] = None) -> Dict:
    for p in norm_params:
        p.requires_grad_(True)
    assert [id(p) for p in wrapper.trainable()] == [id(p) for p in norm_params], \
        "norm_params must be collect_norm_params(dit, target) of the wrapped DiT, in order"
    losses, _, es = _optimize(wrapper, norm_params, lambda: snapshot_params(norm_params),
                              lambda s: restore_params(norm_params, s), cond_latents, train_latents, prompt_embeds,
                              prompt_mask, num_steps, lr, device, early_stopper, train_latents_variants)
    wrapper.dit.eval()
    return {"losses": losses, "early_stopping_info": es}


def optimize_film_adapter(wrapper: FiLMAdapterWrapper, cond_latents, train_latents, prompt_embeds, prompt_mask,
                          num_steps: int = 20, lr: float = 1e-3, device: str = "cuda", dtype: torch.dtype = BF16,
                          early_stopper=None, train_latents_variants: Optional[List[Dict]] = None) -> Dict:
    params = wrapper.trainable()

    def _restore(snap):
        for p, s in zip(params, snap):
            p.data.copy_(s)
    losses, _, es = _optimize(wrapper, params, lambda: [p.data.clone() for p in params], _restore, cond_latents,
                              train_latents, prompt_embeds, prompt_mask, num_steps, lr, device, early_stopper,
                              train_latents_variants)
    wrapper.eval()
    return {"losses": losses, "correction_norm": sum(c.detach().norm().item() for c in wrapper.corrections),
            "early_stopping_info": es}
