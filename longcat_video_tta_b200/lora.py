"""LoRA test-time adaptation -- same names, arguments and return values as the reference's
``lora_experiment/scripts/run_lora_tta.py`` (functions :104-634), executed by the sm_100a engine.

``LoRALinear`` / ``inject_lora_into_dit`` / ``get_lora_parameters`` / ``reset_lora_weights`` / ``save_lora_weights`` keep
the reference's module surgery, parameter order (per block: attn.qkv, attn.proj, cross_attn.q_linear,
cross_attn.kv_linear, cross_attn.proj, then ffn.w1/w2/w3; down before up) and checkpoint layout
(``lora_{i}.down`` / ``lora_{i}.up``).  ``finetune_lora_on_conditioning`` is the hot loop: it keeps the reference's
RNG draw order (one CPU ``torch.randint`` per step, then sigma, then eps) and replaces loss / backward / clip / AdamW
with one fused ``TTAStepper.step``.
"""
from __future__ import annotations

import math
import time
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .stepper import TTAStepper

BF16 = torch.bfloat16


class LoRALinear(nn.Module):
    """Low-rank adapter around a frozen ``nn.Linear`` (reference: run_lora_tta.py:224-260).
    out = original(x) + lora_up(lora_down(x)) * (alpha / rank); down: kaiming-uniform(a=sqrt 5), up: zeros."""

    def __init__(self, original: nn.Linear, rank: int = 8, alpha: float = 16.0, dropout: float = 0.0):
        super().__init__()
        self.original = original
        self.rank, self.alpha = rank, alpha
        self.scaling = alpha / rank
        self.lora_down = nn.Linear(original.in_features, rank, bias=False)
        self.lora_up = nn.Linear(rank, original.out_features, bias=False)
        self.dropout = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        nn.init.kaiming_uniform_(self.lora_down.weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_up.weight)

    @property
    def in_features(self):
        return self.original.in_features

    @property
    def out_features(self):
        return self.original.out_features

    def forward(self, x, *args, **kwargs):
        """Stand-alone (inference) use of one adapted linear through the fused kernel; inside the DiT the engine
        drives the same kernel with fused epilogues instead of calling this."""
        from . import ops
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).to(BF16).contiguous()
        y = torch.empty(x2.shape[0], self.out_features, dtype=BF16, device=x.device)
        r = self.rank
        if r % 8 == 0 and r <= 64:
            xa = torch.empty(x2.shape[0], r, dtype=BF16, device=x.device)
            ops.lora_linear_fwd(x2, self.original.weight, ops.epi(ops.EPI_STORE, y, bias=self.original.bias),
                                A=self.lora_down.weight.detach(), B=self.lora_up.weight.detach(), XA=xa, scale=self.scaling)
        else:
            raise NotImplementedError("stand-alone LoRALinear.forward needs rank % 8 == 0; use the DiT forward")
        return y.reshape(*shp[:-1], self.out_features)


class LoRAModule(nn.Module):
    """Builtin-style adapter (upstream ``LoRAModule`` as the reference uses it, run_lora_tta.py:132-135,175-181,
    201-209): ``lora_down`` Linear(in, n_seperate*r), ``lora_up`` Linear(r, out) or ``.blocks[i]`` Linear(r, out/n),
    ``multiplier``, ``alpha_scale = alpha / r``, ``use_lora``."""

    class _Up(nn.Module):
        def __init__(self, r, out, n):
            super().__init__()
            self.blocks = nn.ModuleList([nn.Linear(r, out // n, bias=False) for _ in range(n)])

    def __init__(self, name, org_module: nn.Linear, multiplier=1.0, lora_dim=4, alpha=1.0, n_seperate=1):
        super().__init__()
        self.lora_name, self.lora_dim = name, lora_dim
        self.lora_down = nn.Linear(org_module.in_features, n_seperate * lora_dim, bias=False)
        self.lora_up = (LoRAModule._Up(lora_dim, org_module.out_features, n_seperate) if n_seperate > 1
                        else nn.Linear(lora_dim, org_module.out_features, bias=False))
        self.multiplier, self.alpha_scale, self.use_lora = multiplier, alpha / lora_dim, True
        reset_builtin_lora_weights([self])


def _parse_target_blocks(target_blocks: str, num_blocks: int) -> Optional[set]:
    """'all' -> None | 'last_N' | 'i,j,k'  (run_lora_tta.py:263-283)."""
    spec = target_blocks.strip().lower()
    if spec == "all":
        return None
    if spec.startswith("last_"):
        n = int(spec.split("_", 1)[1])
        if n <= 0 or n > num_blocks:
            raise ValueError(f"last_{n} invalid for {num_blocks} blocks")
        return set(range(num_blocks - n, num_blocks))
    idx = {int(x.strip()) for x in spec.split(",")}
    for i in idx:
        if i < 0 or i >= num_blocks:
            raise ValueError(f"Block index {i} out of range [0, {num_blocks})")
    return idx


_SITES = (("attn", "qkv", "qkv", 3), ("attn", "proj", "proj", 1), ("cross_attn", "q_linear", "qkv", 1),
          ("cross_attn", "kv_linear", "qkv", 2), ("cross_attn", "proj", "proj", 1))


def _sites(dit, target_modules, target_ffn, target_blocks):
    chosen = _parse_target_blocks(target_blocks, len(dit.blocks))
    for i, blk in enumerate(dit.blocks):
        if chosen is not None and i not in chosen:
            continue
        for parent, name, group, n_sep in _SITES:
            if group in target_modules and hasattr(blk, parent) and hasattr(getattr(blk, parent), name):
                yield i, getattr(blk, parent), name, f"blocks.{i}.{parent}.{name}", n_sep
        if target_ffn and hasattr(blk, "ffn"):
            for name in ("w1", "w2", "w3"):
                if hasattr(blk.ffn, name):
                    yield i, blk.ffn, name, f"blocks.{i}.ffn.{name}", 1


def inject_lora_into_dit(dit, rank: int = 8, alpha: float = 16.0, dropout: float = 0.0, target_modules=("qkv", "proj"),
                         target_ffn: bool = False, target_blocks: str = "all") -> List[LoRALinear]:
    """run_lora_tta.py:286-382."""
    ref = next(dit.parameters())
    chosen = _parse_target_blocks(target_blocks, len(dit.blocks))
    print(f"  LoRA target blocks: {'all (%d)' % len(dit.blocks) if chosen is None else sorted(chosen)}")
    mods = []
    for _, parent, name, _, _ in _sites(dit, target_modules, target_ffn, target_blocks):
        orig = getattr(parent, name)
        if isinstance(orig, nn.Linear):
            m = LoRALinear(orig, rank=rank, alpha=alpha, dropout=dropout).to(device=ref.device, dtype=ref.dtype)
            setattr(parent, name, m)
            mods.append(m)
    return mods


def inject_builtin_lora_into_dit(dit, rank: int = 8, alpha: float = 16.0, target_modules=("qkv", "proj"),
                                 target_ffn: bool = False, target_blocks: str = "all") -> List[LoRAModule]:
    """run_lora_tta.py:104-170: fused qkv gets 3 independent (A_i, B_i), kv_linear 2."""
    ref = next(dit.parameters())
    mods = []
    for _, parent, name, full, n_sep in _sites(dit, target_modules, target_ffn, target_blocks):
        lin = getattr(parent, name)
        if not isinstance(lin, nn.Linear):
            continue
        lora = LoRAModule(full, lin, multiplier=1.0, lora_dim=rank, alpha=alpha, n_seperate=n_sep)
        lora = lora.to(device=ref.device, dtype=ref.dtype)
        object.__setattr__(lin, "_b200_lora", lora)  # the engine finds the adapter here (not a registered submodule)
        mods.append(lora)
    return mods


def unhook_builtin_lora(dit):
    for m in dit.modules():
        if getattr(m, "_b200_lora", None) is not None:
            object.__setattr__(m, "_b200_lora", None)


def get_lora_parameters(lora_modules) -> List[nn.Parameter]:
    params = []
    for m in lora_modules:
        params.extend(m.lora_down.parameters())
        params.extend(m.lora_up.parameters())
    return params


get_builtin_lora_parameters = get_lora_parameters


def count_lora_parameters(lora_modules) -> Dict[str, int]:
    n = sum(p.numel() for p in get_lora_parameters(lora_modules))
    total = sum(p.numel() for m in lora_modules for p in m.parameters())
    return {"total_lora": total, "trainable": n}


count_builtin_lora_parameters = count_lora_parameters


def reset_lora_weights(lora_modules):
    for m in lora_modules:
        nn.init.kaiming_uniform_(m.lora_down.weight, a=math.sqrt(5))
        nn.init.zeros_(m.lora_up.weight)


def reset_builtin_lora_weights(lora_modules):
    for m in lora_modules:
        nn.init.kaiming_uniform_(m.lora_down.weight, a=math.sqrt(5))
        for p in m.lora_up.parameters():
            nn.init.zeros_(p)


def save_lora_weights(lora_modules, path: str):
    state = {}
    for i, m in enumerate(lora_modules):
        state[f"lora_{i}.down"] = m.lora_down.weight.detach().cpu()
        state[f"lora_{i}.up"] = m.lora_up.weight.detach().cpu()
    torch.save(state, path)


def _check_param_list(dit, lora_params):
    """The fused stepper trains EVERY adapter the engine finds on the DiT, in injection order; the reference's optimizer
    trains whatever list it is handed (run_lora_tta.py:455-468).  Refuse a list that is not exactly that set instead of
    silently training (and early-stop snapshotting) different tensors."""
    eng = getattr(dit, "engine", None)
    if eng is None:
        return
    mine = eng.adapter_parameters()
    if len(mine) != len(lora_params) or any(a is not b for a, b in zip(mine, lora_params)):
        raise NotImplementedError(
            f"the parameter list handed in ({len(lora_params)} tensors) is not the adapter set injected into the DiT "
            f"({len(mine)} tensors, injection order): training a subset / re-ordered list is not supported by the fused step")
    flags = [bool(p.requires_grad) for p in lora_params]
    if any(flags) and not all(flags):
        raise NotImplementedError("some adapter tensors are frozen and some are not: the fused step trains all of them")


def _warmup_lr(lr, step, warmup_steps):
    return lr * (step + 1) / warmup_steps if (warmup_steps > 0 and step < warmup_steps) else lr


def finetune_lora_on_conditioning(dit, lora_modules, cond_latents, train_latents, prompt_embeds, prompt_mask,
                                  num_steps: int = 20, lr: float = 2e-4, warmup_steps: int = 3,
                                  weight_decay: float = 0.01, max_grad_norm: float = 1.0, device: str = "cuda",
                                  dtype: torch.dtype = BF16, early_stopper=None, lora_param_fn=None,
                                  train_latents_variants: Optional[List[Dict]] = None, *, master_weights: bool = True,
                                  faithful_bf16: bool = False, process_group=None) -> Dict:
    """run_lora_tta.py:425-547.  Returns {losses, train_time, es_check_time, early_stopping_info}."""
    lora_params = lora_param_fn() if lora_param_fn is not None else get_lora_parameters(lora_modules)
    if not lora_params:
        raise ValueError("No LoRA parameters found.")
    _check_param_list(dit, lora_params)
    stepper = TTAStepper(dit, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay, max_grad_norm=max_grad_norm,
                         master_weights=master_weights, faithful_bf16=faithful_bf16, process_group=process_group)
    if train_latents_variants is None:
        train_latents_variants = [{"latents": train_latents, "name": "orig"}]

    def _save_fn():
        return [p.data.clone() for p in lora_params]

    def _restore(snapshot):
        for p, saved in zip(lora_params, snapshot):
            p.data.copy_(saved)

    dit.train()
    losses, es_check_time = [], 0.0
    train_start = time.time()
    for step in range(num_steps):
        cur_lr = _warmup_lr(lr, step, warmup_steps)
        vi = torch.randint(0, len(train_latents_variants), (1,)).item()     # run_lora_tta.py:499 (CPU stream)
        step_train = train_latents_variants[vi]["latents"]
        B = step_train.shape[0]
        sigma = torch.rand(B, device=device, dtype=torch.float32) * (1.0 - 0.001) + 0.001   # common.py:458
        noise = torch.randn_like(step_train)                                               # common.py:462
        losses.append(stepper.step(cond_latents, step_train, prompt_embeds, prompt_mask, sigma, noise, cur_lr))
        if early_stopper is not None:
            t0 = time.time()
            should_stop, info = early_stopper.step(step + 1, save_fn=_save_fn)
            es_check_time += time.time() - t0
            if should_stop:
                print(f"  Early stopping at step {step + 1}: {info}")
                break
    losses = [float(v) for v in torch.cat(losses).tolist()] if losses else []   # one device->host read per video
    train_time = time.time() - train_start
    dit.eval()
    es_state = None
    if early_stopper is not None:
        early_stopper.restore(restore_fn=_restore)
        es_state = early_stopper.state
    return {"losses": losses, "train_time": train_time, "es_check_time": es_check_time, "early_stopping_info": es_state}


def finetune_lora_batch(dit, lora_modules, batch_data: List[Dict], num_steps: int = 20, lr: float = 2e-4,
                        warmup_steps: int = 3, weight_decay: float = 0.01, max_grad_norm: float = 1.0,
                        device: str = "cuda", dtype: torch.dtype = BF16, lora_param_fn=None, *,
                        master_weights: bool = True, faithful_bf16: bool = False) -> Dict:
    """run_lora_tta.py:558-634: round-robin over K pre-encoded videos held on the host (one video per step)."""
    lora_params = lora_param_fn() if lora_param_fn is not None else get_lora_parameters(lora_modules)
    _check_param_list(dit, lora_params)
    stepper = TTAStepper(dit, eps=1e-8, weight_decay=weight_decay, max_grad_norm=max_grad_norm,
                         master_weights=master_weights, faithful_bf16=faithful_bf16)
    dit.train()
    losses = []
    t0 = time.time()
    for step in range(num_steps):
        cur_lr = _warmup_lr(lr, step, warmup_steps)
        bd = batch_data[step % len(batch_data)]
        cond, train = bd["cond_latents"].to(device, non_blocking=True), bd["train_latents"].to(device, non_blocking=True)
        pe = bd["prompt_embeds"].to(device, non_blocking=True)
        pm = bd["prompt_mask"].to(device, non_blocking=True) if bd["prompt_mask"] is not None else None
        sigma = torch.rand(train.shape[0], device=device, dtype=torch.float32) * (1.0 - 0.001) + 0.001
        noise = torch.randn_like(train)
        losses.append(stepper.step(cond, train, pe, pm, sigma, noise, cur_lr))
    losses = [float(v) for v in torch.cat(losses).tolist()] if losses else []
    dit.eval()
    return {"losses": losses, "train_time": time.time() - t0, "es_check_time": 0.0, "early_stopping_info": None}
