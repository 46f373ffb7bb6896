"""Full-model TTA: every DiT parameter trains.

Same function names, arguments and return dicts as the reference:
  finetune_full_on_conditioning   lora_experiment/scripts/run_full_tta.py:95-219
  reset_dit_weights               lora_experiment/scripts/run_full_tta.py:222-227
  finetune_full_batch             lora_experiment/scripts/run_full_tta.py:230-303

The reference lets autograd build weight gradients for all 13.6 B parameters (bf16) and steps them with
``torch.optim.SGD(momentum=0)`` (default: no optimizer state) or ``AdamW``.  Here the same engine that runs the adapter
methods additionally forms, for every linear its backward walks through, dW = dY^T X (the tcgen05 GEMM in its
weight-gradient form: both operands read token-major in place, fp32 output) and db = colsum(dY), plus the norm, adaLN and
embedder gradients -- all into ONE flat fp32 buffer (54 GB at 13.6 B; the all-reduce payload under data parallelism) --
and a multi-tensor clip + SGD / AdamW steps the bf16 parameters (update computed in fp32, rounded once).
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .lora import _warmup_lr
from .stepper import TTAStepper

BF16 = torch.bfloat16


def _check_all_trainable(dit: nn.Module):
    params = [p for p in dit.parameters() if p.requires_grad]
    if not params:
        raise ValueError("No trainable parameters found. Did you unfreeze the model?")     # run_full_tta.py:126-128
    if len(params) != sum(1 for _ in dit.parameters()):
        raise NotImplementedError("full-model TTA trains EVERY parameter: partially frozen models are what the LoRA / "
                                  "norm-tune entry points are for")
    return params


def finetune_full_on_conditioning(dit: nn.Module, cond_latents, train_latents, prompt_embeds, prompt_mask, num_steps: int = 10,
                                  lr: float = 1e-5, warmup_steps: int = 2, weight_decay: float = 0.01,
                                  max_grad_norm: float = 1.0, device: str = "cuda", dtype: torch.dtype = BF16,
                                  early_stopper=None, train_latents_variants: Optional[List[Dict]] = None,
                                  optimizer_type: str = "sgd", process_group=None) -> Dict:
    """run_full_tta.py:95-219.  Returns {losses, train_time, es_check_time, early_stopping_info}."""
    _check_all_trainable(dit)
    stepper = TTAStepper(dit, full=True, optimizer="adamw" if optimizer_type == "adamw" else "sgd", eps=1e-8,
                         weight_decay=weight_decay, max_grad_norm=max_grad_norm, process_group=process_group)
    if train_latents_variants is None:
        train_latents_variants = [{"latents": train_latents, "name": "orig"}]
    dit.train()
    losses, es_check_time = [], 0.0
    train_start = time.time()
    for step in range(num_steps):
        cur_lr = _warmup_lr(lr, step, warmup_steps)
        vi = torch.randint(0, len(train_latents_variants), (1,)).item()            # run_full_tta.py:160 (CPU stream)
        step_train = train_latents_variants[vi]["latents"]
        sigma = torch.rand(step_train.shape[0], device=device, dtype=torch.float32) * (1.0 - 0.001) + 0.001   # common.py:458
        noise = torch.randn_like(step_train)                                                                  # common.py:462
        losses.append(stepper.step(cond_latents, step_train, prompt_embeds, prompt_mask, sigma, noise, cur_lr))
        if early_stopper is not None:
            t0 = time.time()
            should_stop, info = early_stopper.step(step + 1)
            es_check_time += time.time() - t0
            if should_stop:
                print(f"  Early stopping at step {step + 1}: {info}")
                break
    losses = [float(v) for v in torch.cat(losses).tolist()] if losses else []
    train_time = time.time() - train_start
    dit.eval()
    es_state = None
    if early_stopper is not None:
        def _restore_full(state_dict):                                             # run_full_tta.py:198-205
            for k, v in state_dict.items():
                mod = dit
                parts = k.split(".")
                for part in parts[:-1]:
                    mod = getattr(mod, part)
                getattr(mod, parts[-1]).data.copy_(v)
        early_stopper.restore(restore_fn=_restore_full)
        es_state = early_stopper.state
    dit.engine.disable_full_grads()
    return {"losses": losses, "train_time": train_time, "es_check_time": es_check_time, "early_stopping_info": es_state}


def reset_dit_weights(dit: nn.Module, base_state: Dict[str, torch.Tensor]):
    """run_full_tta.py:222-227: back to the pretrained weights before the next video."""
    with torch.no_grad():
        for name, param in dit.named_parameters():
            if name in base_state:
                param.data.copy_(base_state[name].to(param.device))


def finetune_full_batch(dit: nn.Module, batch_data: List[Dict], num_steps: int = 10, lr: float = 1e-5, warmup_steps: int = 2,
                        weight_decay: float = 0.01, max_grad_norm: float = 1.0, device: str = "cuda",
                        dtype: torch.dtype = BF16, optimizer_type: str = "sgd") -> Dict:
    """run_full_tta.py:230-303: round-robin (``step % K``) over K pre-encoded videos held on the host."""
    _check_all_trainable(dit)
    stepper = TTAStepper(dit, full=True, optimizer="adamw" if optimizer_type == "adamw" else "sgd", eps=1e-8,
                         weight_decay=weight_decay, max_grad_norm=max_grad_norm)
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
    dit.engine.disable_full_grads()
    return {"losses": losses, "train_time": time.time() - t0, "es_check_time": 0.0, "early_stopping_info": None}
