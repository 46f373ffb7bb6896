"""Shared TTA library -- the functions of ``delta_experiment/scripts/common.py`` that sit on the hot path, with the
reference's names, argument meaning and return values:

  split_tta_latents                              common.py:1365-1401
  compute_flow_matching_loss                     common.py:274-343
  compute_flow_matching_loss_fixed               common.py:346-407
  compute_flow_matching_loss_conditioned         common.py:414-489
  compute_flow_matching_loss_conditioned_fixed   common.py:492-559

The differentiable variants call ``dit(...)`` (``B200DiT`` or one of our adapter wrappers), whose forward/backward is
the sm_100a engine behind ``torch.autograd``; the ``_fixed`` (no-grad anchor) variants use the fused forward +
MSE kernels directly.  ``forward_fn`` keeps the reference's seam: when given it is called instead of ``dit``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import os

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16


def _get_model_config(dit):
    if hasattr(dit, "config"):
        return dit.config
    if hasattr(dit, "dit") and hasattr(dit.dit, "config"):
        return dit.dit.config
    raise AttributeError(f"Cannot find config on {type(dit).__name__}.")


def split_tta_latents(latents: torch.Tensor, num_context_latents: int, holdout_fraction: float = 0.25
                      ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """(cond, train, val|None) along the latent-time axis; at least one non-context frame, val >= 1 frame unless that
    would leave no training frame."""
    T = latents.shape[2]
    t_cond = min(num_context_latents, T - 1)
    rest = T - t_cond
    t_val = max(1, int(rest * holdout_fraction))
    t_train = rest - t_val
    if t_train < 1:
        t_train, t_val = rest, 0
    cond = latents[:, :, :t_cond].contiguous()
    train = latents[:, :, t_cond:t_cond + t_train].contiguous()
    val = latents[:, :, t_cond + t_train:].contiguous() if t_val > 0 else None
    return cond, train, val


def estimate_tta_split_budget(tta_total_frames: int, tta_context_frames: int, holdout_fraction: float = 0.25,
                              vae_t_scale: int = 4) -> Dict[str, int]:
    """Latent split sizes ``split_tta_latents`` will produce for a pixel-frame budget (common.py:1493-1517): same keys."""
    def latent_len(n_pixel_frames):
        return 1 + (max(1, int(n_pixel_frames)) - 1) // int(vae_t_scale)
    t_total, t_ctx = latent_len(tta_total_frames), latent_len(tta_context_frames)
    t_cond = min(t_ctx, t_total - 1)
    rest = t_total - t_cond
    t_val = max(1, int(rest * float(holdout_fraction)))
    t_train = rest - t_val
    if t_train < 1:
        t_train, t_val = rest, 0
    return {"total_latents": int(t_total), "cond_latents": int(t_cond), "train_latents": int(t_train),
            "val_latents": int(t_val)}


def _draw(target, device, sigma_min, sigma_max):
    B = target.shape[0]
    sigma = torch.rand(B, device=device, dtype=torch.float32) * (sigma_max - sigma_min) + sigma_min
    return sigma, torch.randn_like(target)


def compute_flow_matching_loss(dit, latents, prompt_embeds, prompt_mask, num_train_timesteps: int = 1000,
                               sigma_min: float = 0.001, sigma_max: float = 1.0, device: str = "cuda",
                               dtype: torch.dtype = BF16, forward_fn=None) -> torch.Tensor:
    cfg = _get_model_config(dit)
    B, _, T = latents.shape[:3]
    sigma, noise = _draw(latents, device, sigma_min, sigma_max)
    s = sigma.view(B, 1, 1, 1, 1)
    noisy = ((1.0 - s) * latents + s * noise).to(dtype)
    timestep = (sigma * num_train_timesteps).unsqueeze(1).expand(B, T // cfg.patch_size[0]).to(dtype)
    pred = forward_fn(noisy, timestep) if forward_fn is not None else dit(
        hidden_states=noisy, timestep=timestep, encoder_hidden_states=prompt_embeds, encoder_attention_mask=prompt_mask)
    return F.mse_loss(pred.to(torch.float32), (noise - latents).to(torch.float32))


def _conditioned_inputs(cfg, cond, target, sigma, noise, dtype, num_train_timesteps):
    B = cond.shape[0]
    pt = cfg.patch_size[0]
    n_cond, n_tgt = cond.shape[2] // pt, target.shape[2] // pt
    s = sigma.view(-1, 1, 1, 1, 1)
    hidden = torch.cat([cond, (1.0 - s) * target + s * noise], dim=2).to(dtype)
    timestep = torch.zeros(B, n_cond + n_tgt, device=cond.device, dtype=dtype)
    timestep[:, n_cond:] = (sigma * num_train_timesteps).reshape(-1, 1).expand(B, n_tgt).to(dtype)
    return hidden, timestep, n_cond


def compute_flow_matching_loss_conditioned(dit, cond_latents, target_latents, prompt_embeds, prompt_mask,
                                           num_train_timesteps: int = 1000, sigma_min: float = 0.001,
                                           sigma_max: float = 1.0, device: str = "cuda", dtype: torch.dtype = BF16,
                                           forward_fn=None) -> torch.Tensor:
    cfg = _get_model_config(dit)
    sigma, noise = _draw(target_latents, device, sigma_min, sigma_max)
    hidden, timestep, n_cond = _conditioned_inputs(cfg, cond_latents, target_latents, sigma, noise, dtype, num_train_timesteps)
    pred = forward_fn(hidden, timestep, n_cond) if forward_fn is not None else dit(
        hidden_states=hidden, timestep=timestep, encoder_hidden_states=prompt_embeds,
        encoder_attention_mask=prompt_mask, num_cond_latents=n_cond)
    t_cond = cond_latents.shape[2]
    return F.mse_loss(pred[:, :, t_cond:].to(torch.float32), (noise - target_latents).to(torch.float32))


def _python_mean(vals: List[torch.Tensor]) -> float:
    """The reference accumulates ``loss.item()`` in a Python float and divides by the count (common.py:528-559); same
    arithmetic here, after ONE device -> host read of all the per-forward losses instead of one sync per forward."""
    if not vals:
        return 0.0
    total = 0.0
    for v in torch.cat(vals).tolist():
        total += v
    return total / len(vals)


def _fused_eval(dit, cond, target, prompt_embeds, prompt_mask, sigma, noise, ctx=None) -> Optional[torch.Tensor]:
    """Forward + MSE through the fused kernels when ``dit`` is a B200DiT or one of our wrappers; device scalar."""
    from .adapters import stepper_for_eval
    st = stepper_for_eval(dit)
    if st is None:
        return None
    return st.eval_loss(cond, target, prompt_embeds, prompt_mask, sigma, noise, ctx=ctx)


def compute_flow_matching_loss_conditioned_fixed(dit, cond_latents, target_latents, prompt_embeds, prompt_mask,
                                                 fixed_sigmas: List[float], fixed_noises: List[torch.Tensor],
                                                 num_train_timesteps: int = 1000, device: str = "cuda",
                                                 dtype: torch.dtype = BF16, forward_fn=None) -> float:
    """Mean anchor loss over sigmas x noises (no grad).  One device->host read for the whole grid.  On the fused path
    the conditioning frames are computed once per call: the first forward fills the per-block context K/V cache, the
    others run the noised rows only (B200TTA_NO_CTX_CACHE=1 disables it)."""
    cfg = _get_model_config(dit)
    vals = []
    use_cache = os.environ.get("B200TTA_NO_CTX_CACHE") is None and len(fixed_sigmas) * len(fixed_noises) > 1
    for sv in fixed_sigmas:
        sigma = torch.tensor([sv], device=device, dtype=torch.float32)
        for noise in fixed_noises:
            v = None
            if forward_fn is None:
                ctx = None if not use_cache else ("fill" if not vals else "use")
                v = _fused_eval(dit, cond_latents, target_latents, prompt_embeds, prompt_mask, sigma, noise, ctx=ctx)
            if v is None:
                with torch.no_grad():
                    hidden, timestep, n_cond = _conditioned_inputs(cfg, cond_latents, target_latents, sigma, noise, dtype,
                                                                   num_train_timesteps)
                    pred = forward_fn(hidden, timestep, n_cond) if forward_fn is not None else dit(
                        hidden_states=hidden, timestep=timestep, encoder_hidden_states=prompt_embeds,
                        encoder_attention_mask=prompt_mask, num_cond_latents=n_cond)
                    v = F.mse_loss(pred[:, :, cond_latents.shape[2]:].to(torch.float32),
                                   (noise - target_latents).to(torch.float32)).reshape(1)
            vals.append(v.reshape(1))
    return _python_mean(vals)


def compute_flow_matching_loss_fixed(dit, latents, prompt_embeds, prompt_mask, fixed_sigmas: List[float],
                                     noise_draws: int = 1, num_train_timesteps: int = 1000, device: str = "cuda",
                                     dtype: torch.dtype = BF16, forward_fn=None) -> float:
    cfg = _get_model_config(dit)
    B, _, T = latents.shape[:3]
    vals = []
    for sv in fixed_sigmas:
        sigma = torch.tensor([sv], device=device, dtype=torch.float32)
        timestep = (sigma * num_train_timesteps).unsqueeze(1).expand(B, T // cfg.patch_size[0]).to(dtype)
        for d in range(noise_draws):
            gen = torch.Generator(device=device)
            gen.manual_seed(42 + d)
            noise = torch.randn(latents.shape, generator=gen, device=device, dtype=latents.dtype)
            noisy = ((1.0 - sigma.view(1, 1, 1, 1, 1)) * latents + sigma.view(1, 1, 1, 1, 1) * noise).to(dtype)
            with torch.no_grad():
                pred = forward_fn(noisy, timestep) if forward_fn is not None else dit(
                    hidden_states=noisy, timestep=timestep, encoder_hidden_states=prompt_embeds,
                    encoder_attention_mask=prompt_mask)
            vals.append(F.mse_loss(pred.to(torch.float32), (noise - latents).to(torch.float32)).reshape(1))
    return _python_mean(vals)
