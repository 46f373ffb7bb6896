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


def resolve_tta_frames(args) -> None:
    """The block every method script runs after parsing (run_lora_tta.py:743-757): the TTA window defaults to the
    conditioning frames, a missing or too large context does too, and the window never reaches past
    ``gen_start_frame`` (no ground-truth leakage).  Resolves ``args.tta_total_frames`` / ``args.tta_context_frames``
    in place."""
    if args.tta_total_frames is None:
        args.tta_total_frames = args.num_cond_frames
    if args.tta_context_frames is None or args.tta_context_frames > args.tta_total_frames:
        args.tta_context_frames = args.num_cond_frames
    if args.tta_total_frames > args.gen_start_frame:
        print(f"[WARN] tta_total_frames ({args.tta_total_frames}) exceeds gen_start_frame ({args.gen_start_frame}); "
              f"clamping to avoid GT leakage.")
        args.tta_total_frames = args.gen_start_frame
    args.tta_context_frames = min(args.tta_context_frames, args.tta_total_frames)


def validate_tta_feature_budget(args, context: str = "") -> Dict:
    """Does the resolved frame budget leave room for what is switched on (common.py:1520-1598)?  Early stopping needs at
    least one held-out latent frame; the CLIP gate (recorded only in this build, checked all the same so that a sweep
    fails where the reference's would) needs enough candidate frames.  ``feature_frame_guard_mode``: fail | warn | off."""
    mode = str(getattr(args, "feature_frame_guard_mode", "fail")).lower()
    mode = mode if mode in ("fail", "warn", "off") else "fail"
    tag = f"[feature_budget:{context}]" if context else "[feature_budget]"
    total = int(getattr(args, "tta_total_frames", 0) or 0)
    ctx = int(getattr(args, "tta_context_frames", 0) or 0)
    holdout = float(getattr(args, "es_holdout_fraction", 0.25) or 0.25)
    split = estimate_tta_split_budget(total, ctx, holdout_fraction=holdout)
    info: Dict = {"split_budget": split}
    issues = []
    if not getattr(args, "es_disable", False) and split["val_latents"] < 1:
        issues.append(f"ES is enabled but estimated val_latents=0 (tta_total_frames={total}, tta_context_frames={ctx}, "
                      f"holdout={holdout}). Increase tta_total_frames and/or reduce tta_context_frames.")
    if getattr(args, "clip_gate_enabled", False):
        sampling = "late_only" if getattr(args, "clip_gate_late_only", False) else \
            str(getattr(args, "clip_gate_sampling_mode", "full_window"))
        late = float(getattr(args, "clip_gate_late_fraction", 0.4) or 0.4)
        backend = str(getattr(args, "clip_gate_backend", "clip") or "clip").lower()
        need = 8 if backend == "xclip" else int(getattr(args, "clip_gate_sample_frames", 4) or 4)
        window = max(1, total)
        have = max(1, int(round(window * min(max(late, 1e-6), 1.0)))) if (sampling or "full_window").lower() == "late_only" \
            else window
        info.update(clip_candidates=int(have), clip_required_frames=int(need))
        if have < need:
            issues.append(f"CLIP gate is enabled but candidate frames are fewer than required (candidates={have}, "
                          f"required={need}, tta_total_frames={total}, sampling_mode={sampling}, late_fraction={late}). "
                          f"Increase tta_total_frames and/or adjust sampling.")
    if mode != "off":
        print(f"{tag} split(total={split['total_latents']}, cond={split['cond_latents']}, train={split['train_latents']}, "
              f"val={split['val_latents']})")
        if "clip_candidates" in info:
            print(f"{tag} clip_candidates={info['clip_candidates']} required={info['clip_required_frames']}")
    if issues and mode != "off":
        msg = f"{tag} " + " | ".join(issues)
        if mode == "fail":
            raise RuntimeError(msg)
        print(f"WARNING: {msg}")
    return info


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
