"""VAE-side helpers of the reference that are the reference's own code (VAE half):
``encode_video`` / ``normalize_latents`` / ``denormalize_latents`` / ``decode_latents``
(delta_experiment/scripts/common.py:158-226), same names, arguments and return values.

The VAE network itself (upstream ``AutoencoderKLWan``, absent from the reference tree) stays whatever object the caller
loaded; what runs here is the per-channel latent normalisation around it -- ``b200tta_latent_affine``, bit-exact with the
reference's arithmetic (every intermediate rounded to the latent dtype).  No CPU path.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import ops


def _channel_stats(vae, latents: torch.Tensor):
    """mean and 1 / std per latent channel, rounded to the latent dtype exactly as common.py:177-187 builds them, then held
    as f32 on the latent's device for the kernel"""
    cfg = vae.config
    shape = (1, cfg.z_dim, 1, 1, 1)
    mean = torch.tensor(cfg.latents_mean).view(shape).to(latents.dtype)
    inv_std = 1.0 / torch.tensor(cfg.latents_std).view(shape).to(latents.dtype)
    if latents.dim() != 5 or latents.shape[1] != cfg.z_dim:
        raise ValueError(f"latents {tuple(latents.shape)}: expected [B, {cfg.z_dim}, T, H, W]")
    return (mean.float().reshape(-1).to(latents.device), inv_std.float().reshape(-1).to(latents.device))


def normalize_latents(vae, latents: torch.Tensor) -> torch.Tensor:
    """common.py:175-189: (latents - mean) * (1 / std), per channel"""
    mean, inv_std = _channel_stats(vae, latents)
    x = latents.contiguous()
    out = torch.empty_like(x)
    ops.latent_affine(out, x, mean, inv_std, inverse=False)
    return out


def denormalize_latents(vae, latents: torch.Tensor) -> torch.Tensor:
    """common.py:192-205: latents / (1 / std) + mean, per channel"""
    mean, inv_std = _channel_stats(vae, latents)
    x = latents.contiguous()
    out = torch.empty_like(x)
    ops.latent_affine(out, x, mean, inv_std, inverse=True)
    return out


def _default_retrieve(posterior):
    """what upstream's ``retrieve_latents`` (longcat_video.pipeline_longcat_video, absent here) does with its defaults on an
    encoder output: a sample of the posterior; falls back to ``.latents``.  Pass upstream's function to be exact."""
    if hasattr(posterior, "latent_dist"):
        return posterior.latent_dist.sample()
    if hasattr(posterior, "latents"):
        return posterior.latents
    raise AttributeError("could not read latents from the VAE's encoder output; pass retrieve_latents=")


def encode_video(vae, pixel_frames: torch.Tensor, normalize: bool = True,
                 retrieve_latents: Optional[Callable] = None) -> torch.Tensor:
    """common.py:158-172: pixel frames [B, C, T, H, W] -> (normalised) VAE latents; ``vae`` is the caller's encoder"""
    with torch.no_grad():
        latents = (retrieve_latents or _default_retrieve)(vae.encode(pixel_frames))
    return normalize_latents(vae, latents) if normalize else latents


def decode_latents(vae, latents: torch.Tensor, denorm: bool = True) -> torch.Tensor:
    """common.py:208-221: latents -> pixel frames [B, C, T, H, W] in [0, 1]; ``vae`` is the caller's decoder"""
    if denorm:
        latents = denormalize_latents(vae, latents)
    with torch.no_grad():
        video = vae.decode(latents.to(vae.dtype), return_dict=False)[0]
    return ((video + 1.0) / 2.0).clamp(0, 1)
