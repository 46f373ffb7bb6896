"""Adapted denoise loop in latent space: the part of ``generate_video_continuation``
(delta_experiment/scripts/common.py:566-611 -> upstream ``pipe.generate_vc``) that runs through the DiT.

The upstream pipeline is not vendored by the reference, so the loop below is OUR statement of a
flow-matching Euler sampler with classifier-free guidance, consistent with the training parametrisation of
common.py:458-488 (x_sigma = (1 - sigma) x0 + sigma eps, the network predicts v = eps - x0 = d x_sigma / d sigma):

    sigmas: 1 = s_0 > s_1 > ... > s_K = 0   (uniform, optionally time-shifted s' = a s / (1 + (a - 1) s))
    x <- eps;  for k: v = v_neg + g (v_pos - v_neg)  (or v_pos without a negative prompt);  x <- x + (s_{k+1} - s_k) v

VAE encode / decode, the text encoder and video I/O stay outside (SURVEY: out of scope).  Conditioning frames are clean
(timestep 0) and only attend to themselves, and they have no cross attention, so their K / V per block are computed ONCE
(first forward) and every later forward -- both guidance branches of all steps -- runs the generated rows only
(engine._block_fwd_noise_rows; upstream's ``use_kv_cache=True``).
"""
from __future__ import annotations

from typing import Optional

import torch

BF16, F32 = torch.bfloat16, torch.float32


def flow_match_sigmas(num_inference_steps: int, shift: float = 1.0, device="cpu") -> torch.Tensor:
    s = torch.linspace(1.0, 0.0, num_inference_steps + 1, dtype=torch.float64, device=device)
    if shift != 1.0:
        s = shift * s / (1.0 + (shift - 1.0) * s)
    return s.to(F32)


@torch.no_grad()
def denoise_latents(model, cond_latents: torch.Tensor, prompt_embeds: torch.Tensor, prompt_mask: Optional[torch.Tensor],
                    num_gen_latent_frames: int, negative_prompt_embeds: Optional[torch.Tensor] = None,
                    negative_prompt_mask: Optional[torch.Tensor] = None, num_inference_steps: int = 50,
                    guidance_scale: float = 4.0, shift: float = 1.0, generator: Optional[torch.Generator] = None,
                    init_noise: Optional[torch.Tensor] = None, use_kv_cache: bool = True) -> torch.Tensor:
    """model: a B200DiT (LoRA injected or not) or one of the adapters.* wrappers.  cond_latents [1,16,Tc,H,W];
    returns the generated latent frames [1,16,num_gen_latent_frames,H,W] (fp32)."""
    from .adapters import stepper_for_eval
    st = stepper_for_eval(model)
    if st is None:
        raise TypeError("denoise_latents needs a B200DiT or a longcat_video_tta_b200.adapters wrapper (no CPU / eager fallback)")
    dev = cond_latents.device
    B, Cl, Tc, Hl, Wl = cond_latents.shape
    shape = (B, Cl, num_gen_latent_frames, Hl, Wl)
    if init_noise is None:
        init_noise = torch.randn(shape, generator=generator, device=generator.device if generator is not None else dev)
    x = init_noise.to(device=dev, dtype=F32).clone()
    sigmas = flow_match_sigmas(num_inference_steps, shift)
    cfg = negative_prompt_embeds is not None and guidance_scale != 1.0
    cond_bf = cond_latents.to(BF16)
    filled = False
    for k in range(num_inference_steps):
        s, s_next = float(sigmas[k]), float(sigmas[k + 1])
        sigma = torch.tensor([s], device=dev, dtype=F32)
        x_bf = x.to(BF16)

        def velocity(pe, pm):
            nonlocal filled
            ctx = None
            if use_kv_cache and Tc > 0:
                ctx = "use" if filled else "fill"
                filled = True
            return st.predict_velocity(cond_bf, x_bf, pe, pm, sigma, ctx=ctx)

        v = velocity(prompt_embeds, prompt_mask)
        if cfg:
            v_neg = velocity(negative_prompt_embeds, negative_prompt_mask)
            v = v_neg + guidance_scale * (v - v_neg)
        x += (s_next - s) * v
    return x
