"""Adapted denoise loop at the headline geometry (13.6B, 480p, 4 conditioning + 20 generated latent frames, LoRA r=16
injected, classifier-free guidance): seconds per Euler step (two forwards) with and without the context K/V cache."""
import sys, time, torch
sys.path.insert(0, '.')
from longcat_video_tta_b200 import lora
from longcat_video_tta_b200.denoise import denoise_latents
from longcat_video_tta_b200.dit import B200DiT
BF16 = torch.bfloat16
dev = torch.device("cuda", 0)
dit = B200DiT.random_init("13.6b", seed=0, device=dev)
torch.manual_seed(7)
lora.inject_lora_into_dit(dit, rank=16, alpha=32.0, target_modules=["qkv", "proj"])
g = torch.Generator().manual_seed(1)
cond = torch.randn(1, 16, 4, 60, 104, generator=g).to(BF16).to(dev)
prompt = torch.randn(1, 1, 512, dit.config.caption_channels, generator=g).to(BF16).to(dev)
neg = torch.randn(1, 1, 512, dit.config.caption_channels, generator=g).to(BF16).to(dev)
mask = torch.ones(1, 512, dtype=torch.int64, device=dev)
noise = torch.randn(1, 16, 20, 60, 104, generator=g)
for label, cache in (("context K/V cache", True), ("full forwards", False)):
    kw = dict(negative_prompt_embeds=neg, negative_prompt_mask=mask, guidance_scale=4.0, init_noise=noise, use_kv_cache=cache)
    denoise_latents(dit, cond, prompt, mask, 20, num_inference_steps=1, **kw)   # warm-up
    torch.cuda.synchronize(); t0 = time.time()
    steps = 4
    out = denoise_latents(dit, cond, prompt, mask, 20, num_inference_steps=steps, **kw)
    torch.cuda.synchronize(); dt = (time.time() - t0) / steps
    print(f"{label:20s} {dt:6.3f} s per Euler step (2 forwards, 37 440 tokens)  -> {50 * dt:6.1f} s per 50-step video; "
          f"latent rms {out.float().pow(2).mean().sqrt().item():.4f}", flush=True)
