"""Oracle restatement of the reference's TTA step host logic (CPU "port").

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference lines it follows; ``tests/test_oracle_vs_reference.py`` (container only)
checks each against the reference's own unmodified function through
``oracle/ref_bridge.py`` and ``tests/golden/`` holds vectors generated that way.

Restated here (reference file:line):
  split_tta_latents              delta_experiment/scripts/common.py:1365-1401
  fm_loss_conditioned            delta_experiment/scripts/common.py:414-489
  fm_loss_conditioned_fixed      delta_experiment/scripts/common.py:492-559
  fm_loss                        delta_experiment/scripts/common.py:274-343
  LoRALinear / inject_lora       lora_experiment/scripts/run_lora_tta.py:224-382
  lora_parameters / reset_lora   lora_experiment/scripts/run_lora_tta.py:385-409
  lora_tta_loop                  lora_experiment/scripts/run_lora_tta.py:425-547
  DeltaA / delta_a_loop          delta_experiment/scripts/run_delta_a.py:88-305
  clip_grad_norm / adamw_step    torch.nn.utils.clip_grad_norm_ / torch.optim.AdamW
                                 as called at run_lora_tta.py:462-468,513-514
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------

def split_tta_latents(latents, num_context_latents: int, holdout_fraction: float = 0.25):
    """common.py:1365-1401 -- (cond, train, val|None) along the latent-time axis."""
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


# ----------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------

def _patch_t(dit) -> int:
    cfg = dit.config if hasattr(dit, "config") else dit.dit.config  # common.py:262-271
    return cfg.patch_size[0]


def build_step_inputs(cond, target, sigma, noise, dtype, num_train_timesteps=1000, patch_t=1):
    """common.py:458-470: noising, concat and per-frame timestep for given (sigma, eps)."""
    B = cond.shape[0]
    s = sigma.view(B, 1, 1, 1, 1)
    noisy = (1.0 - s) * target + s * noise  # fp32 (sigma is fp32)
    hidden = torch.cat([cond, noisy], dim=2).to(dtype)
    n_cond, n_tgt = cond.shape[2] // patch_t, target.shape[2] // patch_t
    timestep = torch.zeros(B, n_cond + n_tgt, device=cond.device, dtype=dtype)
    timestep[:, n_cond:] = (sigma * num_train_timesteps).unsqueeze(1).expand(B, n_tgt).to(dtype)
    return hidden, timestep, n_cond


def fm_loss_given(dit, cond, target, prompt_embeds, prompt_mask, sigma, noise, dtype,
                  forward_fn=None, num_train_timesteps=1000, return_pred=False):
    """The deterministic part of common.py:414-489 for explicit (sigma, eps)."""
    hidden, timestep, n_cond = build_step_inputs(cond, target, sigma, noise, dtype,
                                                 num_train_timesteps, _patch_t(dit))
    if forward_fn is not None:
        pred = forward_fn(hidden, timestep, n_cond)
    else:
        pred = dit(hidden_states=hidden, timestep=timestep, encoder_hidden_states=prompt_embeds,
                   encoder_attention_mask=prompt_mask, num_cond_latents=n_cond)
    t_cond = cond.shape[2]
    loss = F.mse_loss(pred[:, :, t_cond:].to(torch.float32), (noise - target).to(torch.float32))
    return (loss, pred) if return_pred else loss


def draw_sigma_noise(target, device, sigma_min=0.001, sigma_max=1.0, generator=None):
    """common.py:458,462: sigma ~ U[sigma_min, sigma_max] fp32 [B]; eps ~ N(0,1) like target."""
    B = target.shape[0]
    sigma = torch.rand(B, device=device, dtype=torch.float32, generator=generator) * (sigma_max - sigma_min) + sigma_min
    if generator is None:
        noise = torch.randn_like(target)
    else:
        noise = torch.randn(target.shape, device=target.device, dtype=target.dtype, generator=generator)
    return sigma, noise


def fm_loss_conditioned(dit, cond, target, prompt_embeds, prompt_mask, device="cpu",
                        dtype=torch.float32, forward_fn=None, **kw):
    """common.py:414-489."""
    sigma, noise = draw_sigma_noise(target, device)
    return fm_loss_given(dit, cond, target, prompt_embeds, prompt_mask, sigma, noise, dtype, forward_fn, **kw)


def fm_loss_conditioned_fixed(dit, cond, target, prompt_embeds, prompt_mask, fixed_sigmas: Sequence[float],
                              fixed_noises: Sequence[torch.Tensor], device="cpu", dtype=torch.float32,
                              forward_fn=None) -> float:
    """common.py:492-559: mean over sigmas x noises of the no-grad anchor loss."""
    total, count = 0.0, 0
    for sv in fixed_sigmas:
        sigma = torch.tensor([sv], device=device, dtype=torch.float32)
        for noise in fixed_noises:
            with torch.no_grad():
                # reference broadcasts a [1] sigma over the batch (view(1,1,1,1,1))
                hidden, timestep, n_cond = build_step_inputs(
                    cond, target, sigma.expand(cond.shape[0]), noise, dtype, 1000, _patch_t(dit))
                pred = forward_fn(hidden, timestep, n_cond) if forward_fn is not None else dit(
                    hidden_states=hidden, timestep=timestep, encoder_hidden_states=prompt_embeds,
                    encoder_attention_mask=prompt_mask, num_cond_latents=n_cond)
            total += F.mse_loss(pred[:, :, cond.shape[2]:].float(), (noise - target).float()).item()
            count += 1
    return total / max(count, 1)


def fm_loss(dit, latents, prompt_embeds, prompt_mask, device="cpu", dtype=torch.float32, forward_fn=None):
    """common.py:274-343: unconditioned variant -- every frame noised with one sigma."""
    B, _, T = latents.shape[:3]
    sigma, noise = draw_sigma_noise(latents, device)
    s = sigma.view(B, 1, 1, 1, 1)
    noisy = ((1.0 - s) * latents + s * noise).to(dtype)
    timestep = (sigma * 1000).unsqueeze(1).expand(B, T // _patch_t(dit)).to(dtype)
    pred = forward_fn(noisy, timestep) if forward_fn is not None else dit(
        hidden_states=noisy, timestep=timestep, encoder_hidden_states=prompt_embeds,
        encoder_attention_mask=prompt_mask)
    return F.mse_loss(pred.float(), (noise - latents).float())


# ----------------------------------------------------------------------------
# LoRA (run_lora_tta.py:224-418)
# ----------------------------------------------------------------------------

class LoRALinear(nn.Module):
    """orig(x) + up(down(x)) * alpha/rank; down kaiming-uniform(a=sqrt 5), up zeros."""

    def __init__(self, original: nn.Linear, rank=8, alpha=16.0):
        super().__init__()
        self.original, self.rank, self.alpha = original, rank, alpha
        self.scaling = alpha / rank
        self.lora_down = nn.Linear(original.in_features, rank, bias=False)
        self.lora_up = nn.Linear(rank, original.out_features, bias=False)
        reset_lora([self])

    def forward(self, x):
        y = self.original(x)
        return y + self.lora_up(self.lora_down(x.to(self.lora_down.weight.dtype))) * self.scaling


def parse_target_blocks(spec: str, num_blocks: int):
    """run_lora_tta.py:263-283: 'all' | 'last_N' | 'i,j,k' -> None | set."""
    spec = spec.strip().lower()
    if spec == "all":
        return None
    if spec.startswith("last_"):
        n = int(spec.split("_", 1)[1])
        if n <= 0 or n > num_blocks:
            raise ValueError(f"last_{n} invalid for {num_blocks} blocks")
        return set(range(num_blocks - n, num_blocks))
    idx = {int(s.strip()) for s in spec.split(",")}
    for i in idx:
        if not 0 <= i < num_blocks:
            raise ValueError(f"Block index {i} out of range [0, {num_blocks})")
    return idx


LORA_SITES = (  # injection order within a block (run_lora_tta.py:327-380)
    ("attn", "qkv", "qkv"), ("attn", "proj", "proj"),
    ("cross_attn", "q_linear", "qkv"), ("cross_attn", "kv_linear", "qkv"), ("cross_attn", "proj", "proj"),
)


def inject_lora(dit, rank=8, alpha=16.0, target_modules=("qkv", "proj"), target_ffn=False,
                target_blocks="all") -> List[LoRALinear]:
    ref = next(dit.parameters())
    chosen = parse_target_blocks(target_blocks, len(dit.blocks))
    mods = []

    def wrap(parent, name):
        lin = getattr(parent, name, None)
        if isinstance(lin, nn.Linear):
            m = LoRALinear(lin, rank, alpha).to(device=ref.device, dtype=ref.dtype)
            setattr(parent, name, m)
            mods.append(m)

    for i, blk in enumerate(dit.blocks):
        if chosen is not None and i not in chosen:
            continue
        for parent, name, group in LORA_SITES:
            if group in target_modules and hasattr(blk, parent):
                wrap(getattr(blk, parent), name)
        if target_ffn and hasattr(blk, "ffn"):
            for name in ("w1", "w2", "w3"):
                wrap(blk.ffn, name)
    return mods


def lora_parameters(mods) -> List[nn.Parameter]:
    out = []
    for m in mods:
        out += [m.lora_down.weight, m.lora_up.weight]
    return out


def reset_lora(mods):
    for m in mods:
        nn.init.kaiming_uniform_(m.lora_down.weight, a=math.sqrt(5))
        nn.init.zeros_(m.lora_up.weight)


# ----------------------------------------------------------------------------
# clip + AdamW restated explicitly (what torch computes at run_lora_tta.py:513-514)
# ----------------------------------------------------------------------------

def clip_grad_norm(grads: Sequence[torch.Tensor], max_norm: float) -> torch.Tensor:
    """torch.nn.utils.clip_grad_norm_: total = ||(||g_i||)_i||_2 ; g *= min(1, max/(total+1e-6))."""
    norms = torch.stack([g.norm(2) for g in grads])
    total = norms.norm(2)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef.to(g.dtype))
    return total


def adamw_step(params, grads, exp_avg, exp_avg_sq, step: int, lr, betas=(0.9, 0.999), eps=1e-8, wd=0.01):
    """torch.optim.AdamW single-tensor formulation (decoupled decay, bias correction)."""
    b1, b2 = betas
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        p.mul_(1.0 - lr * wd)
        m.lerp_(g, 1.0 - b1)
        v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)


def warmup_lr(lr: float, step: int, warmup_steps: int) -> float:
    """run_lora_tta.py:494-497 (the last warm-up value, == lr, persists afterwards)."""
    if warmup_steps > 0 and step < warmup_steps:
        return lr * (step + 1) / warmup_steps
    return lr


# ----------------------------------------------------------------------------
# loops
# ----------------------------------------------------------------------------

def lora_tta_loop(dit, mods, cond, train, prompt_embeds, prompt_mask, num_steps=20, lr=2e-4,
                  warmup_steps=3, weight_decay=0.01, max_grad_norm=1.0, device="cpu",
                  dtype=torch.float32, on_step: Optional[Callable] = None) -> Dict:
    """run_lora_tta.py:425-547 without early stopping / variants (single variant)."""
    params = lora_parameters(mods)
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    dit.train()
    losses = []
    for step in range(num_steps):
        for p in params:
            p.grad = None
        cur_lr = warmup_lr(lr, step, warmup_steps)
        torch.randint(0, 1, (1,))  # run_lora_tta.py:499 consumes one CPU draw per step
        sigma, noise = draw_sigma_noise(train, device)
        with torch.enable_grad():
            for p in params:
                p.requires_grad_(True)
            loss, pred = fm_loss_given(dit, cond, train, prompt_embeds, prompt_mask, sigma, noise, dtype,
                                       return_pred=True)
            grads = list(torch.autograd.grad(loss, params))
        raw = [g.clone() for g in grads] if on_step is not None else None
        total = clip_grad_norm(grads, max_grad_norm)
        with torch.no_grad():
            adamw_step(params, grads, m, v, step + 1, cur_lr, eps=1e-8, wd=weight_decay)
        losses.append(loss.item())
        if on_step is not None:
            on_step(step=step, sigma=sigma, noise=noise, loss=loss.detach(), pred=pred.detach(),
                    grads=raw, total_norm=total, lr=cur_lr)
    dit.eval()
    return {"losses": losses}


class DeltaA(nn.Module):
    """run_delta_a.py:88-217: one fp32 vector added to the timestep embedding."""

    def __init__(self, dit, dim=512):
        super().__init__()
        self.dit = dit
        self.delta = nn.Parameter(torch.zeros(dim))

    @property
    def config(self):
        return self.dit.config

    def forward(self, hidden_states, timestep, encoder_hidden_states, encoder_attention_mask=None,
                num_cond_latents=0, **kw):
        d = self.dit
        B, _, T, H, W = hidden_states.shape
        grid = (T // d.patch_size[0], H // d.patch_size[1], W // d.patch_size[2])
        dtype = d.x_embedder.proj.weight.dtype
        x = d.x_embedder(hidden_states.to(dtype))
        t = d.t_embedder(timestep.to(dtype).float().flatten(), dtype=torch.float32).reshape(B, grid[0], -1)
        t = t + self.delta[None, None]
        y, y_seqlens = d.embed_text(encoder_hidden_states.to(dtype), encoder_attention_mask)
        for blk in d.blocks:
            x = blk(x, y, t, y_seqlens, grid, num_cond_latents=num_cond_latents)
        x = d.final_layer(x, t, grid)
        return d.unpatchify(x, *grid).to(torch.float32)


def delta_loop(wrapper, params, cond, train, prompt_embeds, prompt_mask, num_steps=20, lr=1e-3,
               device="cpu", dtype=torch.float32, per_tensor_clip=False, on_step=None) -> Dict:
    """optimize_delta_a (run_delta_a.py:224-305) / optimize_delta_b (run_delta_b.py:337-421):
    AdamW(lr, betas=(0.9,0.999), eps=1e-15) (default wd 0.01), no warm-up, clip 1.0
    (per tensor for delta-B, run_delta_b.py:386-388)."""
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    losses = []
    for step in range(num_steps):
        torch.randint(0, 1, (1,))  # run_delta_a.py:262
        sigma, noise = draw_sigma_noise(train, device)
        with torch.enable_grad():
            for p in params:
                p.requires_grad_(True)
            loss = fm_loss_given(wrapper, cond, train, prompt_embeds, prompt_mask, sigma, noise, dtype)
            grads = list(torch.autograd.grad(loss, params))
        raw = [g.clone() for g in grads] if on_step is not None else None
        if per_tensor_clip:
            for g in grads:
                clip_grad_norm([g], 1.0)
        else:
            clip_grad_norm(grads, 1.0)
        with torch.no_grad():
            adamw_step(params, grads, m, v, step + 1, lr, eps=1e-15, wd=0.01)
        losses.append(loss.item())
        if on_step is not None:
            on_step(step=step, sigma=sigma, noise=noise, loss=loss.detach(), grads=raw)
    return {"losses": losses}


# ---------------------------------------------------------------------------------------------------- denoise loop
def flow_match_sigmas(num_inference_steps, shift=1.0):
    s = torch.linspace(1.0, 0.0, num_inference_steps + 1, dtype=torch.float64)
    if shift != 1.0:
        s = shift * s / (1.0 + (shift - 1.0) * s)
    return s.to(torch.float32)


@torch.no_grad()
def denoise_latents(dit, cond, prompt_embeds, prompt_mask, init_noise, num_inference_steps, dtype,
                    negative_prompt_embeds=None, negative_prompt_mask=None, guidance_scale=4.0, shift=1.0):
    """Flow-matching Euler sampler with classifier-free guidance over FULL forwards of the oracle DiT (no context
    cache): the checker for longcat_video_tta_b200.denoise.denoise_latents.  The upstream sampler
    (generate_vc, reached from common.py:566-611) is absent from the reference; this is our statement of it, consistent
    with the training parametrisation of common.py:458-488 (v = eps - x0 = d x_sigma / d sigma)."""
    x = init_noise.to(torch.float32).clone()
    sig = flow_match_sigmas(num_inference_steps, shift)
    B = cond.shape[0]
    for k in range(num_inference_steps):
        s, s_next = float(sig[k]), float(sig[k + 1])
        sigma = torch.full((B,), s, dtype=torch.float32, device=cond.device)
        xt = x.to(dtype)
        hidden, timestep, n_cond = build_step_inputs(cond.to(dtype), xt, sigma, xt, dtype, 1000, _patch_t(dit))

        def vel(pe, pm):
            pred = dit(hidden_states=hidden, timestep=timestep, encoder_hidden_states=pe.to(dtype),
                       encoder_attention_mask=pm, num_cond_latents=n_cond)
            return pred[:, :, cond.shape[2]:].to(torch.float32)

        v = vel(prompt_embeds, prompt_mask)
        if negative_prompt_embeds is not None and guidance_scale != 1.0:
            vn = vel(negative_prompt_embeds, negative_prompt_mask)
            v = vn + guidance_scale * (v - vn)
        x = x + (s_next - s) * v
    return x
