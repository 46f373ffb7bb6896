"""Oracle restatement of the LongCat-Video DiT forward (plain PyTorch, any device).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity unpinned by the
reference: the upstream module (``longcat_video/modules/longcat_video_dit.py`` of
meituan-longcat/LongCat-Video, un-pinned clone) is absent from the reference
tree.  This file restates it from

* the reference's own mirror of the forward protocol
  (``delta_experiment/scripts/run_delta_a.py:134-217``),
* the module / attribute names and shapes the reference touches
  (``lora_experiment/scripts/run_lora_tta.py:146-168``,
  ``delta_experiment/scripts/run_norm_tune_tta.py:78-96``,
  ``delta_experiment/scripts/run_film_tta.py:80-83``),
* SURVEY.md Appendix A for the internals.

The module keeps the upstream attribute layout (``x_embedder.proj``,
``t_embedder(t, dtype)``, ``y_embedder``, ``blocks[i].attn.qkv`` ...) so that the
reference's unmodified adapter code (LoRA injection, delta / norm / FiLM wrappers)
can operate on it through ``oracle/ref_bridge.py``.

Mixed precision: parameters may be fp32 (the oracle proper) or bf16 (restating what
the reference runs on a GPU).  The "fp32 islands" upstream creates with
``amp.autocast(dtype=torch.float32)`` are written as explicit ``.float()`` casts.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# configs (BASELINE.json configs[0] and configs[1])
# ----------------------------------------------------------------------------

def ffn_hidden_dim(hidden_size: int, mlp_ratio: int = 4, multiple_of: int = 256) -> int:
    """Appendix A.4: F = 256 * ceil(int(2 * 4C / 3) / 256)  (C=4096 -> 11008)."""
    h = int(2 * hidden_size * mlp_ratio / 3)
    return multiple_of * ((h + multiple_of - 1) // multiple_of)


def make_config(name: str = "tiny", **overrides) -> SimpleNamespace:
    base = dict(
        in_channels=16, out_channels=16, patch_size=(1, 2, 2),
        adaln_tembed_dim=512, frequency_embedding_size=256, mlp_ratio=4,
        text_tokens_zero_pad=True, rope_base=10000.0, norm_eps=1e-6,
    )
    if name == "tiny":
        base.update(hidden_size=512, depth=2, num_heads=4, caption_channels=512)
    elif name == "13.6b":
        base.update(hidden_size=4096, depth=48, num_heads=32, caption_channels=4096)
    else:
        raise ValueError(name)
    base.update(overrides)
    cfg = SimpleNamespace(**base)
    cfg.head_dim = cfg.hidden_size // cfg.num_heads
    cfg.ffn_dim = ffn_hidden_dim(cfg.hidden_size, cfg.mlp_ratio)
    return cfg


# ----------------------------------------------------------------------------
# leaf modules
# ----------------------------------------------------------------------------

class LayerNormFP32(nn.LayerNorm):
    """LayerNorm evaluated in fp32, result cast back (Appendix A.4)."""

    def forward(self, x):
        w = self.weight.float() if self.weight is not None else None
        b = self.bias.float() if self.bias is not None else None
        return F.layer_norm(x.float(), self.normalized_shape, w, b, self.eps).to(x.dtype)


class RMSNormFP32(nn.Module):
    """x * rsqrt(mean(x^2) + eps) in fp32, cast back, times weight (Appendix A.5)."""

    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        xf = x.float()
        out = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).type_as(x)
        return out * self.weight


def modulate_fp32(norm, x, shift, scale):
    """LN(x.float()) * (1 + scale) + shift -> x.dtype; shift/scale fp32 [B,T,1,C]."""
    dtype = x.dtype
    y = norm(x.float())
    y = y * (scale + 1.0) + shift
    return y.to(dtype)


class PatchEmbed3D(nn.Module):
    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.patch_size = tuple(patch_size)
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)

    def forward(self, x):
        x = self.proj(x)  # [B, C, T, H', W']
        return x.flatten(2).transpose(1, 2)  # token order (t, h, w)


class TimestepEmbedder(nn.Module):
    """cat(cos, sin) sinusoid(256) -> Linear -> SiLU -> Linear  (Appendix A.2)."""

    def __init__(self, t_embed_dim: int, frequency_embedding_size: int = 256):
        super().__init__()
        self.mlp = nn.Sequential(
            nn.Linear(frequency_embedding_size, t_embed_dim, bias=True),
            nn.SiLU(),
            nn.Linear(t_embed_dim, t_embed_dim, bias=True),
        )
        self.frequency_embedding_size = frequency_embedding_size

    @staticmethod
    def timestep_embedding(t, dim, max_period=10000):
        half = dim // 2
        freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
        args = t[:, None].float() * freqs[None]
        return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)

    def forward(self, t, dtype=torch.float32):
        # reference call: t_embedder(timestep.float().flatten(), dtype=torch.float32)
        # under autocast(float32) (run_delta_a.py:162-165): linears evaluate in fp32.
        f = self.timestep_embedding(t, self.frequency_embedding_size).to(dtype)
        l0, l2 = self.mlp[0], self.mlp[2]
        h = F.linear(f.float(), l0.weight.float(), l0.bias.float())
        h = F.silu(h)
        h = F.linear(h, l2.weight.float(), l2.bias.float())
        return h.to(dtype)


class CaptionEmbedder(nn.Module):
    def __init__(self, in_channels: int, hidden_size: int):
        super().__init__()
        self.y_proj = nn.Sequential(
            nn.Linear(in_channels, hidden_size, bias=True),
            nn.GELU(approximate="tanh"),
            nn.Linear(hidden_size, hidden_size, bias=True),
        )

    def forward(self, caption):
        return self.y_proj(caption)


class RotaryPositionalEmbedding3D(nn.Module):
    """Appendix A.6: dim_h = dim_w = 2*(D//6), dim_t = D - dim_h - dim_w; interleaved pairs."""

    def __init__(self, head_dim: int, base: float = 10000.0):
        super().__init__()
        self.head_dim = head_dim
        self.base = base
        self._cache = {}

    def axis_dims(self) -> Tuple[int, int, int]:
        d_hw = 2 * (self.head_dim // 6)
        return self.head_dim - 2 * d_hw, d_hw, d_hw

    def freqs(self, grid: Sequence[int], device) -> torch.Tensor:
        key = (tuple(grid), str(device))
        if key not in self._cache:
            T, H, W = grid
            parts = []
            for n, d in zip((T, H, W), self.axis_dims()):
                inv = 1.0 / (self.base ** (torch.arange(0, d, 2, dtype=torch.float32)[: d // 2] / d))
                ang = torch.arange(n, dtype=torch.float32)[:, None] * inv[None]  # [n, d/2]
                parts.append(ang.repeat_interleave(2, dim=-1))  # [n, d]
            ft, fh, fw = parts
            f = torch.cat([
                ft[:, None, None, :].expand(T, H, W, -1),
                fh[None, :, None, :].expand(T, H, W, -1),
                fw[None, None, :, :].expand(T, H, W, -1),
            ], dim=-1).reshape(T * H * W, self.head_dim)
            self._cache[key] = f.to(device)
        return self._cache[key]

    @staticmethod
    def rotate_half(x):
        x = x.unflatten(-1, (-1, 2))
        x1, x2 = x.unbind(-1)
        return torch.stack((-x2, x1), dim=-1).flatten(-2)

    def forward(self, q, k, grid):
        f = self.freqs(grid, q.device)
        cos, sin = f.cos(), f.sin()
        qf, kf = q.float(), k.float()
        qf = qf * cos + self.rotate_half(qf) * sin
        kf = kf * cos + self.rotate_half(kf) * sin
        return qf.type_as(q), kf.type_as(k)


def _sdpa(q, k, v):
    """softmax(q k^T / sqrt(D)) v with fp32 softmax; q,k,v [B,H,N,D]."""
    scale = q.shape[-1] ** -0.5
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, v.float()).to(q.dtype)


class Attention(nn.Module):
    """Appendix A.5.  ``self.bsa = dict(chunk=(4, 4, 8), sparsity=0.9375)`` switches to the block-sparse variant (A.9,
    our definition: oracle/bsa_oracle.py)."""

    def __init__(self, dim: int, num_heads: int, eps: float, rope_base: float):
        super().__init__()
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.q_norm = RMSNormFP32(self.head_dim, eps)
        self.k_norm = RMSNormFP32(self.head_dim, eps)
        self.proj = nn.Linear(dim, dim)
        self.rope_3d = RotaryPositionalEmbedding3D(self.head_dim, rope_base)
        self.bsa = None

    def forward(self, x, shape=None, num_cond_latents=None):
        B, N, C = x.shape
        qkv = self.qkv(x).view(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)  # [B,H,N,D]
        q, k = self.q_norm(q), self.k_norm(k)
        q, k = self.rope_3d(q, k, shape)
        if self.bsa is not None:
            from .bsa_oracle import block_sparse_attention
            o = block_sparse_attention(q, k, v, shape, num_cond_latents or 0, **self.bsa)
        elif num_cond_latents is not None and num_cond_latents > 0:
            nc = num_cond_latents * (N // shape[0])
            o = torch.cat([_sdpa(q[:, :, :nc], k[:, :, :nc], v[:, :, :nc]), _sdpa(q[:, :, nc:], k, v)], dim=2)
        else:
            o = _sdpa(q, k, v)
        o = o.transpose(1, 2).reshape(B, N, C)
        return self.proj(o)


class MultiHeadCrossAttention(nn.Module):
    """Appendix A.7: noise tokens only are queries; cond rows of the output are zero."""

    def __init__(self, dim: int, num_heads: int, eps: float):
        super().__init__()
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.q_linear = nn.Linear(dim, dim)
        self.kv_linear = nn.Linear(dim, dim * 2)
        self.proj = nn.Linear(dim, dim)
        self.q_norm = RMSNormFP32(self.head_dim, eps)
        self.k_norm = RMSNormFP32(self.head_dim, eps)

    def _process(self, x, cond, kv_seqlen):
        B, N, C = x.shape
        q = self.q_linear(x).view(B, N, self.num_heads, self.head_dim)
        kv = self.kv_linear(cond).view(1, -1, 2, self.num_heads, self.head_dim)
        k, v = kv.unbind(2)  # [1, sum(M), H, D]
        q, k = self.q_norm(q), self.k_norm(k)
        outs, start = [], 0
        for b in range(B):  # varlen: sample b attends to its own text tokens
            m = int(kv_seqlen[b])
            kb, vb = k[:, start:start + m], v[:, start:start + m]
            start += m
            outs.append(_sdpa(q[b:b + 1].transpose(1, 2), kb.transpose(1, 2), vb.transpose(1, 2)).transpose(1, 2))
        o = torch.cat(outs, dim=0).reshape(B, N, C)
        return self.proj(o)

    def forward(self, x, cond, kv_seqlen, num_cond_latents=None, shape=None):
        if num_cond_latents is None or num_cond_latents == 0:
            return self._process(x, cond, kv_seqlen)
        B, N, C = x.shape
        nc = num_cond_latents * (N // shape[0])
        out_noise = self._process(x[:, nc:], cond, kv_seqlen)
        zeros = torch.zeros((B, nc, C), dtype=out_noise.dtype, device=out_noise.device)
        return torch.cat([zeros, out_noise], dim=1)


class FeedForwardSwiGLU(nn.Module):
    def __init__(self, dim: int, hidden_dim: int):
        super().__init__()
        self.w1 = nn.Linear(dim, hidden_dim, bias=False)
        self.w2 = nn.Linear(hidden_dim, dim, bias=False)
        self.w3 = nn.Linear(dim, hidden_dim, bias=False)

    def forward(self, x):
        return self.w2(F.silu(self.w1(x)) * self.w3(x))


def _fp32_linear(seq: nn.Sequential, t):
    """Sequential(SiLU, Linear) evaluated in fp32 (upstream: under autocast(float32)).

    Goes through ``seq(...)`` when parameters are already fp32 so that forward hooks
    registered on ``adaLN_modulation`` (FiLM, run_film_tta.py:146-160) still fire.
    """
    lin = seq[1]
    if lin.weight.dtype == torch.float32:
        return seq(t.float())
    out = F.linear(F.silu(t.float()), lin.weight.float(), lin.bias.float())
    for hook in seq._forward_hooks.values():  # keep hook semantics for bf16 params
        r = hook(seq, (t,), out)
        if r is not None:
            out = r
    return out


class LongCatSingleStreamBlock(nn.Module):
    """Appendix A.4 data flow."""

    def __init__(self, cfg):
        super().__init__()
        C = cfg.hidden_size
        self.hidden_size = C
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(cfg.adaln_tembed_dim, 6 * C, bias=True))
        self.mod_norm_attn = LayerNormFP32(C, eps=cfg.norm_eps, elementwise_affine=False)
        self.mod_norm_ffn = LayerNormFP32(C, eps=cfg.norm_eps, elementwise_affine=False)
        self.pre_crs_attn_norm = LayerNormFP32(C, eps=cfg.norm_eps, elementwise_affine=True)
        self.attn = Attention(C, cfg.num_heads, cfg.norm_eps, cfg.rope_base)
        self.cross_attn = MultiHeadCrossAttention(C, cfg.num_heads, cfg.norm_eps)
        self.ffn = FeedForwardSwiGLU(C, cfg.ffn_dim)

    def forward(self, x, y, t, y_seqlen, latent_shape, num_cond_latents=None):
        x_dtype = x.dtype
        B, N, C = x.shape
        T = latent_shape[0]
        mod = _fp32_linear(self.adaLN_modulation, t).unsqueeze(2)  # [B,T,1,6C] fp32
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, dim=-1)

        x_m = modulate_fp32(self.mod_norm_attn, x.view(B, T, -1, C), shift_msa, scale_msa).view(B, N, C)
        x_s = self.attn(x_m, shape=latent_shape, num_cond_latents=num_cond_latents)
        x = (x.float() + (gate_msa * x_s.view(B, T, -1, C).float()).view(B, N, C)).to(x_dtype)

        x = x + self.cross_attn(self.pre_crs_attn_norm(x), y, y_seqlen,
                                num_cond_latents=num_cond_latents, shape=latent_shape)

        x_m = modulate_fp32(self.mod_norm_ffn, x.view(B, T, -1, C), shift_mlp, scale_mlp).view(B, N, C)
        x_s = self.ffn(x_m)
        x = (x.float() + (gate_mlp * x_s.view(B, T, -1, C).float()).view(B, N, C)).to(x_dtype)
        return x


class FinalLayerFP32(nn.Module):
    """Appendix A.8: LN -> modulate -> Linear(C, prod(patch)*out), all fp32."""

    def __init__(self, cfg):
        super().__init__()
        C = cfg.hidden_size
        n_patch = cfg.patch_size[0] * cfg.patch_size[1] * cfg.patch_size[2]
        self.norm_final = LayerNormFP32(C, eps=cfg.norm_eps, elementwise_affine=False)
        self.linear = nn.Linear(C, n_patch * cfg.out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(cfg.adaln_tembed_dim, 2 * C, bias=True))

    def forward(self, x, t, latent_shape):
        B, N, C = x.shape
        T = latent_shape[0]
        mod = _fp32_linear(self.adaLN_modulation, t).unsqueeze(2)
        shift, scale = mod.chunk(2, dim=-1)
        xf = self.norm_final(x.float().view(B, T, -1, C)) * (scale + 1.0) + shift
        return F.linear(xf.view(B, N, C), self.linear.weight.float(), self.linear.bias.float())


class OracleDiT(nn.Module):
    """Stands in for ``LongCatVideoTransformer3DModel`` (call site common.py:476-482)."""

    def __init__(self, cfg):
        super().__init__()
        self.config = cfg
        self.patch_size = tuple(cfg.patch_size)
        self.text_tokens_zero_pad = cfg.text_tokens_zero_pad
        self.x_embedder = PatchEmbed3D(cfg.patch_size, cfg.in_channels, cfg.hidden_size)
        self.t_embedder = TimestepEmbedder(cfg.adaln_tembed_dim, cfg.frequency_embedding_size)
        self.y_embedder = CaptionEmbedder(cfg.caption_channels, cfg.hidden_size)
        self.blocks = nn.ModuleList([LongCatSingleStreamBlock(cfg) for _ in range(cfg.depth)])
        self.final_layer = FinalLayerFP32(cfg)
        self.gradient_checkpointing = False

    def unpatchify(self, x, N_t, N_h, N_w):
        """[B, N, pt*ph*pw*C_out] -> [B, C_out, T, H, W]; inner layout (pt ph pw c) (A.1)."""
        pt, ph, pw = self.patch_size
        B = x.shape[0]
        x = x.view(B, N_t, N_h, N_w, pt, ph, pw, -1)
        x = x.permute(0, 7, 1, 4, 2, 5, 3, 6)
        return x.reshape(B, -1, N_t * pt, N_h * ph, N_w * pw)

    def embed_text(self, encoder_hidden_states, encoder_attention_mask):
        """run_delta_a.py:170-192: y_embedder -> zero-pad by mask -> pack valid tokens."""
        y = self.y_embedder(encoder_hidden_states)
        if self.text_tokens_zero_pad and encoder_attention_mask is not None:
            y = y * encoder_attention_mask[:, None, :, None]
            encoder_attention_mask = (encoder_attention_mask * 0 + 1).to(encoder_attention_mask.dtype)
        if encoder_attention_mask is not None:
            m = encoder_attention_mask.squeeze(1).squeeze(1)
            y = y.squeeze(1).masked_select(m.unsqueeze(-1) != 0).view(1, -1, y.shape[-1])
            y_seqlens = m.sum(dim=1).tolist()
        else:
            y_seqlens = [y.shape[2]] * y.shape[0]
            y = y.squeeze(1).reshape(1, -1, y.shape[-1])
        return y, y_seqlens

    def forward(self, hidden_states, timestep, encoder_hidden_states,
                encoder_attention_mask=None, num_cond_latents=0, **kwargs):
        B, _, T, H, W = hidden_states.shape
        N_t, N_h, N_w = T // self.patch_size[0], H // self.patch_size[1], W // self.patch_size[2]
        if timestep.dim() == 1:
            timestep = timestep.unsqueeze(1).expand(-1, N_t)
        dtype = self.x_embedder.proj.weight.dtype
        x = self.x_embedder(hidden_states.to(dtype))
        t = self.t_embedder(timestep.to(dtype).float().flatten(), dtype=torch.float32).reshape(B, N_t, -1)
        y, y_seqlens = self.embed_text(encoder_hidden_states.to(dtype), encoder_attention_mask)
        for block in self.blocks:
            if self.gradient_checkpointing and torch.is_grad_enabled():
                # the reference turns this on for training (run_lora_tta.py:806-811): per-block recompute
                from torch.utils.checkpoint import checkpoint
                x = checkpoint(block, x, y, t, y_seqlens, (N_t, N_h, N_w), num_cond_latents=num_cond_latents,
                               use_reentrant=False)
            else:
                x = block(x, y, t, y_seqlens, (N_t, N_h, N_w), num_cond_latents=num_cond_latents)
        x = self.final_layer(x, t, (N_t, N_h, N_w))
        return self.unpatchify(x, N_t, N_h, N_w).to(torch.float32)


def build_oracle_dit(name: str = "tiny", seed: int = 0, dtype=torch.float32, init_std: Optional[float] = None,
                     **overrides) -> OracleDiT:
    """Seeded random-init DiT.  ``init_std`` None -> torch default init (tiny config,
    SURVEY 8d config 1); a float -> N(0, std) weights, zero biases except adaLN
    (13.6 B config: std 0.02 keeps activations finite through 48 blocks)."""
    cfg = make_config(name, **overrides)
    g = torch.Generator().manual_seed(seed)
    with torch.random.fork_rng():
        torch.manual_seed(seed)
        dit = OracleDiT(cfg)
    if init_std is not None:
        with torch.no_grad():
            for n, p in dit.named_parameters():
                if p.dim() >= 2:
                    p.copy_(torch.randn(p.shape, generator=g) * init_std)
                elif n.endswith("bias"):
                    p.copy_(torch.randn(p.shape, generator=g) * init_std)
                else:  # 1-d norm weights: perturb around 1 so their gradients are exercised
                    p.copy_(1.0 + torch.randn(p.shape, generator=g) * init_std)
    dit.requires_grad_(False)
    return dit.to(dtype)
