"""B200DiT: the LongCat-Video DiT as the TTA path sees it.

The module mirrors the attribute set the reference touches on ``LongCatVideoTransformer3DModel`` --
``config.patch_size / adaln_tembed_dim / hidden_size``, ``patch_size``, ``text_tokens_zero_pad``, ``x_embedder.proj``,
``t_embedder``, ``y_embedder``, ``blocks[i].{adaLN_modulation, mod_norm_attn, mod_norm_ffn, pre_crs_attn_norm,
attn.{qkv,proj,q_norm,k_norm}, cross_attn.{q_linear,kv_linear,proj,q_norm,k_norm}, ffn.{w1,w2,w3}}``,
``final_layer`` and ``unpatchify`` (delta_experiment/scripts/common.py:262-271,452; run_delta_a.py:147-214;
lora_experiment/scripts/run_lora_tta.py:146-168; run_norm_tune_tta.py:78-96; run_film_tta.py:80-83) -- so the
reference's own adapter surgery (``inject_lora_into_dit`` replaces ``attn.qkv`` with a wrapper, hooks are put on
``t_embedder`` / ``adaLN_modulation``) keeps working.  The leaves are real ``nn.Linear`` / ``nn.LayerNorm`` modules:
their weights are the storage the sm_100a kernels read.  The leaves' own ``forward`` is never used on the hot path:
``B200DiT.forward`` runs the whole network through ``engine.TTAEngine`` (hand-written kernels behind the C ABI).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn

BF16 = torch.bfloat16


def ffn_hidden_dim(hidden_size: int, mlp_ratio: int = 4, multiple_of: int = 256) -> int:
    h = int(2 * hidden_size * mlp_ratio / 3)
    return multiple_of * ((h + multiple_of - 1) // multiple_of)


def make_config(name: str = "13.6b", **overrides) -> SimpleNamespace:
    base = dict(in_channels=16, out_channels=16, patch_size=(1, 2, 2), adaln_tembed_dim=512,
                frequency_embedding_size=256, mlp_ratio=4, text_tokens_zero_pad=True, rope_base=10000.0, norm_eps=1e-6,
                enable_bsa=False, bsa_params=None)   # bsa_params: dict(sparsity=0.9375, chunk=(4, 4, 8)) -- bsa.py
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
    if cfg.head_dim != 128:
        raise ValueError("the sm_100a attention kernels are built for head_dim 128")
    return cfg


class _RMSNormParams(nn.Module):
    def __init__(self, dim: int, eps: float):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))


class _PatchEmbed3D(nn.Module):
    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.patch_size = tuple(patch_size)
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)


class _TimestepEmbedder(nn.Module):
    def __init__(self, dim: int, freq: int):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(freq, dim, bias=True), nn.SiLU(), nn.Linear(dim, dim, bias=True))
        self.frequency_embedding_size = freq


class _CaptionEmbedder(nn.Module):
    def __init__(self, in_channels: int, hidden: int):
        super().__init__()
        self.y_proj = nn.Sequential(nn.Linear(in_channels, hidden, bias=True), nn.GELU(approximate="tanh"),
                                    nn.Linear(hidden, hidden, bias=True))


class _Attention(nn.Module):
    def __init__(self, dim, heads, eps):
        super().__init__()
        self.dim, self.num_heads, self.head_dim = dim, heads, dim // heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.q_norm = _RMSNormParams(self.head_dim, eps)
        self.k_norm = _RMSNormParams(self.head_dim, eps)
        self.proj = nn.Linear(dim, dim)


class _CrossAttention(nn.Module):
    def __init__(self, dim, heads, eps):
        super().__init__()
        self.dim, self.num_heads, self.head_dim = dim, heads, dim // heads
        self.q_linear = nn.Linear(dim, dim)
        self.kv_linear = nn.Linear(dim, dim * 2)
        self.proj = nn.Linear(dim, dim)
        self.q_norm = _RMSNormParams(self.head_dim, eps)
        self.k_norm = _RMSNormParams(self.head_dim, eps)


class _FFN(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.w1 = nn.Linear(dim, hidden, bias=False)
        self.w2 = nn.Linear(hidden, dim, bias=False)
        self.w3 = nn.Linear(dim, hidden, bias=False)


class LongCatSingleStreamBlock(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        C = cfg.hidden_size
        self.hidden_size = C
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(cfg.adaln_tembed_dim, 6 * C, bias=True))
        self.mod_norm_attn = nn.LayerNorm(C, eps=cfg.norm_eps, elementwise_affine=False)
        self.mod_norm_ffn = nn.LayerNorm(C, eps=cfg.norm_eps, elementwise_affine=False)
        self.pre_crs_attn_norm = nn.LayerNorm(C, eps=cfg.norm_eps, elementwise_affine=True)
        self.attn = _Attention(C, cfg.num_heads, cfg.norm_eps)
        self.cross_attn = _CrossAttention(C, cfg.num_heads, cfg.norm_eps)
        self.ffn = _FFN(C, cfg.ffn_dim)


class _FinalLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        C = cfg.hidden_size
        n_patch = cfg.patch_size[0] * cfg.patch_size[1] * cfg.patch_size[2]
        self.norm_final = nn.LayerNorm(C, eps=cfg.norm_eps, elementwise_affine=False)
        self.linear = nn.Linear(C, n_patch * cfg.out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(cfg.adaln_tembed_dim, 2 * C, bias=True))


class B200DiT(nn.Module):
    """Stands in for ``LongCatVideoTransformer3DModel`` on the TTA path (call site common.py:476-482)."""

    def __init__(self, cfg):
        super().__init__()
        if tuple(cfg.patch_size) != (1, 2, 2) or cfg.in_channels != 16 or cfg.out_channels != 16:
            raise ValueError("B200DiT kernels are specialised for 16 latent channels and patch (1,2,2)")
        self.config = cfg
        self.patch_size = tuple(cfg.patch_size)
        self.text_tokens_zero_pad = cfg.text_tokens_zero_pad
        self.x_embedder = _PatchEmbed3D(cfg.patch_size, cfg.in_channels, cfg.hidden_size)
        self.t_embedder = _TimestepEmbedder(cfg.adaln_tembed_dim, cfg.frequency_embedding_size)
        self.y_embedder = _CaptionEmbedder(cfg.caption_channels, cfg.hidden_size)
        self.blocks = nn.ModuleList([LongCatSingleStreamBlock(cfg) for _ in range(cfg.depth)])
        self.final_layer = _FinalLayer(cfg)
        self.gradient_checkpointing = True  # per-block recompute is how the engine always runs
        self._engine = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def random_init(cls, name: str = "13.6b", seed: int = 0, device="cuda", init_std: float = 0.02, **overrides):
        """Seeded random-init bf16 model created directly on ``device`` (no checkpoints: there is no network)."""
        cfg = make_config(name, **overrides)
        with torch.device("meta"):
            m = cls(cfg)
        m = m.to(BF16).to_empty(device=device)
        g = torch.Generator(device=device).manual_seed(seed)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if p.dim() >= 2 or n.endswith("bias"):
                    p.copy_(torch.randn(p.shape, generator=g, device=device, dtype=torch.float32) * init_std)
                else:
                    p.copy_(1.0 + torch.randn(p.shape, generator=g, device=device, dtype=torch.float32) * init_std)
        m.requires_grad_(False)
        return m

    @classmethod
    def from_oracle(cls, oracle_dit, device="cuda", **overrides):
        """Copy an ``oracle.dit_oracle.OracleDiT`` (tests only hand one in) into bf16 kernel storage."""
        o = oracle_dit.config
        cfg = make_config("tiny", hidden_size=o.hidden_size, depth=o.depth, num_heads=o.num_heads,
                          caption_channels=o.caption_channels, **overrides)
        m = cls(cfg)
        missing, unexpected = m.load_state_dict(oracle_dit.state_dict(), strict=False)
        assert not unexpected and not missing, (missing, unexpected)
        m = m.to(device=device, dtype=BF16)
        m.requires_grad_(False)
        return m

    # ------------------------------------------------------------------ reference-facing API
    def unpatchify(self, x, N_t, N_h, N_w):
        pt, ph, pw = self.patch_size
        B = x.shape[0]
        x = x.view(B, N_t, N_h, N_w, pt, ph, pw, -1).permute(0, 7, 1, 4, 2, 5, 3, 6)
        return x.reshape(B, -1, N_t * pt, N_h * ph, N_w * pw)

    @property
    def engine(self):
        if self._engine is None:
            from .engine import TTAEngine
            self._engine = TTAEngine(self)
        return self._engine

    def forward(self, hidden_states, timestep, encoder_hidden_states, encoder_attention_mask=None,
                num_cond_latents: Optional[int] = 0, **kwargs):
        """Same signature and return value (fp32 [B,16,T,H,W], differentiable wrt adapter parameters) as the
        upstream DiT forward; executed by the sm_100a engine."""
        from .engine import dit_forward_autograd
        return dit_forward_autograd(self, hidden_states, timestep, encoder_hidden_states, encoder_attention_mask,
                                    num_cond_latents or 0)
