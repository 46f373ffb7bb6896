"""Oracle restatement of the UMT5 encoder forward (plain PyTorch, CPU, any float dtype).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): imported by ``tests/`` and by the checker legs of the scratch
benchmarks, never by ``longcat_video_tta_b200``.

What it restates: the text half of SURVEY 8(f) row 4.  The reference's ``encode_prompt``
(``delta_experiment/scripts/common.py:228-255``) tokenises a prompt and calls
``text_encoder(input_ids, mask).last_hidden_state`` on a ``transformers.UMT5EncoderModel``
(``common.py:33,62-64``; the reference environment pins transformers 4.41.0, ``env_setup/01_setup_longcat_env.sbatch``).
The arithmetic therefore lives in a third-party dependency: ``transformers/models/umt5/modeling_umt5.py``
(UMT5Stack / UMT5Block / UMT5Attention / UMT5LayerNorm / UMT5DenseGatedActDense, NewGELUActivation).

Parity is PINNED: transformers (5.5 in this image) is importable here, and ``oracle/make_golden_umt5.py`` runs the real
``UMT5EncoderModel`` on the weights of :func:`tiny_state` and commits its outputs as ``tests/golden/umt5_tiny.pt``;
``tests/test_umt5_oracle_cpu.py`` holds this file to them at 1e-5.

Published algorithm (encoder, eval mode, no dropout), per layer ``l`` on hidden states ``x [B, N, d_model]``:
    h  = rms(x) * ln0_w                       rms(x) = x * rsqrt(mean(x^2, -1) + eps)   (no mean subtraction, no bias)
    q, k, v = h Wq^T, h Wk^T, h Wv^T          split into ``num_heads`` heads of ``d_kv``
    s  = q k^T + bias_l[head, bucket(key - query)] + (1 - mask[key]) * finfo.min         (NO 1/sqrt(d_kv))
    x  = x + (softmax(s) v) Wo^T
    h  = rms(x) * ln1_w
    x  = x + (gelu_new(h Wi0^T) * (h Wi1^T)) Wo_ff^T
and ``rms(x) * final_w`` at the end.  Every layer owns its relative-position table (that is the "U" of UMT5);
``bucket`` is the bidirectional T5 bucketing: half of the buckets per sign, exact up to ``num_buckets/4``, then
logarithmic up to ``max_distance``.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

TINY = dict(vocab_size=384, d_model=256, d_kv=64, d_ff=512, num_layers=2, num_heads=4,
            relative_attention_num_buckets=32, relative_attention_max_distance=128, layer_norm_epsilon=1e-6)
XXL = dict(vocab_size=256384, d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64,
           relative_attention_num_buckets=32, relative_attention_max_distance=128, layer_norm_epsilon=1e-6)


def tiny_state(cfg: dict = TINY, seed: int = 0, bf16_values: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded weights under transformers' UMT5EncoderModel state-dict names (fp32 tensors; with ``bf16_values`` every value
    is representable in bf16, so the bf16 copy the GPU path stores is lossless)."""
    g = torch.Generator().manual_seed(seed)
    d, inner, ff = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"]

    def rn(*shape, std):
        t = torch.randn(*shape, generator=g) * std
        return t.bfloat16().float() if bf16_values else t

    st = {"shared.weight": rn(cfg["vocab_size"], d, std=1.0)}
    st["encoder.embed_tokens.weight"] = st["shared.weight"]
    for l in range(cfg["num_layers"]):
        a, f = f"encoder.block.{l}.layer.0.", f"encoder.block.{l}.layer.1."
        st[a + "SelfAttention.q.weight"] = rn(inner, d, std=2.0 * (d * cfg["d_kv"]) ** -0.5)
        st[a + "SelfAttention.k.weight"] = rn(inner, d, std=2.0 * d ** -0.5)
        st[a + "SelfAttention.v.weight"] = rn(inner, d, std=d ** -0.5)
        st[a + "SelfAttention.o.weight"] = rn(d, inner, std=inner ** -0.5)
        st[a + "SelfAttention.relative_attention_bias.weight"] = rn(cfg["relative_attention_num_buckets"],
                                                                    cfg["num_heads"], std=1.0)
        st[a + "layer_norm.weight"] = 1.0 + rn(d, std=0.1)
        st[f + "DenseReluDense.wi_0.weight"] = rn(ff, d, std=d ** -0.5)
        st[f + "DenseReluDense.wi_1.weight"] = rn(ff, d, std=d ** -0.5)
        st[f + "DenseReluDense.wo.weight"] = rn(d, ff, std=ff ** -0.5)
        st[f + "layer_norm.weight"] = 1.0 + rn(d, std=0.1)
    st["encoder.final_layer_norm.weight"] = 1.0 + rn(d, std=0.1)
    return st


def relative_position_bucket(rel: torch.Tensor, num_buckets: int, max_distance: int) -> torch.Tensor:
    """modeling_umt5.py UMT5Attention._relative_position_bucket, encoder (bidirectional) branch; rel = key - query"""
    half = num_buckets // 2
    bucket = (rel > 0).long() * half
    dist = rel.abs()
    exact = half // 2
    scaled = torch.log(dist.float() / exact) / math.log(max_distance / exact)
    scaled = scaled * (half - exact)
    far = torch.clamp(exact + scaled.long(), max=half - 1)
    return bucket + torch.where(dist < exact, dist, far)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """UMT5LayerNorm.forward: variance in fp32; a half-precision weight makes the normalised value round to that dtype
    before the scale"""
    var = x.float().pow(2).mean(-1, keepdim=True)
    h = x * torch.rsqrt(var + eps)
    if w.dtype in (torch.float16, torch.bfloat16):
        h = h.to(w.dtype)
    return w * h


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    """transformers.activations.NewGELUActivation (tanh form), what ``feed_forward_proj="gated-gelu"`` selects"""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


@torch.no_grad()
def umt5_encode(state: Dict[str, torch.Tensor], cfg: dict, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                dtype=torch.float32) -> torch.Tensor:
    """last_hidden_state [B, N, d_model] of UMT5EncoderModel(input_ids, attention_mask) with all parameters in ``dtype``"""
    W = {k: v.to(dtype) for k, v in state.items()}
    B, N = input_ids.shape
    H, dk, eps = cfg["num_heads"], cfg["d_kv"], cfg["layer_norm_epsilon"]
    x = W["encoder.embed_tokens.weight"][input_ids]
    key_bias = (1.0 - attention_mask[:, None, None, :].to(dtype)) * torch.finfo(dtype).min
    pos = torch.arange(N)
    buckets = relative_position_bucket(pos[None, :] - pos[:, None], cfg["relative_attention_num_buckets"],
                                       cfg["relative_attention_max_distance"])
    for l in range(cfg["num_layers"]):
        a, f = f"encoder.block.{l}.layer.0.", f"encoder.block.{l}.layer.1."
        h = rms_norm(x, W[a + "layer_norm.weight"], eps)
        q = (h @ W[a + "SelfAttention.q.weight"].T).view(B, N, H, dk).transpose(1, 2)
        k = (h @ W[a + "SelfAttention.k.weight"].T).view(B, N, H, dk).transpose(1, 2)
        v = (h @ W[a + "SelfAttention.v.weight"].T).view(B, N, H, dk).transpose(1, 2)
        s = q @ k.transpose(2, 3)
        bias = W[a + "SelfAttention.relative_attention_bias.weight"][buckets].permute(2, 0, 1)[None]
        s = s + (bias + key_bias)
        p = torch.softmax(s.float(), -1).to(s.dtype)
        o = (p @ v).transpose(1, 2).reshape(B, N, H * dk)
        x = x + o @ W[a + "SelfAttention.o.weight"].T
        h = rms_norm(x, W[f + "layer_norm.weight"], eps)
        g = gelu_new(h @ W[f + "DenseReluDense.wi_0.weight"].T) * (h @ W[f + "DenseReluDense.wi_1.weight"].T)
        x = x + g @ W[f + "DenseReluDense.wo.weight"].T
    return rms_norm(x, W["encoder.final_layer_norm.weight"], eps)


def tiny_inputs(cfg: dict = TINY, batch: int = 2, n_tok: int = 96, seed: int = 1):
    """token ids + right-padded masks (like the tokenizer's padding="max_length"): item b keeps n_tok*(2+b)//4 tokens...
    the last item is left unpadded when batch > 2"""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(2, cfg["vocab_size"], (batch, n_tok), generator=g)
    mask = torch.zeros(batch, n_tok, dtype=torch.long)
    for b in range(batch):
        keep = n_tok if b >= 2 else max(1, n_tok * (2 + b) // 4 + 3 * b)
        mask[b, :keep] = 1
        ids[b, keep:] = 0          # the pad id
    return ids, mask
