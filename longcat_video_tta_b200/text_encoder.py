"""UMT5 text encoding on the B200 (text half).

Mirrors the reference's ``encode_prompt`` (delta_experiment/scripts/common.py:228-255): tokenise with
``padding="max_length"``, run the encoder, return ``(last_hidden_state.to(dtype)[:, None], attention_mask)`` -- the
``prompt_embeds [B, 1, N, C]`` / ``prompt_mask [B, N]`` pair every TTA loop takes.  The encoder itself is
:class:`B200UMT5Encoder`, a drop-in for the ``transformers.UMT5EncoderModel`` the reference loads
(common.py:62-64): same call (``encoder(input_ids, attention_mask).last_hidden_state``), weights taken from an existing
model (``from_hf``) or from a state dict under transformers' parameter names.

Data layout: hidden states are bf16 ``[B * N, d_model]`` rows; per layer one tcgen05 GEMM for q|k|v (weights
concatenated once at construction), the attention kernel on strided views of that buffer, the output projection with the
residual add in its epilogue, one GEMM over wi_0 | wi_1 with the gated GELU in its epilogue (the two weights are read in
place), and the down projection with the residual add.  The per-layer relative-position tables are expanded once per
sequence length to ``[heads, 2 N - 1]`` (bucketing is index arithmetic on 2 N - 1 distances, done on the host).

No CPU path: construction and calls require a B200 (the C ABI refuses anything else).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch

from . import ops

_A = "encoder.block.{}.layer.0."
_F = "encoder.block.{}.layer.1."


def _bucket_of_distance(rel: torch.Tensor, num_buckets: int, max_distance: int) -> torch.Tensor:
    """bidirectional T5 bucketing of rel = key - query (int64, CPU): sign selects the half, distances below
    num_buckets/4 keep their own bucket, the rest share logarithmic bins up to max_distance"""
    per_sign = num_buckets // 2
    n_exact = per_sign // 2
    dist = rel.abs()
    log_bin = torch.log(dist.float() / n_exact) / math.log(max_distance / n_exact) * (per_sign - n_exact)
    coarse = (n_exact + log_bin.long()).clamp(max=per_sign - 1)
    return torch.where(rel > 0, per_sign, 0) + torch.where(dist < n_exact, dist, coarse)


class B200UMT5Encoder:
    """Encoder half of UMT5 (gated-GELU feed-forward, one relative-position table per layer) on the b200tta operators."""

    def __init__(self, state: Dict[str, torch.Tensor], *, d_model: int, d_kv: int, d_ff: int, num_layers: int,
                 num_heads: int, relative_attention_num_buckets: int = 32, relative_attention_max_distance: int = 128,
                 layer_norm_epsilon: float = 1e-6, device="cuda", **_unused):
        if d_kv != 64:
            raise NotImplementedError(f"b200tta_t5_attn is built for d_kv = 64 (UMT5-xxl), got {d_kv}")
        if d_model % 64 or (num_heads * d_kv) % 64 or d_ff % 128 or d_model > 4096:
            raise NotImplementedError("need d_model % 64 == 0 (<= 4096), d_ff % 128 == 0")
        self.cfg = SimpleNamespace(d_model=d_model, d_kv=d_kv, d_ff=d_ff, num_layers=num_layers, num_heads=num_heads,
                                   buckets=relative_attention_num_buckets, max_distance=relative_attention_max_distance,
                                   eps=layer_norm_epsilon)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ops._lib.B200TTAError("B200UMT5Encoder runs on a B200 only (no CPU fallback)")

        def w(name):
            t = state[name]
            return t.detach().to(self.device, torch.bfloat16).contiguous()

        self.embed = w("encoder.embed_tokens.weight" if "encoder.embed_tokens.weight" in state else "shared.weight")
        self.layers = []
        for l in range(num_layers):
            a, f = _A.format(l), _F.format(l)
            self.layers.append(SimpleNamespace(
                ln0=w(a + "layer_norm.weight"),
                wqkv=torch.cat([w(a + "SelfAttention.q.weight"), w(a + "SelfAttention.k.weight"),
                                w(a + "SelfAttention.v.weight")], 0).contiguous(),
                wo=w(a + "SelfAttention.o.weight"),
                rel_table=state[a + "SelfAttention.relative_attention_bias.weight"].detach().float().cpu(),
                ln1=w(f + "layer_norm.weight"),
                wi0=w(f + "DenseReluDense.wi_0.weight"), wi1=w(f + "DenseReluDense.wi_1.weight"),
                wo_ff=w(f + "DenseReluDense.wo.weight")))
        self.final_ln = w("encoder.final_layer_norm.weight")
        self._rel = {}       # n_tok -> [layers, heads, 2 n_tok - 1] f32 on the device
        self._ws = {}        # (batch, n_tok) -> workspace

    @classmethod
    def from_hf(cls, model, device="cuda") -> "B200UMT5Encoder":
        """wrap a transformers.UMT5EncoderModel (the object common.py:62-64 loads)"""
        c = model.config
        return cls(model.state_dict(), d_model=c.d_model, d_kv=c.d_kv, d_ff=c.d_ff, num_layers=c.num_layers,
                   num_heads=c.num_heads, relative_attention_num_buckets=c.relative_attention_num_buckets,
                   relative_attention_max_distance=c.relative_attention_max_distance,
                   layer_norm_epsilon=c.layer_norm_epsilon, device=device)

    # -- per-shape state -------------------------------------------------------------------------------------------
    def _rel_bias(self, n_tok: int) -> torch.Tensor:
        t = self._rel.get(n_tok)
        if t is None:
            c = self.cfg
            rel = torch.arange(-(n_tok - 1), n_tok)
            b = _bucket_of_distance(rel, c.buckets, c.max_distance)                      # [2 n - 1]
            t = torch.stack([ly.rel_table[b].t().contiguous() for ly in self.layers])      # [L, heads, 2 n - 1]
            self._rel[n_tok] = t = t.to(self.device)
        return t

    def _workspace(self, batch: int, n_tok: int):
        ws = self._ws.get((batch, n_tok))
        if ws is None:
            c, rows = self.cfg, batch * n_tok
            inner = c.num_heads * c.d_kv

            def buf(cols):
                return torch.empty(rows, cols, dtype=torch.bfloat16, device=self.device)

            ws = SimpleNamespace(x=buf(c.d_model), x2=buf(c.d_model), h=buf(c.d_model), qkv=buf(3 * inner), o=buf(inner),
                                 g=buf(c.d_ff), out=buf(c.d_model))
            self._ws = {(batch, n_tok): ws}        # one geometry at a time (prompts are padded to max_length)
        return ws

    # -- forward ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None):
        if input_ids.dim() != 2:
            raise ValueError(f"input_ids must be [batch, tokens], got {tuple(input_ids.shape)}")
        B, N = input_ids.shape
        if attention_mask is not None and tuple(attention_mask.shape) != (B, N):
            raise ValueError(f"attention_mask {tuple(attention_mask.shape)} does not match input_ids {(B, N)}")
        c, rows = self.cfg, B * N
        inner = c.num_heads * c.d_kv
        ids = input_ids.to(self.device, torch.int64).reshape(-1).contiguous()
        if int(ids.min()) < 0 or int(ids.max()) >= self.embed.shape[0]:
            raise IndexError("input_ids outside the vocabulary")
        valid = None if attention_mask is None else (attention_mask.to(self.device) != 0).to(torch.int32).contiguous()
        ws, rel = self._workspace(B, N), self._rel_bias(N)
        x, x2 = ws.x, ws.x2
        ops.gather_rows(x, self.embed, ids)
        q, k, v = ws.qkv[:, :inner], ws.qkv[:, inner:2 * inner], ws.qkv[:, 2 * inner:]
        for l, ly in enumerate(self.layers):
            ops.t5_rmsnorm(ws.h, x, ly.ln0, c.eps)
            ops.gemm(rows, 3 * inner, [(ws.h, ly.wqkv, c.d_model, False, None)], ops.epi(ops.EPI_STORE, ws.qkv))
            ops.t5_attn(ws.o, q, k, v, rel[l], valid, N, c.num_heads, B)
            ops.gemm(rows, c.d_model, [(ws.o, ly.wo, inner, False, None)], ops.epi(ops.EPI_GATE_RESID, x2, resid=x))
            ops.t5_rmsnorm(ws.h, x2, ly.ln1, c.eps)
            ops.gemm(rows, 2 * c.d_ff, [(ws.h, ly.wi0, c.d_model, False, ly.wi1)], ops.epi(ops.EPI_GEGLU, ws.g))
            ops.gemm(rows, c.d_model, [(ws.g, ly.wo_ff, c.d_ff, False, None)], ops.epi(ops.EPI_GATE_RESID, x, resid=x2))
        ops.t5_rmsnorm(ws.out, x, self.final_ln, c.eps)
        return SimpleNamespace(last_hidden_state=ws.out.view(B, N, c.d_model).clone())

    forward = __call__


def encode_prompt(tokenizer, text_encoder, prompt: str, device: str = "cuda", dtype: torch.dtype = torch.bfloat16,
                  max_length: int = 512) -> Tuple[torch.Tensor, torch.Tensor]:
    """common.py:228-255 with the same arguments and return value; ``text_encoder`` is a :class:`B200UMT5Encoder` (or
    anything with the UMT5EncoderModel call)."""
    tok = tokenizer([prompt], return_tensors="pt", return_attention_mask=True, add_special_tokens=True,
                    padding="max_length", truncation=True, max_length=max_length)
    mask = tok.attention_mask.to(device)
    hidden = text_encoder(tok.input_ids.to(device), mask).last_hidden_state      # [1, max_length, d_model]
    return hidden.to(device=device, dtype=dtype)[:, None], mask
