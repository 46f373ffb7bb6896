"""Block-sparse self-attention for the 720p refinement stage (BASELINE.json configs[4]; upstream ``enable_bsa`` /
``bsa_params``, delta_experiment/scripts/common.py:71-74).

The upstream kernel and its parameters are not vendored by the reference, so the semantics are OUR
definition, chosen to sit on the tcgen05 tile shape:

  * tokens of the (T, H', W') latent grid are grouped into 3-D chunks of ``chunk = (4, 4, 8)`` = 128 tokens -- one MMA
    tile of rows -- (SURVEY A.9 recalls 4x4x4 = 64-token chunks upstream; with 128 the key budget per query is the
    same at the same sparsity and no tile is half empty);
  * every query chunk attends the ``ceil((1 - sparsity) * n_chunks)`` key chunks with the highest mean-pooled score
    ``mean(q) . mean(k)`` (per head, after q/k RMSNorm + RoPE), always including itself;
  * with a clean-context / noised split the context chunks only see context chunks (the dense path's segments).

``block_permutation`` gives the token order the kernels expect (block-major), ``select_blocks`` the CSR lists of
``b200tta_attn_bsa_fwd/_bwd``.  Host-side index work only: the attention itself runs in csrc/attn_fwd.cu / attn_bwd.cu.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

BLOCK = 128


def block_permutation(T: int, Hg: int, Wg: int, chunk: Tuple[int, int, int] = (4, 4, 8), device="cpu"):
    """perm[i] = row-major (t, h, w) token index that sits at block-major position i; inv[perm[i]] = i."""
    ct, ch, cw = chunk
    if ct * ch * cw != BLOCK:
        raise ValueError(f"chunk {chunk} must hold {BLOCK} tokens")
    if T % ct or Hg % ch or Wg % cw:
        raise ValueError(f"token grid ({T},{Hg},{Wg}) is not divisible by the chunk {chunk}")
    idx = torch.arange(T * Hg * Wg, device=device).view(T // ct, ct, Hg // ch, ch, Wg // cw, cw)
    perm = idx.permute(0, 2, 4, 1, 3, 5).reshape(-1)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(perm.numel(), device=device)
    return perm, inv


@dataclass
class BlockLists:
    mask: torch.Tensor     # [H, nb, nb] bool: query block i of head h attends key block j
    q_off: torch.Tensor    # int32 [H * nb + 1]
    q_idx: torch.Tensor    # int32 [nnz]
    k_off: torch.Tensor
    k_idx: torch.Tensor

    @property
    def density(self) -> float:
        return float(self.mask.float().mean())


def lists_from_mask(mask: torch.Tensor) -> BlockLists:
    """CSR lists (ascending block indices) of a [H, nb, nb] boolean block mask and of its transpose."""
    def csr(m):
        counts = m.sum(-1).reshape(-1)
        off = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=m.device)
        off[1:] = torch.cumsum(counts, 0).to(torch.int32)
        idx = m.nonzero(as_tuple=False)[:, 2].to(torch.int32).contiguous()
        return off, idx
    q_off, q_idx = csr(mask)
    k_off, k_idx = csr(mask.transpose(1, 2).contiguous())
    return BlockLists(mask, q_off, q_idx, k_off, k_idx)


def select_blocks(q: torch.Tensor, k: torch.Tensor, sparsity: float = 0.9375, n_context_blocks: int = 0) -> BlockLists:
    """q, k: [n_tok, H, D] in block-major order (after RMSNorm + RoPE).  Returns the block lists of the definition above.
    No host synchronisation: every row keeps a number of blocks that is known up front (``row_counts``)."""
    n, H, D = q.shape
    if n % BLOCK:
        raise ValueError(f"n_tok={n} must be a multiple of {BLOCK}")
    nb = n // BLOCK
    keep, keep_ctx = row_counts(nb, sparsity, n_context_blocks)
    # (mean with dtype=float32 up-casts inside the reduction: no fp32 copy of the 755 MB q / k tensors)
    qm = q.reshape(nb, BLOCK, H, D).mean(1, dtype=torch.float32)       # [nb, H, D]
    km = k.reshape(nb, BLOCK, H, D).mean(1, dtype=torch.float32)
    score = torch.einsum("ihd,jhd->hij", qm, km)          # [H, nb, nb]
    eye = torch.eye(nb, dtype=torch.bool, device=q.device)[None]
    score = score.masked_fill(eye, float("inf"))          # the chunk itself is always kept
    if n_context_blocks > 0:
        score[:, :n_context_blocks, n_context_blocks:] = float("-inf")   # context queries never see noised keys
    top = score.topk(keep, dim=-1).indices
    mask = torch.zeros(H, nb, nb, dtype=torch.bool, device=q.device)
    mask.scatter_(2, top, True)
    if n_context_blocks > 0:
        mask[:, :n_context_blocks, n_context_blocks:] = False
    nnz = H * (n_context_blocks * keep_ctx + (nb - n_context_blocks) * keep)
    # query lists: the per-row counts are fixed, so the offsets are analytic
    counts = torch.full((nb,), keep, dtype=torch.int64)
    counts[:n_context_blocks] = keep_ctx
    q_off = torch.zeros(H * nb + 1, dtype=torch.int64)
    q_off[1:] = torch.cumsum(counts.repeat(H), 0)
    q_idx = mask.nonzero_static(size=nnz)[:, 2].to(torch.int32).contiguous()
    mt = mask.transpose(1, 2).contiguous()
    k_off = torch.zeros(H * nb + 1, dtype=torch.int32, device=q.device)
    k_off[1:] = torch.cumsum(mt.sum(-1).reshape(-1), 0).to(torch.int32)
    k_idx = mt.nonzero_static(size=nnz)[:, 2].to(torch.int32).contiguous()
    return BlockLists(mask, q_off.to(device=q.device, dtype=torch.int32), q_idx, k_off, k_idx)


def row_counts(nb: int, sparsity: float, n_context_blocks: int = 0) -> Tuple[int, int]:
    """(key blocks kept per noised query block, per context query block)"""
    keep = min(nb, max(1, math.ceil((1.0 - sparsity) * nb)))
    return keep, (min(keep, n_context_blocks) if n_context_blocks > 0 else keep)


def bsa_attention(q, k, v, lists: BlockLists, softmax_scale: Optional[float] = None):
    """Forward helper: returns (o [n,H,D] bf16, lse [H,n] f32) for block-major q/k/v."""
    from . import ops
    n, H, D = q.shape
    o = torch.empty(n, H, D, dtype=torch.bfloat16, device=q.device)
    lse = torch.empty(H, n, dtype=torch.float32, device=q.device)
    ops.attn_bsa_fwd(q, k, v, o, lse, lists.q_off, lists.q_idx, softmax_scale if softmax_scale is not None else D ** -0.5)
    return o, lse
