"""CPU/any-device restatement of the block-sparse attention definition (test infrastructure only; imported by tests/).

The upstream block-sparse kernel is absent from the reference ("spec to be fixed by us"), so this file
restates OUR definition (longcat_video_tta_b200/bsa.py docstring) independently of the product code: plain loops over
chunks, fp32 masked softmax.  Parity unpinned by the reference.
"""
import math

import torch


def chunk_permutation(T, Hg, Wg, chunk=(4, 4, 8)):
    ct, ch, cw = chunk
    order = []
    for t0 in range(0, T, ct):
        for h0 in range(0, Hg, ch):
            for w0 in range(0, Wg, cw):
                for t in range(t0, t0 + ct):
                    for h in range(h0, h0 + ch):
                        for w in range(w0, w0 + cw):
                            order.append((t * Hg + h) * Wg + w)
    return torch.tensor(order)


def select_mask(q, k, sparsity, n_context_blocks=0, block=128):
    """q, k [n, H, D] fp32 block-major -> [H, nb, nb] bool"""
    n, H, D = q.shape
    nb = n // block
    keep = max(1, math.ceil((1.0 - sparsity) * nb))
    mask = torch.zeros(H, nb, nb, dtype=torch.bool)
    for h in range(H):
        qm = torch.stack([q[i * block:(i + 1) * block, h].double().mean(0) for i in range(nb)])
        km = torch.stack([k[j * block:(j + 1) * block, h].double().mean(0) for j in range(nb)])
        s = qm @ km.t()
        for i in range(nb):
            allowed = range(n_context_blocks) if i < n_context_blocks else range(nb)
            ranked = sorted((j for j in allowed if j != i), key=lambda j: -float(s[i, j]))
            chosen = [i] + ranked[:min(keep, len(allowed)) - 1]
            mask[h, i, chosen] = True
    return mask


def masked_attention(q, k, v, mask, scale, block=128):
    """fp32 reference: q,k,v [n,H,D]; block mask [H,nb,nb] -> o [n,H,D], lse [H,n]"""
    n, H, D = q.shape
    tok = mask.repeat_interleave(block, 1).repeat_interleave(block, 2)       # [H, n, n]
    s = torch.einsum("qhd,khd->hqk", q, k) * scale
    s = s.masked_fill(~tok.to(s.device), float("-inf"))
    lse = torch.logsumexp(s, -1)
    o = torch.einsum("hqk,khd->qhd", torch.softmax(s, -1), v)
    return o, lse


def block_sparse_attention(q, k, v, shape, num_cond_latents, chunk=(4, 4, 8), sparsity=0.9375):
    """Self-attention of the oracle DiT under the block-sparse definition.  q, k, v [B, H, N, D] (after q/k norm + RoPE),
    shape = (T, H', W') token grid, context = the first num_cond_latents frames (must be whole chunks)."""
    B, H, N, D = q.shape
    T, Hg, Wg = shape
    ct, ch, cw = chunk
    block = ct * ch * cw
    if num_cond_latents % ct:
        raise ValueError("the context frames must be whole chunks along T")
    perm = chunk_permutation(T, Hg, Wg, chunk).to(q.device)
    n_ctx = (num_cond_latents // ct) * (Hg // ch) * (Wg // cw)
    outs = []
    for b in range(B):
        qb, kb, vb = (t[b][:, perm].transpose(0, 1) for t in (q, k, v))            # [N, H, D] block-major
        mask = select_mask(qb.detach().float().cpu(), kb.detach().float().cpu(), sparsity, n_ctx, block)
        o, _ = masked_attention(qb.float(), kb.float(), vb.float(), mask.to(q.device), D ** -0.5, block)
        ob = torch.empty_like(o)
        ob[perm] = o                                                                 # back to (t, h, w) row-major
        outs.append(ob.transpose(0, 1).to(q.dtype))
    return torch.stack(outs)
