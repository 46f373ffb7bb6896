"""CPU model of the lazy-reference-maximum online softmax prepared in scratch/variants/attn_fwd_lazy_max_UNTESTED.cu.txt:
block-wise attention for a batch of rows where, after the first K/V block, the row maximum is only computed when the
block's row sum exceeds 2^8 (P rounded to bf16, O and l accumulated in fp32, exactly as the kernel does).  Compared with
exact softmax attention on random, growing-maximum and overflowing inputs; prints how often the exact path ran."""
import torch

BN, THRESH, LIMIT = 128, 8.0, 256.0
LOG2E = 1.4426950408889634


def lazy_attention(q, k, v, scale):
    """q [R,D], k/v [N,D] fp32 (values already bf16-representable) -> o [R,D], lse [R], fraction of exact-path blocks"""
    R, D = q.shape
    s_all = (q @ k.t()) * (scale * LOG2E)          # log2 units, fp32 like the TMEM accumulator
    m = torch.full((R,), float("-inf")); l = torch.zeros(R); o = torch.zeros(R, D)
    exact_blocks = 0
    nb = (k.shape[0] + BN - 1) // BN
    for j in range(nb):
        s = s_all[:, j * BN:(j + 1) * BN]
        vj = v[j * BN:(j + 1) * BN]

        def exact():
            nonlocal m, l, o
            mb = s.max(dim=1).values
            if j == 0:
                m = mb.clone()
                return False
            grow = mb > m + THRESH
            # the kernel decides per warp (32 rows); per row is the finest version of the same rule
            if not grow.any():
                return False
            alpha = torch.where(grow, torch.exp2(m - mb), torch.ones_like(m))
            m = torch.where(grow, mb, m)
            l = l * alpha; o = o * alpha[:, None]
            return True

        if j == 0:
            exact(); exact_blocks += 1
            p = torch.exp2(s - m[:, None])
        else:
            p = torch.exp2(s - m[:, None])
            ls = p.sum(dim=1)
            if (~(ls <= LIMIT)).any():             # catches inf / nan as well
                exact_blocks += 1
                if exact():
                    p = torch.exp2(s - m[:, None])
        l = l + p.sum(dim=1)
        o = o + p.to(torch.bfloat16).float() @ vj
    return o / l[:, None], (m + torch.log2(l)) / LOG2E, exact_blocks / nb


def reference(q, k, v, scale):
    s = (q @ k.t()) * scale
    return torch.softmax(s, dim=-1) @ v, torch.logsumexp(s, dim=-1)


def case(name, kscale):
    g = torch.Generator().manual_seed(0)
    R, N, D = 128, 4096, 128
    q = torch.randn(R, D, generator=g).bfloat16().float()
    k = (torch.randn(N, D, generator=g) * kscale[:, None]).bfloat16().float()
    v = torch.randn(N, D, generator=g).bfloat16().float()
    o, lse, frac = lazy_attention(q, k, v, D ** -0.5)
    ro, rlse = reference(q, k, v, D ** -0.5)
    err = (o - ro).abs().max().item() / ro.abs().max().item()
    cosv = torch.nn.functional.cosine_similarity(o.flatten(), ro.flatten(), dim=0).item()
    print(f"{name:34s} exact-path blocks {frac * 100:5.1f} %   max err / max|ref| {err:.2e}   cosine {cosv:.6f}   "
          f"lse err {(lse - rlse).abs().max().item():.2e}   finite {bool(torch.isfinite(o).all())}")
    assert torch.isfinite(o).all() and cosv > 0.999 and err < 2e-2


nblk = 4096 // BN
case("random N(0,1) keys", torch.ones(4096))
case("magnitude grows block by block", torch.linspace(0.25, 20.0, nblk).repeat_interleave(BN))
case("jumps (x0.25 ... x20 ... x1)", torch.tensor([0.25, 1, 4, 0.5, 10, 1, 2, 20] * (nblk // 8)).repeat_interleave(BN))
case("one huge block in the middle (x200)", torch.cat([torch.ones(2048), torch.full((128,), 200.0), torch.ones(4096 - 2176)]))
case("decaying magnitude (stale-high m)", torch.linspace(20.0, 0.05, nblk).repeat_interleave(BN))
