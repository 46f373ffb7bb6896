"""Block-sparse attention (BASELINE.json configs[4]): the list-driven tcgen05 kernels against an fp32 masked-softmax
reference, plus the host-side chunk permutation and top-k selection against the independent restatement in
oracle/bsa_oracle.py.  Tolerance: bf16 storage, rtol 2e-2 / cosine >= 0.999 (north_star)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF16, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from longcat_video_tta_b200 import ops as o
    o.selfcheck()
    return o


def close(got, ref, rtol=2e-2, min_cos=0.999):
    got, ref = got.float(), ref.float()
    assert torch.isfinite(got).all()
    atol = 2e-2 * ref.abs().max().item() + 1e-6
    c = (torch.dot(got.flatten(), ref.flatten()) / (got.norm() * ref.norm() + 1e-30)).item()
    assert c >= min_cos, f"cosine {c}"
    assert torch.allclose(got, ref, rtol=rtol, atol=atol), f"max err {(got - ref).abs().max().item()} vs atol {atol}"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(BF16)


def random_mask(H, nb, density, seed):
    g = torch.Generator().manual_seed(seed)
    m = torch.rand(H, nb, nb, generator=g) < density
    m |= torch.eye(nb, dtype=torch.bool)[None]
    return m.cuda()


@pytest.mark.parametrize("nb,H,density", [(4, 2, 0.5), (8, 3, 0.3), (16, 2, 0.15), (6, 1, 1.0)])
def test_bsa_forward_and_backward_match_masked_reference(ops, nb, H, density):
    from oracle.bsa_oracle import masked_attention
    from longcat_video_tta_b200.bsa import lists_from_mask
    n, D = nb * 128, 128
    scale = D ** -0.5
    qkv = rnd(n, 3, H, D, seed=nb)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    lists = lists_from_mask(random_mask(H, nb, density, seed=nb))
    o = torch.zeros(n, H, D, dtype=BF16, device="cuda")
    lse = torch.zeros(H, n, dtype=F32, device="cuda")
    ops.attn_bsa_fwd(q, k, v, o, lse, lists.q_off, lists.q_idx, scale)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ro, rl = masked_attention(qf, kf, vf, lists.mask, scale)
    close(o, ro)
    close(lse, rl, rtol=1e-3)
    do = rnd(n, H, D, seed=99)
    dqkv = torch.full((n, 3, H, D), float("nan"), dtype=BF16, device="cuda")
    delta = torch.empty(H, n, dtype=F32, device="cuda")
    ops.attn_bsa_bwd(dqkv[:, 0], dqkv[:, 1], dqkv[:, 2], do, o, lse, delta, q, k, v, lists.q_off, lists.q_idx,
                     lists.k_off, lists.k_idx, scale)
    ro.backward(do.float())
    close(dqkv[:, 0], qf.grad)
    close(dqkv[:, 1], kf.grad)
    close(dqkv[:, 2], vf.grad)


def test_full_mask_equals_dense_attention(ops):
    from longcat_video_tta_b200.bsa import lists_from_mask
    nb, H, D = 5, 2, 128
    n = nb * 128
    qkv = rnd(n, 3, H, D, seed=5)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    lists = lists_from_mask(torch.ones(H, nb, nb, dtype=torch.bool, device="cuda"))
    o1, o2 = (torch.zeros(n, H, D, dtype=BF16, device="cuda") for _ in range(2))
    l1, l2 = (torch.zeros(H, n, dtype=F32, device="cuda") for _ in range(2))
    ops.attn_bsa_fwd(q, k, v, o1, l1, lists.q_off, lists.q_idx, D ** -0.5)
    ops.attn_fwd(q, k, v, o2, l2, [(0, n, n)], D ** -0.5)
    assert torch.allclose(l1, l2, rtol=1e-5, atol=1e-5)
    assert (o1.float() - o2.float()).abs().max().item() <= 2 ** -7


def test_permutation_and_selection_match_oracle(ops):
    from oracle import bsa_oracle as O
    from longcat_video_tta_b200 import bsa
    T, Hg, Wg, H, D = 8, 8, 16, 2, 128
    perm, inv = bsa.block_permutation(T, Hg, Wg)
    assert torch.equal(perm, O.chunk_permutation(T, Hg, Wg))
    assert torch.equal(perm[inv], torch.arange(T * Hg * Wg))
    n = T * Hg * Wg
    nb = n // 128
    g = torch.Generator().manual_seed(3)
    q, k = torch.randn(n, H, D, generator=g), torch.randn(n, H, D, generator=g)
    for sparsity, n_ctx in ((0.75, 0), (0.5, 2), (0.9375, 0)):
        lists = bsa.select_blocks(q.cuda(), k.cuda(), sparsity=sparsity, n_context_blocks=n_ctx)
        want = O.select_mask(q, k, sparsity, n_ctx)
        assert torch.equal(lists.mask.cpu(), want), f"selection differs at sparsity {sparsity}"
        assert bool(lists.mask.diagonal(dim1=1, dim2=2).all())
        if n_ctx:
            assert not bool(lists.mask[:, :n_ctx, n_ctx:].any())
        # CSR lists are the mask, row by row, ascending
        off, idx = lists.q_off.cpu(), lists.q_idx.cpu()
        for h in range(H):
            for i in range(nb):
                row = idx[off[h * nb + i]:off[h * nb + i + 1]]
                assert torch.equal(row.long(), lists.mask[h, i].nonzero().flatten().cpu())
        koff, kidx = lists.k_off.cpu(), lists.k_idx.cpu()
        for h in range(H):
            for j in range(nb):
                col = kidx[koff[h * nb + j]:koff[h * nb + j + 1]]
                assert torch.equal(col.long(), lists.mask[h, :, j].nonzero().flatten().cpu())


def test_bsa_at_720p_geometry_slices(ops):
    """configs[4] geometry: 24 x 48 x 80 token grid = 92 160 tokens = 720 blocks, sparsity 0.9375 (45 key blocks per
    query block), 2 heads; fp32 spot checks on a few query blocks + softmax-rows-sum-to-one property."""
    from longcat_video_tta_b200 import bsa
    T, Hg, Wg, H, D = 24, 48, 80, 2, 128
    n = T * Hg * Wg
    scale = D ** -0.5
    qkv = rnd(n, 3, H, D, seed=7)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    lists = bsa.select_blocks(q, k, sparsity=0.9375, n_context_blocks=(4 * Hg * Wg) // 128)
    nb = n // 128
    assert int(lists.mask[0, nb - 1].sum()) == 45
    o, lse = bsa.bsa_attention(q, k, v, lists, scale)
    o1, _ = bsa.bsa_attention(q, k, torch.ones_like(v), lists, scale)
    assert (o1.float() - 1).abs().max().item() <= 2 ** -7
    for h, i in ((0, 0), (1, 37), (0, nb - 1), (1, 400)):
        cols = lists.mask[h, i].nonzero().flatten()
        kk = torch.cat([k[j * 128:(j + 1) * 128, h] for j in cols.tolist()]).float()
        vv = torch.cat([v[j * 128:(j + 1) * 128, h] for j in cols.tolist()]).float()
        s = (q[i * 128:(i + 1) * 128, h].float() @ kk.t()) * scale
        close(o[i * 128:(i + 1) * 128, h], torch.softmax(s, -1) @ vv)
        close(lse[h, i * 128:(i + 1) * 128], torch.logsumexp(s, -1), rtol=1e-3)


def test_tta_step_with_block_sparse_attention_matches_oracle():
    """configs[4] on the tiny DiT: 4 context + 4 noised latent frames of 32x32 (2 048 tokens = 16 blocks, sparsity 0.5),
    LoRA r=8 with non-zero B: loss and adapter gradients of one fused step vs the oracle DiT running the block-sparse
    definition in plain PyTorch (bf16 and fp32)."""
    import copy
    from oracle import tta_oracle as T
    from oracle.dit_oracle import build_oracle_dit
    from longcat_video_tta_b200 import lora
    from longcat_video_tta_b200.dit import B200DiT
    from longcat_video_tta_b200.stepper import TTAStepper
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    bsa_cfg = dict(chunk=(4, 4, 8), sparsity=0.5)
    g = torch.Generator().manual_seed(21)
    cond, train = torch.randn(1, 16, 4, 32, 32, generator=g), torch.randn(1, 16, 4, 32, 32, generator=g)
    prompt = torch.randn(1, 1, 512, 512, generator=g)
    mask = torch.zeros(1, 512, dtype=torch.int64)
    mask[:, :128] = 1
    sigma, eps = torch.tensor([0.37]), torch.randn(train.shape, generator=g)
    oracle = build_oracle_dit("tiny", seed=0)
    dit = B200DiT.from_oracle(oracle, enable_bsa=True, bsa_params=dict(sparsity=0.5))
    torch.manual_seed(11)
    mods = lora.inject_lora_into_dit(dit, rank=8, alpha=16.0)
    refs = {}
    for name, dtype in (("bf16", BF16), ("fp32", F32)):
        o = copy.deepcopy(oracle)
        for blk in o.blocks:
            blk.attn.bsa = dict(bsa_cfg)
        torch.manual_seed(11)
        omods = T.inject_lora(o, rank=8, alpha=16.0)
        gen = torch.Generator().manual_seed(3)
        with torch.no_grad():
            for m, om in zip(mods, omods):
                b = (torch.randn(m.lora_up.weight.shape, generator=gen) * 0.02).to(BF16)
                if name == "bf16":
                    m.lora_up.weight.copy_(b.cuda())
                om.lora_up.weight.copy_(b.float())
                om.lora_down.weight.copy_(m.lora_down.weight.float().cpu())
        o = o.to(dtype).cuda()
        params = T.lora_parameters(omods)
        for p in params:
            p.requires_grad_(True)
        cast = (lambda t: t.to(BF16).cuda()) if dtype == BF16 else (lambda t: t.to(BF16).float().cuda())
        loss = T.fm_loss_given(o, cast(cond), cast(train), cast(prompt), mask.cuda(), sigma.cuda(), cast(eps), dtype)
        refs[name] = (loss.item(), [x.float() for x in torch.autograd.grad(loss, params)])
    st = TTAStepper(dit)
    loss = st.forward_backward(cond.to(BF16).cuda(), train.to(BF16).cuda(), prompt.to(BF16).cuda(), mask.cuda(),
                               sigma.cuda(), eps.to(BF16).cuda()).item()
    mine = [g_.float() for site in dit.engine.lora_sites() for g_ in site.param_grads()]
    lists = dit.engine.ws.bsa_lists[0]
    assert lists is not None and abs(lists.density - (8 * 8 + 8 * 8) / 256) < 1e-6
    flat = lambda gs: torch.cat([x.flatten().cpu() for x in gs])
    cosf = lambda a, b: (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()
    c_bf, c_32, c_ref = cosf(flat(mine), flat(refs["bf16"][1])), cosf(flat(mine), flat(refs["fp32"][1])), \
        cosf(flat(refs["bf16"][1]), flat(refs["fp32"][1]))
    print(f"BSA step: loss mine {loss:.5f} bf16-torch {refs['bf16'][0]:.5f} fp32 {refs['fp32'][0]:.5f}; grads mine~bf16 {c_bf:.5f} "
          f"mine~fp32 {c_32:.5f} bf16~fp32 {c_ref:.5f}")
    assert abs(loss - refs["fp32"][0]) <= 2e-2 * refs["fp32"][0]
    assert c_bf > 0.999
    assert c_32 >= min(0.999, c_ref - 5e-3)
