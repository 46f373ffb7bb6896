"""Host-side logic of the block-sparse attention variant (no GPU): chunk permutation, top-k block selection and CSR lists
against the independent restatement in oracle/bsa_oracle.py."""
import pytest
import torch


def test_block_permutation_is_chunk_major_and_invertible():
    from oracle import bsa_oracle as O
    from longcat_video_tta_b200 import bsa
    for T, Hg, Wg in ((4, 4, 8), (8, 8, 16), (4, 12, 24)):
        perm, inv = bsa.block_permutation(T, Hg, Wg)
        assert torch.equal(perm, O.chunk_permutation(T, Hg, Wg))
        assert torch.equal(perm[inv], torch.arange(T * Hg * Wg))
        # the 128 tokens of a block span exactly 4 frames x 4 rows x 8 columns
        blk = perm[:128]
        t, h, w = blk // (Hg * Wg), (blk // Wg) % Hg, blk % Wg
        assert (t.max() - t.min(), h.max() - h.min(), w.max() - w.min()) == (3, 3, 7)
    with pytest.raises(ValueError):
        bsa.block_permutation(5, 8, 16)


@pytest.mark.parametrize("sparsity,n_ctx", [(0.75, 0), (0.5, 2), (0.9375, 0), (0.0, 1)])
def test_selection_matches_oracle_and_lists_are_consistent(sparsity, n_ctx):
    from oracle import bsa_oracle as O
    from longcat_video_tta_b200 import bsa
    n, H, D = 8 * 128, 2, 128
    g = torch.Generator().manual_seed(3)
    q, k = torch.randn(n, H, D, generator=g), torch.randn(n, H, D, generator=g)
    lists = bsa.select_blocks(q, k, sparsity=sparsity, n_context_blocks=n_ctx)
    assert torch.equal(lists.mask, O.select_mask(q, k, sparsity, n_ctx))
    nb = n // 128
    assert bool(lists.mask.diagonal(dim1=1, dim2=2).all())          # the chunk itself is always attended
    if n_ctx:
        assert not bool(lists.mask[:, :n_ctx, n_ctx:].any())        # context queries never see noised keys
    assert lists.q_off.dtype == torch.int32 and int(lists.q_off[-1]) == int(lists.mask.sum()) == lists.q_idx.numel()
    for h in range(H):
        for i in range(nb):
            row = lists.q_idx[lists.q_off[h * nb + i]:lists.q_off[h * nb + i + 1]].long()
            assert torch.equal(row, lists.mask[h, i].nonzero().flatten())
            col = lists.k_idx[lists.k_off[h * nb + i]:lists.k_off[h * nb + i + 1]].long()
            assert torch.equal(col, lists.mask[h, :, i].nonzero().flatten())
