"""CPU: the UMT5 oracle against the committed outputs of transformers' UMT5EncoderModel (tests/golden/umt5_tiny.pt, made by
oracle/make_golden_umt5.py), and the host-side logic of longcat_video_tta_b200.text_encoder (bucketing, encode_prompt's
contract, the refusal to run without a B200)."""
from types import SimpleNamespace

import pytest
import torch

from oracle import umt5_oracle as uo


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(golden_dir / "umt5_tiny.pt")


def test_oracle_matches_transformers_outputs(golden):
    st = uo.tiny_state(golden["cfg"], golden["state_seed"])
    for c in golden["cases"]:
        y = uo.umt5_encode(st, golden["cfg"], c["input_ids"], c["attention_mask"])
        ref = c["last_hidden_state_fp32"]
        assert y.shape == ref.shape
        assert torch.allclose(y, ref, rtol=0, atol=1e-5 * ref.abs().max().item())
        yb = uo.umt5_encode(st, golden["cfg"], c["input_ids"], c["attention_mask"], dtype=torch.bfloat16)
        rb = c["last_hidden_state_bf16"].float()
        # a bf16 model: same rounding points (norm, projections, softmax weights); allow one-ulp reorderings
        assert (yb.float() - rb).norm() / rb.norm() < 5e-3


def test_oracle_matches_live_transformers_when_importable(golden):
    tr = pytest.importorskip("transformers")
    if not hasattr(tr, "UMT5EncoderModel"):
        pytest.skip("this transformers has no UMT5")
    from oracle.make_golden_umt5 import hf_model
    cfg = dict(uo.TINY, num_layers=1, d_ff=256, vocab_size=128)
    st = uo.tiny_state(cfg, seed=5)
    ids, mask = uo.tiny_inputs(cfg, batch=2, n_tok=40, seed=9)
    with torch.no_grad():
        ref = hf_model(cfg, st, torch.float32)(ids, mask).last_hidden_state
    y = uo.umt5_encode(st, cfg, ids, mask)
    assert torch.allclose(y, ref, rtol=0, atol=1e-5 * ref.abs().max().item())


def test_padding_does_not_leak_into_valid_rows(golden):
    """property of the masked softmax: what sits in padded positions cannot change the rows of real tokens"""
    cfg = golden["cfg"]
    st = uo.tiny_state(cfg, golden["state_seed"])
    c = golden["cases"][0]
    ids, mask = c["input_ids"].clone(), c["attention_mask"]
    y0 = uo.umt5_encode(st, cfg, ids, mask)
    ids[mask == 0] = 7
    y1 = uo.umt5_encode(st, cfg, ids, mask)
    keep = mask.bool()
    assert torch.allclose(y0[keep], y1[keep], rtol=0, atol=1e-5)
    assert not torch.allclose(y0[~keep], y1[~keep], atol=1e-3)


@pytest.mark.parametrize("buckets,max_distance", [(32, 128), (16, 64), (64, 256)])
def test_product_bucketing_is_the_oracles(buckets, max_distance):
    from longcat_video_tta_b200.text_encoder import _bucket_of_distance
    rel = torch.arange(-700, 701)
    got = _bucket_of_distance(rel, buckets, max_distance)
    ref = uo.relative_position_bucket(rel, buckets, max_distance)
    assert torch.equal(got, ref)
    assert int(got.min()) == 0 and int(got.max()) == buckets - 1
    # known answers of the published scheme (32 buckets, max distance 128): own bucket below 8, sign in the upper half
    if (buckets, max_distance) == (32, 128):
        t = dict(zip(rel.tolist(), got.tolist()))
        assert [t[-d] for d in range(8)] == list(range(8))
        assert [t[d] for d in range(1, 8)] == [16 + d for d in range(1, 8)]
        assert t[-8] == 8 and t[-127] == 15 and t[-128] == 15 and t[-700] == 15 and t[700] == 31


def test_encode_prompt_contract():
    """common.py:228-255: tokenizer called with padding to max_length; embeds [1, 1, N, C] in `dtype`, mask [1, N]"""
    from longcat_video_tta_b200.text_encoder import encode_prompt
    seen = {}

    def tokenizer(texts, **kw):
        seen.update(kw, texts=texts)
        n = kw["max_length"]
        ids = torch.zeros(1, n, dtype=torch.long)
        ids[0, :5] = torch.tensor([11, 12, 13, 14, 1])
        m = torch.zeros(1, n, dtype=torch.long)
        m[0, :5] = 1
        return SimpleNamespace(input_ids=ids, attention_mask=m)

    def encoder(ids, mask):
        seen["enc"] = (ids.clone(), mask.clone())
        return SimpleNamespace(last_hidden_state=ids.float()[..., None].expand(-1, -1, 8) * 0.5)

    emb, mask = encode_prompt(tokenizer, encoder, "a cat", device="cpu", dtype=torch.bfloat16, max_length=16)
    assert seen["texts"] == ["a cat"] and seen["padding"] == "max_length" and seen["truncation"] is True
    assert seen["add_special_tokens"] is True and seen["return_attention_mask"] is True and seen["return_tensors"] == "pt"
    assert emb.shape == (1, 1, 16, 8) and emb.dtype == torch.bfloat16
    assert mask.shape == (1, 16) and int(mask.sum()) == 5
    assert float(emb[0, 0, 0, 0]) == 5.5 and float(emb[0, 0, 6, 0]) == 0.0


def test_encoder_refuses_to_run_without_a_b200():
    from longcat_video_tta_b200 import _lib
    from longcat_video_tta_b200.text_encoder import B200UMT5Encoder
    cfg = dict(uo.TINY, num_layers=1)
    st = uo.tiny_state(cfg, seed=0)
    kw = {k: v for k, v in cfg.items() if k != "vocab_size"}
    with pytest.raises(_lib.B200TTAError):
        B200UMT5Encoder(st, device="cpu", **kw)
    with pytest.raises(NotImplementedError):
        B200UMT5Encoder(st, device="cpu", **dict(kw, d_kv=32))
