"""GPU parity of the UMT5 text-encoder path (text half) through the C ABI: the two kernels against plain
fp32 PyTorch, the GEGLU / residual GEMM epilogues, and the whole encoder against the committed outputs of transformers'
UMT5EncoderModel (tests/golden/umt5_tiny.pt).  Tolerance: the encoder stores bf16 activations like the bf16 model the
reference runs, so the bar against the fp32 outputs is the bf16 transformers model's own distance (x1.25) and the
north-star cosine > 0.999; the kernels alone are held to rtol 2e-2."""
import math

import pytest
import torch

from oracle import umt5_oracle as uo

pytestmark = pytest.mark.gpu
BF16, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from longcat_video_tta_b200 import ops as o
    o.selfcheck()
    return o


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(BF16)


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm())).item()


@pytest.mark.parametrize("rows,C", [(5, 256), (96, 256), (512, 4096), (130, 1024)])
def test_t5_rmsnorm_bit_exact_against_the_bf16_formula(ops, rows, C):
    x, w = rnd(rows, C, scale=3.0, seed=1), (1 + 0.1 * rnd(C, seed=2).float()).to(BF16)
    y = torch.empty_like(x)
    ops.t5_rmsnorm(y, x, w, 1e-6)
    ref = uo.rms_norm(x.cpu(), w.cpu(), 1e-6).cuda()      # the oracle's (= transformers') bf16 rounding points
    assert ref.dtype == BF16
    # the variance is a 4096-term fp32 sum in a different order: allow the last bf16 bit on a handful of elements
    diff = (y.float() - ref.float()).abs()
    assert (diff <= 2 ** -7 * ref.float().abs() + 1e-30).all()
    assert (diff > 0).float().mean().item() < 0.02


def attn_ref(q, k, v, rel, valid, n_tok, heads, batch):
    q, k, v = (t.float().view(batch, n_tok, heads, 64).transpose(1, 2) for t in (q, k, v))
    s = q @ k.transpose(2, 3)
    idx = torch.arange(n_tok, device=q.device)
    bias = rel[:, (idx[None, :] - idx[:, None]) + n_tok - 1]                # [heads, q, key]
    s = s + bias[None]
    if valid is not None:
        s = s + (1.0 - valid[:, None, None, :].float()) * torch.finfo(torch.float32).min
    p = torch.softmax(s, -1)
    return (p @ v).transpose(1, 2).reshape(batch * n_tok, heads * 64)


@pytest.mark.parametrize("n_tok,heads,batch,masked", [(64, 4, 1, False), (96, 4, 2, True), (200, 3, 1, True),
                                                      (512, 8, 2, True), (33, 2, 3, True)])
def test_t5_attn_against_fp32(ops, n_tok, heads, batch, masked):
    inner = heads * 64
    qkv = rnd(batch * n_tok, 3 * inner, scale=0.35, seed=3)      # strided views, as the encoder passes them
    q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
    rel = torch.randn(heads, 2 * n_tok - 1, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    valid = None
    if masked:
        valid = torch.ones(batch, n_tok, dtype=torch.int32, device="cuda")
        for b in range(batch):
            valid[b, max(1, n_tok * (b + 2) // 5):] = 0
    o = torch.full((batch * n_tok, inner), float("nan"), dtype=BF16, device="cuda")
    ops.t5_attn(o, q, k, v, rel, valid, n_tok, heads, batch)
    ref = attn_ref(q, k, v, rel, valid, n_tok, heads, batch)
    assert torch.isfinite(o.float()).all()
    assert cos(o, ref) > 0.9999
    assert torch.allclose(o.float(), ref, rtol=2e-2, atol=2e-2 * ref.abs().max().item())


def test_t5_attn_fully_masked_item_is_uniform_like_transformers(ops):
    """finfo.min on every key: the scores are absorbed and the softmax is uniform over ALL keys (transformers' behaviour)"""
    n_tok, heads = 64, 2
    qkv = rnd(n_tok, 3 * heads * 64, seed=5)
    inner = heads * 64
    rel = torch.zeros(heads, 2 * n_tok - 1, device="cuda")
    valid = torch.zeros(1, n_tok, dtype=torch.int32, device="cuda")
    o = torch.empty(n_tok, inner, dtype=BF16, device="cuda")
    ops.t5_attn(o, qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], rel, valid, n_tok, heads, 1)
    ref = qkv[:, 2 * inner:].float().mean(0, keepdim=True).expand(n_tok, -1)
    assert torch.allclose(o.float(), ref, atol=2e-2)


def test_geglu_and_residual_epilogues(ops):
    rows, d, ff = 200, 256, 512
    x, w0, w1 = rnd(rows, d, seed=6), rnd(ff, d, scale=d ** -0.5, seed=7), rnd(ff, d, scale=d ** -0.5, seed=8)
    g = torch.empty(rows, ff, dtype=BF16, device="cuda")
    ops.gemm(rows, 2 * ff, [(x, w0, d, False, w1)], ops.epi(ops.EPI_GEGLU, g))
    ref = uo.gelu_new(x.float() @ w0.float().T) * (x.float() @ w1.float().T)
    assert cos(g, ref) > 0.9999 and torch.allclose(g.float(), ref, rtol=2e-2, atol=2e-2 * ref.abs().max().item())
    wo, res = rnd(d, ff, scale=ff ** -0.5, seed=9), rnd(rows, d, seed=10)
    y = torch.empty(rows, d, dtype=BF16, device="cuda")
    ops.gemm(rows, d, [(g, wo, ff, False, None)], ops.epi(ops.EPI_GATE_RESID, y, resid=res))
    ref2 = res.float() + g.float() @ wo.float().T
    assert cos(y, ref2) > 0.9999 and torch.allclose(y.float(), ref2, rtol=2e-2, atol=2e-2 * ref2.abs().max().item())


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(golden_dir / "umt5_tiny.pt")


def build_encoder(cfg, state):
    from longcat_video_tta_b200.text_encoder import B200UMT5Encoder
    kw = {k: v for k, v in cfg.items() if k != "vocab_size"}
    return B200UMT5Encoder(state, device="cuda", **kw)


def test_encoder_against_transformers_golden(ops, golden):
    cfg = golden["cfg"]
    enc = build_encoder(cfg, uo.tiny_state(cfg, golden["state_seed"]))
    n0 = ops.kernel_launches()
    for c in golden["cases"]:
        out = enc(c["input_ids"].cuda(), c["attention_mask"].cuda()).last_hidden_state
        ref32, ref16 = c["last_hidden_state_fp32"].cuda(), c["last_hidden_state_bf16"].cuda()
        assert out.shape == ref32.shape and out.dtype == BF16
        e_ours, e_hf16 = rel_l2(out, ref32), rel_l2(ref16, ref32)
        print(f"n_tok {c['n_tok']}: rel-L2 vs transformers fp32 {e_ours:.2e} (transformers bf16: {e_hf16:.2e}), "
              f"cosine {cos(out, ref32):.6f}")
        assert e_ours <= 1.25 * e_hf16 and e_ours < 3e-2
        assert cos(out, ref32) > 0.999
        keep = c["attention_mask"].bool().cuda()
        assert cos(out[keep], ref32[keep]) > 0.999
    assert ops.kernel_launches() - n0 >= 2 * (2 + 7 * cfg["num_layers"])     # our kernels ran, per call


def test_encoder_without_mask_and_repeatability(ops, golden):
    cfg = golden["cfg"]
    st = uo.tiny_state(cfg, golden["state_seed"])
    enc = build_encoder(cfg, st)
    ids, _ = uo.tiny_inputs(cfg, batch=1, n_tok=128, seed=3)
    ref = uo.umt5_encode(st, cfg, ids, torch.ones_like(ids)).cuda()
    a = enc(ids.cuda()).last_hidden_state
    b = enc(ids.cuda(), torch.ones_like(ids).cuda()).last_hidden_state
    assert cos(a, ref) > 0.999 and rel_l2(a, ref) < 3e-2
    # the CTA-pair GEMM's summation order is not fixed run to run; bf16 activations turn last-bit differences
    # into one-ulp flips that the sharp attention of this seeded model amplifies: 2.2e-3 - 2.6e-3 measured
    assert rel_l2(b, a) < 1e-2
    with pytest.raises(IndexError):
        enc(torch.full((1, 64), cfg["vocab_size"], device="cuda"))
    with pytest.raises(ValueError):
        enc(ids.cuda(), torch.ones(1, 5, device="cuda"))


def test_encode_prompt_feeds_the_dit_format(ops, golden):
    """encode_prompt -> prompt_embeds [1, 1, N, C] bf16 + mask [1, N], what finetune_lora_on_conditioning takes"""
    from types import SimpleNamespace

    from longcat_video_tta_b200.text_encoder import encode_prompt
    cfg = golden["cfg"]
    enc = build_encoder(cfg, uo.tiny_state(cfg, golden["state_seed"]))
    c = golden["cases"][0]

    def tokenizer(texts, max_length, **kw):
        return SimpleNamespace(input_ids=c["input_ids"][:1, :max_length], attention_mask=c["attention_mask"][:1, :max_length])

    emb, mask = encode_prompt(tokenizer, enc, "a prompt", device="cuda", max_length=96)
    assert emb.shape == (1, 1, 96, cfg["d_model"]) and emb.dtype == BF16 and emb.is_cuda
    assert mask.shape == (1, 96) and mask.is_cuda
    ref = c["last_hidden_state_fp32"][:1].cuda()
    assert cos(emb[:, 0], ref) > 0.999


def test_from_hf_wraps_a_live_transformers_model(ops, golden):
    """the drop-in path of INTEGRATION.md: B200UMT5Encoder.from_hf(model) on the object common.py:62-64 loads (bf16, on
    the GPU), same call, compared with that model in fp32"""
    tr = pytest.importorskip("transformers")
    if not hasattr(tr, "UMT5EncoderModel"):
        pytest.skip("this transformers has no UMT5")
    from longcat_video_tta_b200.text_encoder import B200UMT5Encoder
    from oracle.make_golden_umt5 import hf_model
    cfg = golden["cfg"]
    st = uo.tiny_state(cfg, seed=11)
    m32 = hf_model(cfg, st, torch.float32).cuda()
    m16 = hf_model(cfg, st, torch.bfloat16).cuda()
    enc = B200UMT5Encoder.from_hf(m16)
    ids, mask = uo.tiny_inputs(cfg, batch=3, n_tok=72, seed=12)
    ids, mask = ids.cuda(), mask.cuda()
    with torch.no_grad():
        ref, lib = m32(ids, mask).last_hidden_state, m16(ids, mask).last_hidden_state
    out = enc(ids, mask).last_hidden_state
    assert out.shape == ref.shape
    assert rel_l2(out, ref) <= 1.5 * rel_l2(lib, ref) and cos(out, ref) > 0.999     # (the golden test holds the 1.25 x bar)
