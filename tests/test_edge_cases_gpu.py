"""Edge geometries of the TTA step on the GPU against the fp32 oracle (same bf16-valued weights; north-star bars: loss
rtol 2e-2, every adapter gradient cosine > 0.999): ragged and extreme text masks (one real token, scattered tokens, all
512), sequences shorter than one 128-row tile, token counts and frame boundaries that do not fall on tile boundaries,
single-frame context / target.  The reference's tokenizer always leaves at least the EOS token, so an empty prompt is
not a case of the path (common.py:236-246)."""
import pytest
import torch

from parity_util import BF16, COS_BAR, NORM_RTOL, build_pair, cos, expect, run_engine

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _true_fp32():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def make_inputs(T_c, T_t, Hl, Wl, Cc, mask_kind, seed=5):
    g = torch.Generator(device="cuda").manual_seed(seed)
    cond = torch.randn(1, 16, T_c, Hl, Wl, generator=g, device="cuda").to(BF16)
    train = torch.randn(1, 16, T_t, Hl, Wl, generator=g, device="cuda").to(BF16)
    prompt = torch.randn(1, 1, 512, Cc, generator=g, device="cuda").to(BF16)
    mask = torch.zeros(1, 512, dtype=torch.int64, device="cuda")
    if mask_kind == "one":
        mask[:, 0] = 1
    elif mask_kind == "scattered":          # masked_select keeps the real tokens in order (run_delta_a.py:170-192)
        mask[:, ::3] = 1
        mask[:, 200:260] = 0
        mask[:, 511] = 1
    elif mask_kind == "all":
        mask[:] = 1
    else:
        mask[:, :int(mask_kind)] = 1
    sigma = torch.tensor([0.42], device="cuda")
    eps = torch.randn(1, 16, T_t, Hl, Wl, generator=g, device="cuda").to(BF16)
    return cond, train, prompt, mask, sigma, eps


CASES = {
    # name: (T_cond, T_train, latent H, latent W, text mask)
    "one_text_token": (2, 2, 32, 32, "one"),
    "scattered_text_mask": (2, 2, 32, 32, "scattered"),
    "all_512_text_tokens": (2, 2, 32, 32, "all"),
    "shorter_than_one_tile": (1, 1, 12, 20, "7"),          # 2 x 60 = 120 tokens
    "ragged_frames_and_tiles": (3, 2, 20, 28, "129"),      # 5 x 140 = 700 tokens, context ends at 420
    "single_target_frame": (4, 1, 16, 24, "64"),           # 5 x 96 = 480 tokens
}


@pytest.mark.parametrize("name", list(CASES))
def test_edge_geometry_against_fp32_oracle(name):
    from oracle import tta_oracle as T
    T_c, T_t, Hl, Wl, mk = CASES[name]
    oracle, omods, dit, _ = build_pair("tiny", 0, sharpen=2.0, target_ffn=False)
    inputs = make_inputs(T_c, T_t, Hl, Wl, dit.config.caption_channels, mk)
    loss, _ = run_engine(dit, *inputs, taps=False)
    params = T.lora_parameters(omods)
    for p in params:
        p.requires_grad_(True)
    cond, train, prompt, mask, sigma, eps = inputs
    oloss = T.fm_loss_given(oracle, cond, train, prompt, mask, sigma, eps, BF16)
    ograds = torch.autograd.grad(oloss, params)
    expect(abs(loss - oloss.item()) <= NORM_RTOL * abs(oloss.item()), f"{name}: loss {loss} vs {oloss.item()}")
    mine = [gr for s in dit.engine.lora_sites() for gr in s.param_grads()]
    assert len(mine) == len(ograds)
    cs = [cos(a, b) for a, b in zip(mine, ograds)]
    print(f"[{name}] loss {loss:.6f} vs {oloss.item():.6f}; worst adapter-gradient cosine {min(cs):.6f}")
    for i, (gm, go) in enumerate(zip(mine, ograds)):
        assert torch.isfinite(gm).all()
        expect(cs[i] > COS_BAR, f"{name}: adapter gradient {i} cosine {cs[i]}")
