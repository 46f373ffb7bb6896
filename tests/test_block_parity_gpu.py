"""Per-block parity of the engine against an fp32 oracle fed the ENGINE'S OWN block inputs (north_star: adapter gradients
cosine > 0.999 against an fp32 reference, rtol 2e-2).

The whole-model comparison against an fp32 run (tests/test_step_gpu.py) mixes two things: the error of one block's
kernels, and the drift of a bf16 activation stream through depth, which on a random-init model with near-uniform
attention is amplified by cancellation (sum_i g_i (o_i A^T) with o_i ~ mean(V): the fp32 and the bf16 streams differ
in the small token-dependent part of o).  These tests separate them: for every block the fp32 oracle block
(bf16-rounded weights, fp32 arithmetic, TF32 off) gets the block input x_b, the text rows y, the time embedding t and
the upstream gradient dL/dx_{b+1} exactly as the engine held them (bf16 values), and has to agree with what the engine
produced for that block: x_{b+1}, dL/dx_b and every adapter gradient of the block.

Cases: the tiny config (BASELINE.json configs[0]) with all eight linears adapted, stash on and off; the same with sharp
(low-entropy) attention; and ONE block of the 13.6 B architecture (hidden 4096, 32 heads, FFN 11008) on 4 680 and on
the headline 37 440 tokens (the latter against the fp32 oracle with a chunked exact attention).
"""
import pytest
import torch

from parity_util import (BF16, F32, COS_BAR, NORM_RTOL, cos, rel, expect, build_pair, run_engine, tiny_case, wide_inputs,
                         chunked_exact_sdpa)

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


def check_blocks(oracle, omods, dit, dbg, label, sites_per_block, cos_bar=COS_BAR):
    """Feed every oracle block the engine's own inputs / upstream gradient; returns the worst cosine seen."""
    from oracle import tta_oracle as T
    eng = dit.engine
    ws, geo = eng.ws, eng.geo
    y = ws.y.float()[None]
    t = ws.t.float()[None]
    shape = (geo.T, geo.gh, geo.gw)
    worst = 1.0
    sites = eng.lora_sites()
    for b, blk in enumerate(oracle.blocks):
        x = ws.xs[b].float()[None].clone().requires_grad_(True)
        g_out = dbg[(b, "dx_out")][None]
        params = T.lora_parameters(omods[b * sites_per_block:(b + 1) * sites_per_block])
        for p in params:
            p.requires_grad_(True)
        out = blk(x, y, t, [geo.M], shape, num_cond_latents=geo.n_cond)
        grads = torch.autograd.grad((out * g_out).sum(), [x] + params)
        # forward of the block
        c_f, r_f = cos(ws.xs[b + 1], out[0]), rel(ws.xs[b + 1], out[0])
        # dL/dx_b
        c_x, r_x = cos(dbg[(b, "dx_in")], grads[0][0]), rel(dbg[(b, "dx_in")], grads[0][0])
        print(f"[{label}] block {b}: x_out cosine {c_f:.6f} rel {r_f:.3g} | dx_in cosine {c_x:.6f} rel {r_x:.3g}")
        expect(c_f > cos_bar and r_f < NORM_RTOL, f"block {b}: forward cosine {c_f} rel {r_f}")
        expect(c_x > cos_bar, f"block {b}: dX cosine {c_x}")
        worst = min(worst, c_f, c_x)
        mine = [gr for s in sites[b * sites_per_block:(b + 1) * sites_per_block] for gr in s.param_grads()]
        assert len(mine) == len(params)
        for i, (gm, go) in enumerate(zip(mine, grads[1:])):
            s = sites[b * sites_per_block + i // 2]
            c = cos(gm, go)
            nr = (gm.double().norm() / go.double().norm()).item()
            print(f"[{label}]   {s.name:32s} {'dA' if i % 2 == 0 else 'dB'}: cosine {c:.6f} norm ratio {nr:.4f}")
            expect(c > cos_bar, f"{s.name} {'dA' if i % 2 == 0 else 'dB'}: cosine {c} vs the fp32 oracle block")
            expect(abs(nr - 1) < NORM_RTOL, f"{s.name}: gradient norm ratio {nr}")
            worst = min(worst, c)
    return worst


@pytest.mark.parametrize("stash_gb", [None, 0])
@pytest.mark.parametrize("sharpen", [1.0, 2.0])
@pytest.mark.parametrize("target_ffn", [True, False])
def test_tiny_blocks_match_fp32_oracle_block_by_block(stash_gb, sharpen, target_ffn):
    """target_ffn=False runs the w1|w3 co-tiled SwiGLU-epilogue GEMM (the headline path), True the separate ones."""
    oracle, omods, dit, _ = build_pair("tiny", 0, sharpen=sharpen, target_ffn=target_ffn)
    loss, dbg = run_engine(dit, *tiny_case(), stash_gb=stash_gb)
    worst = check_blocks(oracle, omods, dit, dbg, f"tiny sharpen={sharpen} ffn={target_ffn} stash={stash_gb}",
                         8 if target_ffn else 5)
    print(f"tiny per-block worst cosine {worst:.6f}")


@pytest.mark.parametrize("sharpen", [1.0, 2.0])
def test_tiny_whole_model_vs_fp32(sharpen):
    """Both blocks end to end against the fp32 oracle.  With the default random init attention is near uniform and the
    bf16 activation stream costs gradient direction on the self-attention adapters (reported, bar 0.97 as bf16 PyTorch
    itself lands there: test_step_gpu.py); with realistic (sharper) attention the north-star bar must hold."""
    from oracle import tta_oracle as T
    oracle, omods, dit, _ = build_pair("tiny", 0, sharpen=sharpen, target_ffn=False)
    inputs = tiny_case()
    loss, _ = run_engine(dit, *inputs, taps=False)
    params = T.lora_parameters(omods)
    for p in params:
        p.requires_grad_(True)
    cond, train, prompt, mask, sigma, eps = inputs
    oloss = T.fm_loss_given(oracle, cond, train, prompt, mask, sigma, eps, BF16)
    ograds = torch.autograd.grad(oloss, params)
    expect(abs(loss - oloss.item()) <= 2e-2 * abs(oloss.item()), "loss")
    mine = [gr for s in dit.engine.lora_sites() for gr in s.param_grads()]
    cs = [cos(a, b) for a, b in zip(mine, ograds)]
    print(f"[tiny whole model sharpen={sharpen}] loss {loss:.6f} vs {oloss.item():.6f}; adapter-gradient cosines: "
          + " ".join(f"{c:.5f}" for c in cs))
    allc = cos(torch.cat([m.flatten() for m in mine]), torch.cat([o.flatten() for o in ograds]))
    print(f"[tiny whole model sharpen={sharpen}] all adapters concatenated: {allc:.6f}")
    expect(min(cs) > (COS_BAR if sharpen > 1.0 else 0.97), f"worst whole-model cosine {min(cs)}")


def _whole_model_check(oracle, omods, dit, inputs, loss_mine, label):
    """depth-1 model: the engine's loss and adapter gradients against the fp32 oracle run end to end."""
    from oracle import tta_oracle as T
    cond, train, prompt, mask, sigma, eps = inputs
    params = T.lora_parameters(omods)
    for p in params:
        p.requires_grad_(True)
    # dtype=BF16: hidden states / timestep are quantised exactly as the reference does on a GPU (common.py:466-470);
    # the fp32 oracle then up-casts them, so both sides start from identical numbers
    oloss = T.fm_loss_given(oracle, cond, train, prompt, mask, sigma, eps, BF16)
    ograds = torch.autograd.grad(oloss, params)
    print(f"[{label}] loss mine {loss_mine:.6f} fp32 oracle {oloss.item():.6f}")
    expect(abs(loss_mine - oloss.item()) <= 2e-2 * abs(oloss.item()), "loss")
    mine = [gr for s in dit.engine.lora_sites() for gr in s.param_grads()]
    worst = 1.0
    for i, (gm, go) in enumerate(zip(mine, ograds)):
        c = cos(gm, go)
        nr = (gm.double().norm() / go.double().norm()).item()
        print(f"[{label}]   whole-model adapter grad {i:2d}: cosine {c:.6f} norm ratio {nr:.4f}")
        expect(c > COS_BAR and abs(nr - 1) < NORM_RTOL, f"whole-model adapter grad {i}: cosine {c} norm ratio {nr}")
        worst = min(worst, c)
    return worst


@pytest.mark.parametrize("stash_gb,target_ffn", [(None, True), (0, False), (None, False)])
def test_one_headline_width_block_4680_tokens(stash_gb, target_ffn):
    """hidden 4096, 32 heads x 128, FFN 11008, one block, [1 context | 2 noised] latent frames of 60 x 104."""
    oracle, omods, dit, _ = build_pair("13.6b", 3, init_std=0.02, depth=1, target_ffn=target_ffn)
    inputs = wide_inputs(1, 2, 60, 104, 4096)
    loss, dbg = run_engine(dit, *inputs, stash_gb=stash_gb)
    worst = check_blocks(oracle, omods, dit, dbg, f"4096-wide ffn={target_ffn} stash={stash_gb}", 8 if target_ffn else 5)
    worst = min(worst, _whole_model_check(oracle, omods, dit, inputs, loss, f"4096-wide stash={stash_gb}"))
    print(f"headline-width block, 4 680 tokens: worst cosine {worst:.6f}")


@pytest.mark.parametrize("stash_gb", [None, 0])
def test_one_headline_width_block_headline_tokens(stash_gb, monkeypatch):
    """The assembly at the BASELINE.json configs[1] geometry: one 13.6 B-architecture block on 37 440 tokens
    ([4 context | 20 noised] frames of 60 x 104, 512 text tokens), loss and every adapter gradient against the fp32
    oracle (exact attention, query-chunked), stash on and off."""
    if torch.cuda.mem_get_info()[1] < 150 << 30:
        pytest.skip("needs a 180 GB device")
    import oracle.dit_oracle as D
    monkeypatch.setattr(D, "_sdpa", chunked_exact_sdpa)
    oracle, omods, dit, _ = build_pair("13.6b", 4, init_std=0.02, depth=1, target_ffn=False)
    inputs = wide_inputs(4, 20, 60, 104, 4096)
    loss, _ = run_engine(dit, *inputs, stash_gb=stash_gb, taps=False)
    assert dit.engine.geo.N == 37440
    worst = _whole_model_check(oracle, omods, dit, inputs, loss, f"37 440 tokens stash={stash_gb}")
    print(f"headline-width block, 37 440 tokens: worst cosine {worst:.6f}")
