"""CPU: the oracle port (oracle/tta_oracle.py + oracle/dit_oracle.py) reproduces the
fixtures that the reference's own unmodified code produced (oracle/make_golden.py)."""
import torch
import pytest

from oracle.dit_oracle import build_oracle_dit
from oracle import tta_oracle as T
from oracle.make_golden import tiny_inputs, tiny_split, LOOP_SEED


@pytest.fixture(scope="module")
def tiny():
    torch.set_num_threads(min(8, torch.get_num_threads()))
    latents, prompt, mask = tiny_inputs()
    cond, train, val = tiny_split(latents)
    return cond, train, val, prompt, mask


def test_split_table(golden_dir):
    table = torch.load(golden_dir / "split_table.pt")
    for (Tn, ctx, hf), want in table.items():
        c, t, v = T.split_tta_latents(torch.zeros(1, 1, Tn, 1, 1), ctx, hf)
        assert (c.shape[2], t.shape[2], 0 if v is None else v.shape[2]) == want
    # the two cases SURVEY 3.1 probes
    assert table[(4, 4, 0.25)] == (3, 1, 0)
    assert table[(30, 4, 0.25)] == (4, 20, 6)


def test_lora_loop_matches_reference_golden(golden_dir, tiny):
    cond, train, val, prompt, mask = tiny
    g = torch.load(golden_dir / "lora_tiny.pt")
    dit = build_oracle_dit("tiny", seed=0)
    torch.manual_seed(g["config"]["adapter_seed"])
    mods = T.inject_lora(dit, rank=16, alpha=32.0, target_modules=("qkv", "proj"))
    rec = {}

    def on_step(step, **kw):
        if step == 0:
            rec.update(kw)

    torch.manual_seed(LOOP_SEED)
    out = T.lora_tta_loop(dit, mods, cond, train, prompt, mask, num_steps=5, lr=2e-4, warmup_steps=3,
                          weight_decay=0.01, max_grad_norm=1.0, on_step=on_step)
    assert out["losses"] == pytest.approx(g["losses"], rel=1e-6)
    assert torch.allclose(rec["pred"], g["pred_step0"], rtol=1e-5, atol=1e-6)
    coef = torch.clamp(1.0 / (rec["total_norm"] + 1e-6), max=1.0)
    for mine, want in zip(rec["grads"], g["clipped_grads_step0"]):
        assert torch.allclose(mine * coef, want, rtol=1e-4, atol=1e-7)
    for p, want in zip(T.lora_parameters(mods), g["params_after_5"]):
        assert torch.allclose(p, want, rtol=1e-4, atol=1e-6)
    # KAT: B == 0 at init  =>  dA == 0 and dB != 0 at step 0
    for i, gr in enumerate(rec["grads"]):
        if i % 2 == 0:
            assert gr.abs().max() == 0
        else:
            assert gr.abs().max() > 0


def test_delta_a_matches_reference_golden(golden_dir, tiny):
    cond, train, val, prompt, mask = tiny
    g = torch.load(golden_dir / "delta_tiny.pt")["delta_a"]
    w = T.DeltaA(build_oracle_dit("tiny", seed=0), 512)
    torch.manual_seed(LOOP_SEED)
    out = T.delta_loop(w, [w.delta], cond, train, prompt, mask, num_steps=3, lr=1e-3)
    assert out["losses"] == pytest.approx(g["losses"], rel=1e-6)
    assert torch.allclose(w.delta.detach(), g["params"][0], rtol=1e-4, atol=1e-7)


def test_anchor_loss_matches_reference_golden(golden_dir, tiny):
    import hashlib
    cond, train, val, prompt, mask = tiny
    g = torch.load(golden_dir / "anchor_tiny.pt")
    base = int(hashlib.md5(g["video_id"].encode()).hexdigest()[:8], 16) % (2 ** 31)  # early_stopping.py:166
    noises = []
    for d in range(2):
        gen = torch.Generator().manual_seed(base + d)
        noises.append(torch.randn(val.shape, generator=gen))
    dit = build_oracle_dit("tiny", seed=0)
    loss = T.fm_loss_conditioned_fixed(dit, cond, val, prompt, mask, [0.25, 0.5, 0.75], noises)
    assert loss == pytest.approx(g["anchor_loss0"], rel=1e-6)


def test_cross_attn_cond_rows_zero_and_timestep_zero(tiny):
    """KATs from SURVEY 8c-ii that hold regardless of upstream details."""
    cond, train, val, prompt, mask = tiny
    dit = build_oracle_dit("tiny", seed=0)
    blk = dit.blocks[0]
    x = torch.randn(1, 1024, 512)
    y, seqlens = dit.embed_text(prompt, mask)
    out = blk.cross_attn(x, y, seqlens, num_cond_latents=2, shape=(4, 16, 16))
    assert out[:, :512].abs().max() == 0 and out[:, 512:].abs().max() > 0
    sigma = torch.tensor([0.5])
    hidden, timestep, n_cond = T.build_step_inputs(cond, train, sigma, torch.randn_like(train), torch.float32)
    assert n_cond == 2 and timestep[0, :2].abs().max() == 0 and (timestep[0, 2:] == 500).all()
    assert torch.equal(hidden[:, :, :2], cond)
