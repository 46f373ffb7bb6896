"""GPU end-to-end parity of the TTA step (BASELINE.json configs[0]: tiny DiT, LoRA r=16) through the public API.

Two oracles:
  * the committed golden vectors produced by the reference's own unmodified code (fp32 weights, CPU);
  * the in-repo oracle re-run here with bf16-rounded weights/inputs, which isolates kernel error from the bf16
    storage of the frozen backbone.
Tolerances are the ones BASELINE.json's north_star states: rtol 2e-2, cosine similarity > 0.999.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

BF16, F32 = torch.bfloat16, torch.float32


def cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle.dit_oracle import build_oracle_dit
    from oracle.make_golden import tiny_inputs, tiny_split
    from longcat_video_tta_b200.dit import B200DiT
    latents, prompt, mask = tiny_inputs()
    cond, train, val = tiny_split(latents)
    oracle = build_oracle_dit("tiny", seed=0)
    return dict(oracle=oracle, cond=cond, train=train, val=val, prompt=prompt, mask=mask, B200DiT=B200DiT)


def replay_draws(train, n_steps, seed=42):
    """The reference loop's RNG order on the CPU generator (run_lora_tta.py:499, common.py:458,462)."""
    torch.manual_seed(seed)
    out = []
    for _ in range(n_steps):
        torch.randint(0, 1, (1,))
        sigma = torch.rand(1, dtype=F32) * (1.0 - 0.001) + 0.001
        out.append((sigma, torch.randn_like(train)))
    return out


def copy_to_bf16(oracle):
    import copy
    return copy.deepcopy(oracle).to(BF16).cuda()


def rounded_oracle(oracle):
    import copy
    o = copy.deepcopy(oracle)
    with torch.no_grad():
        for p in o.parameters():
            p.copy_(p.to(BF16).float())
    return o.cuda()


def test_forward_matches_oracle(setup):
    from oracle import tta_oracle as T
    s = setup
    dit = s["B200DiT"].from_oracle(s["oracle"])
    ro = rounded_oracle(s["oracle"])
    (sigma, eps), = replay_draws(s["train"], 1)
    cond, train, prompt = (s[k].to(BF16).cuda() for k in ("cond", "train", "prompt"))
    mask = s["mask"].cuda()
    hidden, timestep, n_cond = T.build_step_inputs(cond, train, sigma.cuda(), eps.to(BF16).cuda(), BF16)
    with torch.no_grad():
        got = dit(hidden, timestep, prompt, mask, num_cond_latents=n_cond)
        want = ro(hidden.float(), timestep, prompt.float(), mask, num_cond_latents=n_cond)
        # what the reference itself runs on a GPU: the same network with bf16 parameters and activations
        bf = copy_to_bf16(s["oracle"])(hidden, timestep, prompt, mask, num_cond_latents=n_cond)
    assert got.shape == want.shape == (1, 16, 4, 32, 32) and got.dtype == F32
    c = cos(got, want)
    rel = ((got - want).norm() / want.norm()).item()
    rel_bf = ((bf - want).norm() / want.norm()).item()
    err = (got - want).abs().max().item()
    print(f"forward: cosine {c:.6f} rel-L2 {rel:.4g} (bf16 PyTorch oracle: {rel_bf:.4g}) max|err| {err:.4g} "
          f"(ref max {want.abs().max().item():.4g})")
    assert c > 0.999
    # adapted-denoise latent tolerance: relative L2 error <= 1e-2 and no worse than 1.5x what plain bf16 PyTorch
    # (the reference's own arithmetic) loses against fp32 on the same inputs
    assert rel < 1e-2
    assert rel <= 1.5 * rel_bf + 1e-3


def bf16_torch_grads(oracle, adapter_seed, cond, train, prompt, mask, sigma, eps):
    """The same step in plain bf16 PyTorch (what the reference executes on a GPU): oracle modules cast to bf16,
    autograd backward.  Returns (loss, grads in injection order)."""
    import copy
    from oracle import tta_oracle as T
    ob = copy.deepcopy(oracle)
    torch.manual_seed(adapter_seed)
    mods = T.inject_lora(ob, rank=16, alpha=32.0, target_modules=("qkv", "proj"))
    ob = ob.to(BF16).cuda()
    params = T.lora_parameters(mods)
    for p in params:
        p.requires_grad_(True)
    loss = T.fm_loss_given(ob, cond, train, prompt, mask, sigma, eps, BF16)
    return loss.item(), torch.autograd.grad(loss, params)


def test_autograd_path_grads_match_golden_and_oracle(setup, golden_dir):
    """dit(...) + the reference-style loss + loss.backward(): the drop-in seam (common.py:476-488).

    Four-way comparison of the step-0 adapter gradients:
      fp32    the oracle in fp32 arithmetic on the bf16-VALUED weights / inputs the GPU path stores   <- north-star bar
      golden  fp32 CPU with UN-ROUNDED fp32 weights, produced by the reference's own code (tests/golden/lora_tiny.pt)
      bf16    the identical step in plain bf16 PyTorch on this GPU      (the arithmetic the reference really runs)
      mine    the sm_100a kernels
    North-star tolerance (cosine > 0.999, norm within 2e-2) is asserted against ``fp32`` and against ``bf16`` for every
    tensor.  The golden run differs from all three others by the storage format of the frozen backbone (its weights
    were never rounded to bf16; the reference's GPU backbone is bf16): on this tiny random-init model that alone moves
    the self-attention adapter gradients to cosine ~0.98, for bf16 PyTorch exactly as for the kernels (round 1 mistook
    this for kernel error; tests/test_block_parity_gpu.py separates the two).  Against the golden the bound is therefore
    "at least as close as bf16 PyTorch is".
    """
    from oracle import tta_oracle as T
    from longcat_video_tta_b200 import lora
    s = setup
    g = torch.load(golden_dir / "lora_tiny.pt")
    dit = s["B200DiT"].from_oracle(s["oracle"])
    torch.manual_seed(g["config"]["adapter_seed"])
    mods = lora.inject_lora_into_dit(dit, rank=16, alpha=32.0, target_modules=["qkv", "proj"])
    params = lora.get_lora_parameters(mods)
    assert len(params) == 20 and all(p.is_cuda and p.dtype == BF16 for p in params)
    (sigma, eps), = replay_draws(s["train"], 1)
    cond, train, prompt = (s[k].to(BF16).cuda() for k in ("cond", "train", "prompt"))
    mask = s["mask"].cuda()
    sigma, eps = sigma.cuda(), eps.to(BF16).cuda()
    for p in params:
        p.requires_grad_(True)
    loss = T.fm_loss_given(dit, cond, train, prompt, mask, sigma, eps, BF16)
    loss.backward()
    bloss, bgrads = bf16_torch_grads(s["oracle"], g["config"]["adapter_seed"], cond, train, prompt, mask, sigma, eps)
    # fp32 arithmetic on the same bf16-valued weights and adapter factors (TF32 off)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        ro = rounded_oracle(s["oracle"])
        torch.manual_seed(g["config"]["adapter_seed"])
        romods = T.inject_lora(ro, rank=16, alpha=32.0, target_modules=("qkv", "proj"))
        with torch.no_grad():
            for m, om in zip(mods, romods):
                om.lora_down.weight.copy_(m.lora_down.weight.float())
        rparams = T.lora_parameters(romods)
        for p in rparams:
            p.requires_grad_(True)
        rloss = T.fm_loss_given(ro, cond, train, prompt, mask, sigma, eps, BF16)
        rgrads = torch.autograd.grad(rloss, rparams)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    assert abs(loss.item() - rloss.item()) <= 2e-2 * rloss.item()
    print(f"loss mine {loss.item():.6f} bf16-torch {bloss:.6f} golden(fp32 reference) {g['losses'][0]:.6f}")
    assert abs(loss.item() - g["losses"][0]) <= 2e-2 * g["losses"][0]
    # golden holds CLIPPED grads of step 0; clipping is a positive scalar -> compare directions + norm ratio
    gold = g["clipped_grads_step0"]
    total = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in params)).item()
    coef = min(1.0, 1.0 / (total + 1e-6))
    mine_all, gold_all = [], []
    for i, (p, want, bg, rg) in enumerate(zip(params, gold, bgrads, rgrads)):
        got = p.grad.float().cpu() * coef
        if i % 2 == 0:   # KAT: B == 0 at init => dA == 0 exactly
            assert got.abs().max() == 0 and want.abs().max() == 0
            continue
        c_gold, c_bf, c_bf_gold, c_fp32 = cos(got, want), cos(p.grad, bg), cos(bg.cpu(), want), cos(p.grad, rg)
        print(f"param {i:2d}: mine~fp32(same weights) {c_fp32:.6f}  mine~bf16-torch {c_bf:.6f}  "
              f"mine~golden {c_gold:.5f}  bf16-torch~golden {c_bf_gold:.5f}")
        assert c_fp32 > 0.999, f"param {i}: cosine vs the fp32 oracle on the same weights {c_fp32}"
        assert abs(p.grad.float().norm() / rg.float().norm() - 1) < 2e-2
        assert c_bf > 0.999, f"param {i}: cosine vs bf16 PyTorch {c_bf}"
        assert abs(p.grad.float().norm() / bg.float().norm() - 1) < 2e-2
        assert c_gold >= min(0.999, c_bf_gold - 5e-3), f"param {i}: {c_gold} vs bf16 PyTorch's {c_bf_gold}"
        mine_all.append(got.flatten())
        gold_all.append(want.flatten())
    c_all = cos(torch.cat(mine_all), torch.cat(gold_all))
    print(f"all adapter gradients concatenated: cosine vs golden {c_all:.5f}")
    assert c_all > 0.97


def test_five_step_loop_matches_reference_golden(setup, golden_dir):
    """finetune_lora_on_conditioning with the reference's draws, lr schedule, clip and AdamW."""
    from longcat_video_tta_b200 import lora
    from longcat_video_tta_b200.stepper import TTAStepper
    s = setup
    g = torch.load(golden_dir / "lora_tiny.pt")
    dit = s["B200DiT"].from_oracle(s["oracle"])
    torch.manual_seed(g["config"]["adapter_seed"])
    mods = lora.inject_lora_into_dit(dit, rank=16, alpha=32.0, target_modules=["qkv", "proj"])
    params = lora.get_lora_parameters(mods)
    init = [p.detach().float().cpu().clone() for p in params]
    cond, train, prompt = (s[k].to(BF16).cuda() for k in ("cond", "train", "prompt"))
    mask = s["mask"].cuda()
    stepper = TTAStepper(dit, eps=1e-8, weight_decay=0.01, max_grad_norm=1.0, master_weights=True)
    losses = []
    for step, (sigma, eps) in enumerate(replay_draws(s["train"], 5)):
        lr = lora._warmup_lr(2e-4, step, 3)
        losses.append(stepper.step(cond, train, prompt, mask, sigma.cuda(), eps.to(BF16).cuda(), lr).item())
    print("losses", losses, "golden", g["losses"])
    for got, want in zip(losses, g["losses"]):
        assert abs(got - want) <= 2e-2 * want
    # post-step adapter parameters (fp32 masters) vs the reference's fp32 run
    masters = [e["master"] for e in stepper.group.entries]
    for i, (m, want, p0) in enumerate(zip(masters, g["params_after_5"], init)):
        got = m.float().cpu()
        d_got, d_want = got - p0, want - p0
        if d_want.abs().max() == 0:
            continue
        c = cos(d_got, d_want)
        # AdamW's m/sqrt(v) turns gradient-direction noise into sign flips on near-zero entries: the cross-attention
        # adapters (well conditioned) must track the fp32 reference closely, the self-attention ones loosely
        # (see test_autograd_path_grads_match_golden_and_oracle for why bf16 storage bounds them)
        site = (i // 2) % 5
        print(f"param {i:2d}: update cosine vs golden {c:.4f}")
        assert c > (0.97 if site >= 2 else 0.75), f"param {i}: update cosine {c}"
        assert (got - want).abs().max() <= 1.2e-3, f"param {i}: {(got - want).abs().max()}"


def test_drop_in_loop_runs_and_restores_like_reference(setup):
    """The public drop-in entry point with its own RNG draws on the device (signature of run_lora_tta.py:425-442)."""
    from longcat_video_tta_b200 import lora
    s = setup
    dit = s["B200DiT"].from_oracle(s["oracle"])
    torch.manual_seed(7)
    mods = lora.inject_lora_into_dit(dit, rank=16, alpha=32.0)
    cond, train, prompt = (s[k].to(BF16).cuda() for k in ("cond", "train", "prompt"))
    torch.manual_seed(42)
    out = lora.finetune_lora_on_conditioning(dit, mods, cond, train, prompt, s["mask"].cuda(), num_steps=4, lr=2e-4,
                                             warmup_steps=3, device="cuda", dtype=BF16)
    assert set(out) == {"losses", "train_time", "es_check_time", "early_stopping_info"}
    assert len(out["losses"]) == 4 and all(1.0 < v < 4.0 for v in out["losses"])
    ups = [m.lora_up.weight for m in mods]
    assert all(u.abs().max() > 0 for u in ups)  # B moved away from its zero init
    lora.reset_lora_weights(mods)
    assert all(m.lora_up.weight.abs().max() == 0 for m in mods)


def test_builtin_lora_two_steps_match_reference_golden(setup, golden_dir):
    """inject_builtin_lora_into_dit (upstream LoRAModule style: 3 independent (A_i, B_i) on the fused qkv, 2 on
    kv_linear; run_lora_tta.py:104-170) -> block-diagonal up-projection inside one fused rank-48 / rank-32 segment."""
    from longcat_video_tta_b200 import lora
    from longcat_video_tta_b200.stepper import TTAStepper
    s = setup
    g = torch.load(golden_dir / "lora_builtin_tiny.pt")
    dit = s["B200DiT"].from_oracle(s["oracle"])
    torch.manual_seed(7)
    mods = lora.inject_builtin_lora_into_dit(dit, rank=16, alpha=32.0, target_modules=["qkv", "proj"])
    params = lora.get_builtin_lora_parameters(mods)
    assert [tuple(p.shape) for p in params[:4]] == [(48, 512), (512, 16), (512, 16), (512, 16)]
    init = [p.detach().float().cpu().clone() for p in params]
    cond, train, prompt = (s[k].to(BF16).cuda() for k in ("cond", "train", "prompt"))
    stepper = TTAStepper(dit, eps=1e-8, weight_decay=0.01, max_grad_norm=1.0)
    losses = [stepper.step(cond, train, prompt, s["mask"].cuda(), sg.cuda(), e.to(BF16).cuda(), lora._warmup_lr(2e-4, i, 3)).item()
              for i, (sg, e) in enumerate(replay_draws(s["train"], 2))]
    print("builtin losses", losses, "golden", g["losses"])
    for a, b in zip(losses, g["losses"]):
        assert abs(a - b) <= 2e-2 * b
    # engine parameter order == lora.parameters() order (down, then up blocks) for every module
    eng_params = [p for st in dit.engine.lora_sites() for p in st.params]
    assert [id(p) for p in eng_params] == [id(p) for p in params]
    masters = [e["master"].float().cpu() for e in stepper.group.entries]
    dg = torch.cat([(m - i).flatten() for m, i in zip(masters, init)])
    dw = torch.cat([(w - i).flatten() for w, i in zip(g["params_after_2"], init)])
    c = cos(dg, dw)
    print(f"builtin update cosine vs golden {c:.4f}")
    assert c > 0.95


def test_lora_rank4_with_ffn_targets_matches_bf16_torch(setup):
    """rank 4 (padded to 8 inside the kernels) on qkv, proj AND the FFN (w1/w2/w3: unfused SwiGLU path)."""
    import copy
    from oracle import tta_oracle as T
    from longcat_video_tta_b200 import lora
    from longcat_video_tta_b200.stepper import TTAStepper
    s = setup
    dit = s["B200DiT"].from_oracle(s["oracle"])
    torch.manual_seed(11)
    mods = lora.inject_lora_into_dit(dit, rank=4, alpha=8.0, target_modules=["qkv", "proj"], target_ffn=True)
    ob = copy.deepcopy(s["oracle"])
    torch.manual_seed(11)
    omods = T.inject_lora(ob, rank=4, alpha=8.0, target_ffn=True)
    ob = ob.to(BF16).cuda()
    gen = torch.Generator().manual_seed(3)
    with torch.no_grad():   # non-zero B so that dA carries signal too
        for m, om in zip(mods, omods):
            b = (torch.randn(m.lora_up.weight.shape, generator=gen) * 0.02).to(BF16).cuda()
            m.lora_up.weight.copy_(b)
            om.lora_up.weight.copy_(b)
            om.lora_down.weight.copy_(m.lora_down.weight)
    (sigma, eps), = replay_draws(s["train"], 1)
    cond, train, prompt = (s[k].to(BF16).cuda() for k in ("cond", "train", "prompt"))
    mask, sigma, eps = s["mask"].cuda(), sigma.cuda(), eps.to(BF16).cuda()
    oparams = T.lora_parameters(omods)
    for p in oparams:
        p.requires_grad_(True)
    oloss = T.fm_loss_given(ob, cond, train, prompt, mask, sigma, eps, BF16)
    ograds = torch.autograd.grad(oloss, oparams)
    st = TTAStepper(dit)
    loss = st.forward_backward(cond, train, prompt, mask, sigma, eps).item()
    assert abs(loss - oloss.item()) <= 2e-2 * oloss.item()
    mine = [g for site in dit.engine.lora_sites() for g in site.param_grads()]
    assert len(mine) == len(ograds) == 2 * 8 * 2
    worst = min(cos(a, b) for a, b in zip(mine, ograds))
    print(f"rank-4 + FFN LoRA: loss {loss:.5f} vs bf16 torch {oloss.item():.5f}; worst per-tensor gradient cosine {worst:.5f}")
    assert worst > 0.999


def test_unconditioned_loss_variant(setup):
    """compute_flow_matching_loss (common.py:274-343): every frame noised, no context segment (num_cond_latents = 0)."""
    from oracle import tta_oracle as T
    s = setup
    dit = s["B200DiT"].from_oracle(s["oracle"])
    ro = rounded_oracle(s["oracle"])
    lat = torch.cat([s["cond"], s["train"]], dim=2).to(BF16).cuda()
    prompt, mask = s["prompt"].to(BF16).cuda(), s["mask"].cuda()
    t = torch.full((1, 4), 421.0, device="cuda", dtype=BF16)
    with torch.no_grad():
        got = dit(hidden_states=lat, timestep=t, encoder_hidden_states=prompt, encoder_attention_mask=mask)
        want = ro(lat.float(), t, prompt.float(), mask)
    rel = ((got - want).norm() / want.norm()).item()
    print(f"unconditioned forward rel-L2 {rel:.4g}")
    assert rel < 1e-2


@pytest.mark.parametrize("target_ffn", [False, True])
def test_activation_stash_matches_full_recompute(setup, monkeypatch, target_ffn):
    """The spare-HBM activation stash (engine._ensure_stash) only replaces recomputation by stored copies of the same
    values: adapter gradients with the stash off, partial (some blocks) and full must agree to run-to-run noise (the
    attention backward sums its sub-block contributions in a timing-dependent order, so not bit for bit)."""
    from longcat_video_tta_b200 import lora
    from longcat_video_tta_b200.stepper import TTAStepper
    s = setup
    monkeypatch.setenv("B200TTA_DETERMINISTIC", "1")   # compare runs: take the run-to-run GEMM noise out
    (sigma, eps), = replay_draws(s["train"], 1)
    cond, train, prompt = (s[k].to(BF16).cuda() for k in ("cond", "train", "prompt"))
    mask, sigma, eps = s["mask"].cuda(), sigma.cuda(), eps.to(BF16).cuda()
    results = {}
    for label, cap in (("off", "0"), ("partial", None), ("full", "4")):
        dit = s["B200DiT"].from_oracle(s["oracle"])
        torch.manual_seed(5)
        mods = lora.inject_lora_into_dit(dit, rank=8, alpha=16.0, target_modules=["qkv", "proj"], target_ffn=target_ffn)
        gen = torch.Generator().manual_seed(9)
        with torch.no_grad():
            for m in mods:
                m.lora_up.weight.copy_((torch.randn(m.lora_up.weight.shape, generator=gen) * 0.02).to(BF16).cuda())
        st = TTAStepper(dit)
        eng = dit.engine
        if cap is None:
            # a cap that holds x1/x2 everywhere, qkv for some blocks and h1/h3 for none
            geo = st._geometry(cond, train, eng.pack_text(prompt, mask))
            per_x = geo.N * eng.C * 2
            cap = str((2 * eng.L * per_x + (eng.L // 2) * 3 * per_x + per_x) / float(1 << 30))
        monkeypatch.setenv("B200TTA_STASH_GB", cap)
        loss = st.forward_backward(cond, train, prompt, mask, sigma, eps).item()
        stash = eng._stash
        counts = {k: stash[k] for k in ("k_x1", "k_x2", "k_qkv", "k_h")}
        results[label] = (loss, [g.clone() for site in eng.lora_sites() for g in site.param_grads()], counts)
    assert results["off"][2] == dict(k_x1=0, k_x2=0, k_qkv=0, k_h=0)
    L = len(s["B200DiT"].from_oracle(s["oracle"]).blocks)
    assert results["full"][2] == dict(k_x1=L, k_x2=L, k_qkv=L, k_h=L)
    part = results["partial"][2]
    assert part["k_x1"] == L and part["k_x2"] == L and 0 < part["k_qkv"] < L and part["k_h"] == 0
    for label in ("partial", "full"):
        assert abs(results[label][0] - results["off"][0]) <= 1e-6 * abs(results["off"][0])
        worst = max(((a - b).norm() / (b.norm() + 1e-30)).item() for a, b in zip(results[label][1], results["off"][1]))
        print(f"stash={label}: worst per-tensor relative L2 difference to full recompute {worst:.3g}")
        assert worst < 1e-3, f"stash={label}: gradient differs from the full recompute"


def test_deterministic_switch_gives_the_same_bits_every_run(setup, monkeypatch):
    """B200TTA_DETERMINISTIC=1 (single-issuer GEMM): two forwards of the same inputs agree bit for bit.  The default
    CTA-pair GEMM sums its k-blocks in a timing-dependent order, so there the two runs only agree to rounding."""
    s = setup
    (sigma, eps), = replay_draws(s["train"], 1)
    cond, train, prompt = (s[k].to(BF16).cuda() for k in ("cond", "train", "prompt"))
    mask, sigma = s["mask"].cuda(), sigma.cuda()
    dit = s["B200DiT"].from_oracle(s["oracle"])
    lat = torch.cat([cond, train], dim=2)
    t = torch.full((1, lat.shape[2]), 421.0, device="cuda", dtype=BF16)

    def fwd():
        with torch.no_grad():
            return dit(hidden_states=lat, timestep=t, encoder_hidden_states=prompt, encoder_attention_mask=mask)

    monkeypatch.setenv("B200TTA_DETERMINISTIC", "1")
    a, b = fwd(), fwd()
    assert torch.equal(a, b)
    monkeypatch.delenv("B200TTA_DETERMINISTIC")
    c = fwd()
    rel = ((c.float() - a.float()).norm() / a.float().norm()).item()
    print(f"default (CTA-pair) GEMM vs deterministic GEMM forward: rel-L2 {rel:.3g}")
    assert rel < 5e-3


def test_cuda_graph_replay_matches_eager(setup, monkeypatch):
    """TTAStepper(cuda_graph=True): first step eager, second captured, later ones replayed from static input buffers --
    same losses and the same post-step adapter parameters as the launch-by-launch path on identical draws."""
    from longcat_video_tta_b200 import lora
    from longcat_video_tta_b200.stepper import TTAStepper
    s = setup
    monkeypatch.setenv("B200TTA_DETERMINISTIC", "1")
    cond, train, prompt = (s[k].to(BF16).cuda() for k in ("cond", "train", "prompt"))
    mask = s["mask"].cuda()
    draws = replay_draws(s["train"], 5)
    out = {}
    for mode in (False, True):
        dit = s["B200DiT"].from_oracle(s["oracle"])
        torch.manual_seed(7)
        mods = lora.inject_lora_into_dit(dit, rank=16, alpha=32.0, target_modules=["qkv", "proj"])
        st = TTAStepper(dit, cuda_graph=mode)
        losses = [st.step(cond, train, prompt, mask, sg.cuda(), e.to(BF16).cuda(), lora._warmup_lr(2e-4, i, 3)).item()
                  for i, (sg, e) in enumerate(draws)]
        out[mode] = (losses, [p.detach().float().clone() for p in lora.get_lora_parameters(mods)])
        if mode:
            assert st._graph_state["calls"] == 5 and "graph" in st._graph_state
    print(f"cuda graph vs eager: losses {out[True][0]} vs {out[False][0]}")
    # step 0 runs eagerly in both modes and step 1 is the first replay on identical parameters up to the kernels' own
    # run-to-run summation order: tight.  Later steps sit behind AdamW's m / sqrt(v), which turns rounding-level
    # gradient differences on near-zero entries into +-lr parameter differences: loose.
    for i, (a, b) in enumerate(zip(out[False][0], out[True][0])):
        assert abs(a - b) <= (1e-5 if i < 2 else 2e-3) * abs(a), (i, out[False][0], out[True][0])
    init = out[False][1]
    c = min(cos(a, b) for a, b in zip(out[False][1][1::2], out[True][1][1::2]))
    print(f"cuda graph vs eager: worst cosine between the trained lora_up tensors {c:.6f}")
    assert c > 0.99
