"""GPU parity of the loops / loss variants that had only host-side pinning in round 1:

  * finetune_lora_batch            lora_experiment/scripts/run_lora_tta.py:558-634  (round-robin over K host-resident videos)
  * compute_flow_matching_loss     delta_experiment/scripts/common.py:274-343       (unconditioned: BACKWARD with N_c = 0)
  * forward hooks on adaLN_modulation  delta_experiment/scripts/run_film_tta.py:146-163  (the reference FiLM seam through dit(...))

Oracle: the fp32 restatement with bf16-valued weights on the same GPU (TF32 off), sharpened to realistic attention
entropy where whole-model gradient directions are compared (see tests/test_block_parity_gpu.py for why)."""
import pytest
import torch

from parity_util import BF16, F32, COS_BAR, cos, build_pair, tiny_case

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


def _videos(k):
    """K different pre-encoded videos on the HOST, as run_lora_tta.py:1037-1062 prepares them (different latents,
    different prompts, different numbers of valid text tokens)."""
    out = []
    for i in range(k):
        g = torch.Generator().manual_seed(100 + i)
        mask = torch.zeros(1, 512, dtype=torch.int64)
        mask[:, : 64 + 32 * i] = 1
        out.append(dict(cond_latents=torch.randn(1, 16, 2, 32, 32, generator=g).to(BF16),
                        train_latents=torch.randn(1, 16, 2, 32, 32, generator=g).to(BF16),
                        prompt_embeds=torch.randn(1, 1, 512, 512, generator=g).to(BF16), prompt_mask=mask))
    return out


def test_finetune_lora_batch_matches_oracle_loop():
    from oracle import tta_oracle as T
    from longcat_video_tta_b200 import lora
    oracle, omods, dit, mods = build_pair("tiny", 0, sharpen=2.0, target_ffn=False)
    with torch.no_grad():   # the reference starts from B = 0
        for m, om in zip(mods, omods):
            m.lora_up.weight.zero_()
            om.lora_up.weight.zero_()
    vids, steps, lr, warm = _videos(3), 6, 2e-4, 3
    init = [p.detach().float().clone() for p in lora.get_lora_parameters(mods)]
    torch.manual_seed(42)
    torch.cuda.manual_seed(42)
    out = lora.finetune_lora_batch(dit, mods, vids, num_steps=steps, lr=lr, warmup_steps=warm, device="cuda", dtype=BF16)
    assert set(out) == {"losses", "train_time", "es_check_time", "early_stopping_info"} and len(out["losses"]) == steps
    # the same loop on the oracle: same device RNG stream, same video order, clip 1.0, AdamW(wd .01, eps 1e-8)
    params = T.lora_parameters(omods)
    for p in params:
        p.requires_grad_(True)
    m1 = [torch.zeros_like(p) for p in params]
    m2 = [torch.zeros_like(p) for p in params]
    torch.manual_seed(42)
    torch.cuda.manual_seed(42)
    olosses = []
    for step in range(steps):
        bd = vids[step % len(vids)]
        cond, train = bd["cond_latents"].cuda(), bd["train_latents"].cuda()
        pe, pm = bd["prompt_embeds"].cuda(), bd["prompt_mask"].cuda()
        sigma = torch.rand(1, device="cuda", dtype=F32) * (1.0 - 0.001) + 0.001
        noise = torch.randn_like(train)
        loss = T.fm_loss_given(oracle, cond, train, pe, pm, sigma, noise, BF16)
        grads = [g.clone() for g in torch.autograd.grad(loss, params)]
        T.clip_grad_norm(grads, 1.0)
        with torch.no_grad():
            T.adamw_step(params, grads, m1, m2, step + 1, T.warmup_lr(lr, step, warm), eps=1e-8, wd=0.01)
        olosses.append(loss.item())
    print("batch loop losses", out["losses"], "oracle", olosses)
    for a, b in zip(out["losses"], olosses):
        assert abs(a - b) <= 2e-2 * abs(b)
    got = torch.cat([(p.detach().float() - i).flatten() for p, i in zip(lora.get_lora_parameters(mods), init)])
    want = torch.cat([(p.detach() - i).flatten() for p, i in zip(params, init)])
    c = cos(got, want)
    print(f"batch loop: adapter update cosine vs the fp32 oracle loop {c:.5f}")
    assert c > 0.99   # six AdamW steps: m/sqrt(v) is sign-like on near-zero entries, so looser than the gradient bar


def test_unconditioned_loss_backward_matches_fp32_oracle():
    """N_c = 0: one attention segment, cross attention on every row, loss on every frame -- through the autograd seam
    the reference uses (pred = dit(...); loss.backward(); common.py:328-343)."""
    from oracle import tta_oracle as T
    from longcat_video_tta_b200 import lora
    oracle, omods, dit, mods = build_pair("tiny", 0, sharpen=2.0, target_ffn=False)
    cond, train, prompt, mask, sigma, eps4 = tiny_case()
    lat = torch.cat([cond, train], dim=2)
    g = torch.Generator(device="cuda").manual_seed(5)
    noise = torch.randn(lat.shape, generator=g, device="cuda").to(BF16)
    s = sigma.view(1, 1, 1, 1, 1)
    noisy = ((1.0 - s) * lat + s * noise).to(BF16)
    timestep = (sigma * 1000).unsqueeze(1).expand(1, lat.shape[2]).to(BF16)
    target = (noise - lat).float()
    params = lora.get_lora_parameters(mods)
    for p in params:
        p.requires_grad_(True)
    pred = dit(hidden_states=noisy, timestep=timestep, encoder_hidden_states=prompt, encoder_attention_mask=mask)
    loss = torch.nn.functional.mse_loss(pred.float(), target)
    loss.backward()
    oparams = T.lora_parameters(omods)
    for p in oparams:
        p.requires_grad_(True)
    opred = oracle(hidden_states=noisy, timestep=timestep, encoder_hidden_states=prompt, encoder_attention_mask=mask)
    oloss = torch.nn.functional.mse_loss(opred.float(), target)
    ograds = torch.autograd.grad(oloss, oparams)
    print(f"unconditioned: loss {loss.item():.6f} vs fp32 oracle {oloss.item():.6f}")
    assert abs(loss.item() - oloss.item()) <= 2e-2 * abs(oloss.item())
    cs = [cos(p.grad, og) for p, og in zip(params, ograds)]
    print("unconditioned adapter-gradient cosines:", " ".join(f"{c:.5f}" for c in cs))
    assert min(cs) > COS_BAR
    for p, og in zip(params, ograds):
        assert abs(p.grad.float().norm().item() / og.norm().item() - 1) < 2e-2


@pytest.mark.parametrize("mode", ["full", "scale_only"])
def test_reference_style_film_hooks_through_dit_call(mode):
    """Hooks of the reference's form on every adaLN_modulation (run_film_tta.py:146-163): dit(...) must (i) apply them
    in the forward, (ii) return their gradients from loss.backward() -- equal to what adapters.FiLMAdapterWrapper (the
    fused-stepper form of the same adapter) computes, and to the fp32 oracle carrying the same hooks."""
    from oracle import tta_oracle as T
    from longcat_video_tta_b200 import adapters
    oracle, _, dit, _ = build_pair("tiny", 0, sharpen=2.0, target_ffn=False, rank=16)
    # no LoRA in this test: undo the injection on both models
    for model in (dit, oracle):
        for blk in model.blocks:
            for parent, names in ((blk.attn, ("qkv", "proj")), (blk.cross_attn, ("q_linear", "kv_linear", "proj"))):
                for n in names:
                    setattr(parent, n, getattr(parent, n).original)
    C = dit.config.hidden_size
    w = adapters.FiLMAdapterWrapper(dit, num_groups=2, hidden_size=C, film_mode=mode)
    gen = torch.Generator(device="cuda").manual_seed(3)
    with torch.no_grad():
        for c in w.corrections:
            c.copy_(torch.randn(c.shape, generator=gen, device="cuda") * 0.05)
    cond, train, prompt, mask, sigma, eps = tiny_case()

    def hooks_on(model, corrections):
        hs = []
        for b, blk in enumerate(model.blocks):
            corr = corrections[w._get_group_idx(b)]

            def _make(correction):
                def _hook(_m, _i, output):
                    return output + w._expand_correction(correction).unsqueeze(0).unsqueeze(0).to(output.dtype)
                return _hook
            hs.append(blk.adaLN_modulation.register_forward_hook(_make(corr)))
        return hs

    # (a) fused-stepper form
    for c in w.corrections:
        c.requires_grad_(True)
    la = T.fm_loss_given(w, cond, train, prompt, mask, sigma, eps, BF16)
    ga = torch.autograd.grad(la, list(w.corrections))
    # (b) hooks on the bare dit
    hs = hooks_on(dit, w.corrections)
    lb = T.fm_loss_given(dit, cond, train, prompt, mask, sigma, eps, BF16)
    gb = torch.autograd.grad(lb, list(w.corrections))
    for h in hs:
        h.remove()
    assert abs(la.item() - lb.item()) <= 1e-5 * abs(la.item())
    for a, b in zip(ga, gb):
        assert cos(a, b) > 0.9999     # two runs of the same kernels: run-to-run summation order only
    # (c) fp32 oracle with the same hooks
    ocorr = [torch.nn.Parameter(c.detach().float().clone()) for c in w.corrections]
    hs = hooks_on(oracle, ocorr)
    lo = T.fm_loss_given(oracle, cond, train, prompt, mask, sigma, eps, BF16)
    go = torch.autograd.grad(lo, ocorr)
    for h in hs:
        h.remove()
    print(f"FiLM hooks ({mode}): loss {lb.item():.6f} vs fp32 oracle {lo.item():.6f}; gradient cosines "
          + " ".join(f"{cos(a, b):.5f}" for a, b in zip(gb, go)))
    assert abs(lb.item() - lo.item()) <= 2e-2 * abs(lo.item())
    assert min(cos(a, b) for a, b in zip(gb, go)) > COS_BAR
    # forward only with hooks (generation after TTA): no autograd graph needed
    hs = hooks_on(dit, [c.detach() for c in w.corrections])
    with torch.no_grad():
        hidden, timestep, n_cond = T.build_step_inputs(cond, train, sigma, eps, BF16)
        p1 = dit(hidden, timestep, prompt, mask, num_cond_latents=n_cond)
        p2 = w(hidden, timestep, prompt, mask, num_cond_latents=n_cond)
    for h in hs:
        h.remove()
    assert torch.equal(p1, p2) or ((p1 - p2).norm() / p2.norm()).item() < 1e-3
