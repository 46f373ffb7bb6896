"""Generate ``tests/golden/*.pt`` by running the reference's own, unmodified TTA code
(imported from /root/reference via ``oracle/ref_bridge.py``) on the oracle DiT.

TEST INFRASTRUCTURE ONLY; container-only (needs /root/reference).  Re-run with
    python -m oracle.make_golden
Everything is CPU fp32, BASELINE.json configs[0] ("tiny LongCat-Video DiT (2 blocks,
hidden 512, random init) LoRA r=16 TTA step on synthetic 17-frame 256x256 latent").

Inputs are regenerated from seeds by ``tiny_inputs`` (also used by the tests), so the
fixtures hold only outputs: per-step losses, the step-0 prediction, step-0 clipped
adapter gradients and post-step adapter parameters.
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import ref_bridge  # noqa: E402
from oracle.dit_oracle import build_oracle_dit  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
LOOP_SEED = 42


def tiny_inputs(device="cpu", dtype=torch.float32):
    """latent [1,16,5,32,32] seed 1 (17 px frames @256x256),
    text [1,1,512,512] seed 2 with the first 128 tokens valid."""
    g1 = torch.Generator().manual_seed(1)
    latents = torch.randn(1, 16, 5, 32, 32, generator=g1)
    g2 = torch.Generator().manual_seed(2)
    prompt = torch.randn(1, 1, 512, 512, generator=g2)
    mask = torch.zeros(1, 512, dtype=torch.int64)
    mask[:, :128] = 1
    return latents.to(device, dtype), prompt.to(device, dtype), mask.to(device)


def tiny_split(latents):
    """tta_total=17, tta_context=5 px frames -> num_ctx_lat = 1 + (5-1)//4 = 2
    (run_lora_tta.py:1088-1093) -> cond 2, train 2, val 1."""
    from oracle.tta_oracle import split_tta_latents
    return split_tta_latents(latents, 2, 0.25)


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    cm = ref_bridge.load("common")
    rl = ref_bridge.load("run_lora_tta")

    latents, prompt, mask = tiny_inputs()
    cond, train, val = cm.split_tta_latents(latents, 2, 0.25)
    assert (cond.shape[2], train.shape[2], val.shape[2]) == (2, 2, 1)

    # ---- split table (pure index arithmetic) --------------------------------------------
    table = {}
    for T in range(2, 34):
        for ctx in (1, 2, 4, 8):
            for hf in (0.25, 0.5):
                c, t, v = cm.split_tta_latents(torch.zeros(1, 1, T, 1, 1), ctx, hf)
                table[(T, ctx, hf)] = (c.shape[2], t.shape[2], 0 if v is None else v.shape[2])
    torch.save(table, GOLDEN / "split_table.pt")

    # ---- LoRA r=16 alpha=32 qkv,proj all blocks; lr 2e-4 warm-up 3 wd .01 clip 1 ---------
    def lora_run(num_steps, builtin=False):
        dit = build_oracle_dit("tiny", seed=0)
        torch.manual_seed(7)  # adapter init stream (kaiming A)
        if builtin:
            mods = rl.inject_builtin_lora_into_dit(dit, rank=16, alpha=32.0, target_modules=["qkv", "proj"])
            params = rl.get_builtin_lora_parameters(mods)
            for p in params:
                p.requires_grad_(True)
        else:
            mods = rl.inject_lora_into_dit(dit, rank=16, alpha=32.0, target_modules=["qkv", "proj"])
            params = rl.get_lora_parameters(mods)
        preds = []
        h = dit.register_forward_hook(lambda m, i, o: preds.append(o.detach().clone()))
        torch.manual_seed(LOOP_SEED)
        out = rl.finetune_lora_on_conditioning(
            dit, mods, cond, train, prompt, mask, num_steps=num_steps, lr=2e-4, warmup_steps=3,
            weight_decay=0.01, max_grad_norm=1.0, device="cpu", dtype=torch.float32,
            lora_param_fn=(lambda: params) if builtin else None)
        h.remove()
        return dit, mods, params, preds, out

    _, _, params1, preds1, out1 = lora_run(1)
    _, _, params5, _, out5 = lora_run(5)
    assert abs(out1["losses"][0] - out5["losses"][0]) == 0.0
    torch.save({
        "config": dict(rank=16, alpha=32.0, targets=["qkv", "proj"], lr=2e-4, warmup=3, wd=0.01, clip=1.0,
                       adapter_seed=7, loop_seed=LOOP_SEED, dit_seed=0),
        "losses": out5["losses"],
        "pred_step0": preds1[0],
        "clipped_grads_step0": [p.grad.detach().clone() for p in params1],
        "params_after_1": [p.detach().clone() for p in params1],
        "params_after_5": [p.detach().clone() for p in params5],
    }, GOLDEN / "lora_tiny.pt")
    print("lora losses", out5["losses"])

    _, _, paramsb, _, outb = lora_run(2, builtin=True)
    torch.save({"losses": outb["losses"], "params_after_2": [p.detach().clone() for p in paramsb]},
               GOLDEN / "lora_builtin_tiny.pt")
    print("builtin lora losses", outb["losses"])

    # ---- delta-A / delta-B / delta-C / norm-tune / FiLM: 3 steps, lr 1e-3 -----------------
    res = {}
    da = ref_bridge.load("run_delta_a")
    w = da.DeltaAWrapper(build_oracle_dit("tiny", seed=0), adaln_tembed_dim=512)
    torch.manual_seed(LOOP_SEED)
    o = da.optimize_delta_a(w, cond, train, prompt, mask, num_steps=3, lr=1e-3, device="cpu", dtype=torch.float32)
    res["delta_a"] = {"losses": o["losses"], "params": [w.delta.detach().clone()]}

    db = ref_bridge.load("run_delta_b")
    w = db.DeltaBWrapper(build_oracle_dit("tiny", seed=0), num_groups=2, adaln_tembed_dim=512, hidden_size=512,
                         delta_target="timestep")
    torch.manual_seed(LOOP_SEED)
    o = db.optimize_delta_b(w, cond, train, prompt, mask, num_steps=3, lr=1e-3, device="cpu", dtype=torch.float32)
    res["delta_b_timestep_g2"] = {"losses": o["losses"], "params": [d.detach().clone() for d in w.deltas]}

    w = db.DeltaBWrapper(build_oracle_dit("tiny", seed=0), num_groups=2, adaln_tembed_dim=512, hidden_size=512,
                         delta_target="hidden", delta_dim=512)
    torch.manual_seed(LOOP_SEED)
    o = db.optimize_delta_b(w, cond, train, prompt, mask, num_steps=3, lr=1e-3, device="cpu", dtype=torch.float32)
    res["delta_b_hidden_g2"] = {"losses": o["losses"],
                                "params": [d.detach().clone() for d in w.deltas] + [w.delta_final.detach().clone()]}

    dc = ref_bridge.load("run_delta_c")
    w = dc.DeltaCWrapper(build_oracle_dit("tiny", seed=0), mode="per_channel", out_channels=16)
    torch.manual_seed(LOOP_SEED)
    o = dc.optimize_delta_c(w, cond, train, prompt, mask, num_steps=3, lr=1e-3, device="cpu", dtype=torch.float32)
    res["delta_c"] = {"losses": o["losses"], "params": [w.delta_out.detach().clone()]}

    nt = ref_bridge.load("run_norm_tune_tta")
    dit = build_oracle_dit("tiny", seed=0)
    nparams = nt.collect_norm_params(dit, "all_norm")
    for p in nparams:
        p.requires_grad_(True)
    w = nt.NormTuneForward(dit)
    torch.manual_seed(LOOP_SEED)
    o = nt.optimize_norm_params(w, nparams, cond, train, prompt, mask, num_steps=3, lr=1e-3, device="cpu",
                                dtype=torch.float32)
    res["norm_all"] = {"losses": o["losses"], "params": [p.detach().clone() for p in nparams]}

    fm = ref_bridge.load("run_film_tta")
    w = fm.FiLMAdapterWrapper(build_oracle_dit("tiny", seed=0), num_groups=2, hidden_size=512, film_mode="full")
    w.apply_to_dit()  # run_film_tta.py:441 -- corrections act through adaLN_modulation forward hooks
    torch.manual_seed(LOOP_SEED)
    o = fm.optimize_film_adapter(w, cond, train, prompt, mask, num_steps=3, lr=1e-3, device="cpu", dtype=torch.float32)
    res["film_full_g2"] = {"losses": o["losses"], "params": [c.detach().clone() for c in w.corrections]}

    for k, v in res.items():
        print(k, v["losses"])
    torch.save(res, GOLDEN / "delta_tiny.pt")

    # ---- early-stopper anchor loss (forward-only, fixed sigmas / noises) ------------------
    es = ref_bridge.load("early_stopping")
    stopper = es.AnchoredEarlyStopper(check_every=1, patience=2, anchor_sigmas=[0.25, 0.5, 0.75], noise_draws=2)
    dit = build_oracle_dit("tiny", seed=0)
    stopper.setup(dit, cond, val, prompt, mask, device="cpu", dtype=torch.float32, video_id="golden_video")
    torch.save({"anchor_loss0": stopper.best_loss, "video_id": "golden_video"}, GOLDEN / "anchor_tiny.pt")
    print("anchor", stopper.best_loss)


if __name__ == "__main__":
    main()
