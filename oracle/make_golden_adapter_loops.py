"""Test infrastructure, not product code.  Generates tests/golden/adapter_loops.pt: the reference's own ``optimize_*``
loops for delta-A / delta-B / delta-C / norm-tune / FiLM (delta_experiment/scripts/run_delta_{a,b,c}.py,
run_norm_tune_tta.py, run_film_tta.py, imported from /root/reference through oracle/ref_bridge.py) run on CPU with the
DiT arithmetic replaced by a recorder.  Recorded: the optimizer's hyper-parameters, the learning rate at every
``optimizer.step()``, how gradients are clipped (one call over all tensors, or one call per tensor for delta-B), what the
loop feeds the DiT per step (variant picked, sigma / noise drawn, in the order drawn), where a stopper cuts the loop, and
the keys of the returned dict.

Run here (needs /root/reference):  python oracle/make_golden_adapter_loops.py"""
import pathlib
import sys
import types

import torch
import torch.nn as nn

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_bridge  # noqa: E402
from oracle.make_golden_loop_schedule import ScriptedStopper, videos  # noqa: E402

SEED = 2468


class StubDiT(nn.Module):
    def __init__(self, n_blocks=4):
        super().__init__()
        self.blocks = nn.ModuleList([nn.Linear(1, 1) for _ in range(n_blocks)])
        self.config = types.SimpleNamespace(patch_size=(1, 2, 2))


def variants_for(v):
    g = torch.Generator().manual_seed(13)
    return [{"latents": v["train_latents"], "name": "orig"}, {"latents": torch.randn(1, 16, 2, 4, 4, generator=g), "name": "aug"}]


def main():
    cm = ref_bridge.load("common")
    mods = {k: ref_bridge.load(n) for k, n in (("delta_a", "run_delta_a"), ("delta_b", "run_delta_b"), ("delta_c", "run_delta_c"),
                                              ("norm", "run_norm_tune_tta"), ("film", "run_film_tta"))}
    rec = {"adamw": [], "lrs": [], "clips": [], "calls": []}
    trainables = []

    class RecordingAdamW(torch.optim.AdamW):
        def __init__(self, params, **kw):
            super().__init__(params, **kw)
            g = self.param_groups[0]
            rec["adamw"].append({"lr": g["lr"], "betas": tuple(g["betas"]), "eps": g["eps"], "weight_decay": g["weight_decay"],
                                 "n_tensors": len(g["params"])})

        def step(self, *a, **k):
            rec["lrs"].append(float(self.param_groups[0]["lr"]))
            return super().step(*a, **k)

    real_clip = torch.nn.utils.clip_grad_norm_

    def recording_clip(params, max_norm, *a, **k):
        params = [params] if isinstance(params, torch.Tensor) else list(params)
        rec["clips"].append((len(params), float(max_norm)))
        return real_clip(params, max_norm, *a, **k)

    def recording_loss(dit, cond_latents, target_latents, prompt_embeds, prompt_mask, device="cuda", dtype=torch.bfloat16, **kw):
        def fwd(hidden, timestep, n_cond):
            rec["calls"].append({"hidden": hidden.detach().clone(), "timestep": timestep.detach().clone(), "n_cond": n_cond})
            return hidden.to(torch.float32) * (1.0 + sum(p.float().sum() for p in trainables))
        return cm.compute_flow_matching_loss_conditioned(dit=dit, cond_latents=cond_latents, target_latents=target_latents,
                                                         prompt_embeds=prompt_embeds, prompt_mask=prompt_mask, device=device,
                                                         dtype=dtype, forward_fn=fwd, **kw)

    torch.nn.utils.clip_grad_norm_ = recording_clip
    for m in mods.values():
        m.AdamW = RecordingAdamW
        m.compute_flow_matching_loss_conditioned = recording_loss

    v = videos(1)[0]
    args = (v["cond_latents"], v["train_latents"], v["prompt_embeds"], v["prompt_mask"])
    out = {}

    def run(name, fn, wrapper, params, extra=(), **kw):
        for k in rec:
            rec[k].clear()
        trainables[:] = params
        torch.manual_seed(SEED)
        r = fn(wrapper, *extra, *args, device="cpu", dtype=torch.float32, **kw)
        out[name] = {"adamw": list(rec["adamw"]), "lrs": list(rec["lrs"]), "clips": list(rec["clips"]), "calls": list(rec["calls"]),
                     "keys": sorted(r.keys()), "n_losses": len(r["losses"]), "early_stopping_info": r["early_stopping_info"],
                     "stopper_steps": getattr(kw.get("early_stopper"), "seen", None)}
        print(name, out[name]["adamw"], out[name]["lrs"][:3], out[name]["clips"][:6], out[name]["keys"], out[name]["stopper_steps"])

    w = mods["delta_a"].DeltaAWrapper(StubDiT(), adaln_tembed_dim=8)
    run("delta_a", mods["delta_a"].optimize_delta_a, w, [w.delta], num_steps=9, lr=5e-3, early_stopper=ScriptedStopper(),
        train_latents_variants=variants_for(v))
    w = mods["delta_b"].DeltaBWrapper(StubDiT(), num_groups=3, adaln_tembed_dim=8, hidden_size=16, delta_target="hidden", delta_dim=4)
    run("delta_b", mods["delta_b"].optimize_delta_b, w, list(w.deltas.parameters()) + [w.delta_final], num_steps=3, lr=1e-3)
    w = mods["delta_c"].DeltaCWrapper(StubDiT(), mode="per_channel", out_channels=16)
    run("delta_c", mods["delta_c"].optimize_delta_c, w, [w.delta_out], num_steps=3, lr=2e-3, train_latents_variants=variants_for(v))
    dit = StubDiT()
    norm_params = [dit.blocks[0].weight, dit.blocks[0].bias, dit.blocks[3].weight]
    w = mods["norm"].NormTuneForward(dit)
    run("norm", mods["norm"].optimize_norm_params, w, norm_params, extra=(norm_params,), num_steps=3, lr=1e-4)
    w = mods["film"].FiLMAdapterWrapper(StubDiT(), num_groups=2, hidden_size=4, film_mode="shift_scale")
    run("film", mods["film"].optimize_film_adapter, w, list(w.corrections.parameters()), num_steps=3, lr=1e-3)
    out["variants"] = [x["latents"] for x in variants_for(v)]
    torch.nn.utils.clip_grad_norm_ = real_clip
    torch.save(out, ROOT / "tests" / "golden" / "adapter_loops.pt")


if __name__ == "__main__":
    main()
