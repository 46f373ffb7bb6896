"""Test infrastructure, not product code.  Generates tests/golden/delta_a_batch_loop.pt: the reference's own
``_optimize_delta_a_batch`` (delta_experiment/scripts/run_delta_a.py:308-362, imported through oracle/ref_bridge.py) --
one shared delta vector trained round-robin over K pre-encoded videos held on the host -- run on CPU with the DiT
arithmetic replaced by a recorder.  Recorded like oracle/make_golden_adapter_loops.py: optimizer hyper-parameters, the
learning rate and clipping of every step, the DiT inputs per step (which video, sigma / noise in the order drawn) and the
keys of the returned dict.

Run here (needs /root/reference):  python oracle/make_golden_delta_a_batch.py"""
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_bridge  # noqa: E402
from oracle.make_golden_adapter_loops import SEED, StubDiT  # noqa: E402
from oracle.make_golden_loop_schedule import videos  # noqa: E402

K, STEPS, LR = 3, 7, 5e-3


def main():
    cm = ref_bridge.load("common")
    mod = ref_bridge.load("run_delta_a")
    rec = {"adamw": [], "lrs": [], "clips": [], "calls": []}

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

    w = mod.DeltaAWrapper(StubDiT(), adaln_tembed_dim=8)

    def recording_loss(dit, cond_latents, target_latents, prompt_embeds, prompt_mask, device="cuda", dtype=torch.bfloat16, **kw):
        def fwd(hidden, timestep, n_cond):
            rec["calls"].append({"hidden": hidden.detach().clone(), "timestep": timestep.detach().clone(), "n_cond": n_cond})
            return hidden.to(torch.float32) * (1.0 + w.delta.float().sum())
        return cm.compute_flow_matching_loss_conditioned(dit=dit, cond_latents=cond_latents, target_latents=target_latents,
                                                         prompt_embeds=prompt_embeds, prompt_mask=prompt_mask, device=device,
                                                         dtype=dtype, forward_fn=fwd, **kw)

    torch.nn.utils.clip_grad_norm_ = recording_clip
    mod.AdamW = RecordingAdamW
    mod.compute_flow_matching_loss_conditioned = recording_loss
    torch.manual_seed(SEED)
    r = mod._optimize_delta_a_batch(w, videos(K), num_steps=STEPS, lr=LR, device="cpu", dtype=torch.float32)
    torch.nn.utils.clip_grad_norm_ = real_clip
    out = {"adamw": rec["adamw"], "lrs": rec["lrs"], "clips": rec["clips"], "calls": rec["calls"], "keys": sorted(r.keys()),
           "n_losses": len(r["losses"]), "early_stopping_info": r["early_stopping_info"], "es_check_time": r["es_check_time"]}
    print(out["adamw"], out["lrs"], out["clips"], out["keys"], len(out["calls"]))
    torch.save(out, ROOT / "tests" / "golden" / "delta_a_batch_loop.pt")


if __name__ == "__main__":
    main()
