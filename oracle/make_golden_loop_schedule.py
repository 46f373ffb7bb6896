"""Test infrastructure, not product code.  Generates tests/golden/loop_schedule.pt: the reference's own LoRA loops
(lora_experiment/scripts/run_lora_tta.py:425-547 ``finetune_lora_on_conditioning`` with augmented variants and an early
stopper, :558-634 ``finetune_lora_batch`` round-robin over videos), imported from /root/reference through
oracle/ref_bridge.py and run on CPU with a one-parameter stand-in for the DiT.  Recorded per step: what the loop feeds
the DiT (``hidden_states`` / ``timestep`` / ``num_cond_latents`` -- i.e. which video or variant it picked and which
sigma / noise it drew, in the order it drew them) and the learning rate in force at ``optimizer.step()``; plus where the
early stopper cut the loop and the keys of the returned dict.

Run here (needs /root/reference):  python oracle/make_golden_loop_schedule.py"""
import pathlib
import sys
import types

import torch
import torch.nn as nn

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_bridge  # noqa: E402

SEED = 4321


class OneParamDiT(nn.Module):
    """pred = hidden * w: enough for loss.backward() / clip / AdamW to run; records what it is fed"""

    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.tensor(0.5))
        self.config = types.SimpleNamespace(patch_size=(1, 2, 2))
        self.calls = []

    def forward(self, hidden_states, timestep, encoder_hidden_states=None, encoder_attention_mask=None, num_cond_latents=0):
        self.calls.append({"hidden": hidden_states.detach().clone(), "timestep": timestep.detach().clone(),
                           "n_cond": num_cond_latents, "prompt_sum": float(encoder_hidden_states.float().sum())})
        return hidden_states.to(torch.float32) * self.w


def videos(n):
    g = torch.Generator().manual_seed(9)
    out = []
    for i in range(n):
        out.append({"cond_latents": torch.randn(1, 16, 1 + i % 2, 4, 4, generator=g), "train_latents": torch.randn(1, 16, 2, 4, 4, generator=g),
                    "prompt_embeds": torch.randn(1, 1, 4, 8, generator=g), "prompt_mask": torch.ones(1, 4, dtype=torch.int64) if i != 1 else None})
    return out


class ScriptedStopper:
    """stands in for AnchoredEarlyStopper (its own state machine is pinned by early_stopper_trace.pt): stops at step 7"""

    def __init__(self):
        self.seen, self.state = [], {"scripted": True}

    def step(self, current_step, save_fn=None):
        self.seen.append(current_step)
        if current_step == 2:
            self.snap = save_fn()
        return current_step >= 7, {"anchor_loss": 0.0}

    def restore(self, restore_fn=None):
        restore_fn(self.snap)


def main():
    rl = ref_bridge.load("run_lora_tta")
    lrs = []

    class RecordingAdamW(torch.optim.AdamW):
        def step(self, *a, **k):
            lrs.append(float(self.param_groups[0]["lr"]))
            return super().step(*a, **k)

    rl.AdamW = RecordingAdamW
    out = {}

    # ---- finetune_lora_batch: 3 videos round-robin, 8 steps, warm-up 3
    dit = OneParamDiT()
    torch.manual_seed(SEED)
    r = rl.finetune_lora_batch(dit, None, videos(3), num_steps=8, lr=2e-4, warmup_steps=3, device="cpu", dtype=torch.float32,
                               lora_param_fn=lambda: [dit.w])
    out["batch"] = {"calls": dit.calls, "lrs": list(lrs), "keys": sorted(r.keys()), "n_losses": len(r["losses"]),
                    "es_check_time": r["es_check_time"], "early_stopping_info": r["early_stopping_info"]}
    lrs.clear()

    # ---- finetune_lora_on_conditioning: 3 augmented variants (randint on the CPU stream), stopper cuts at step 7 of 12
    v = videos(1)[0]
    g = torch.Generator().manual_seed(11)
    variants = [{"latents": v["train_latents"], "name": "orig"}] + [
        {"latents": torch.randn(1, 16, 2, 4, 4, generator=g), "name": f"aug{i}"} for i in range(2)]
    dit = OneParamDiT()
    stopper = ScriptedStopper()
    torch.manual_seed(SEED)
    r = rl.finetune_lora_on_conditioning(dit, None, v["cond_latents"], v["train_latents"], v["prompt_embeds"], v["prompt_mask"],
                                         num_steps=12, lr=1e-3, warmup_steps=5, device="cpu", dtype=torch.float32,
                                         early_stopper=stopper, lora_param_fn=lambda: [dit.w], train_latents_variants=variants)
    out["single"] = {"calls": dit.calls, "lrs": list(lrs), "keys": sorted(r.keys()), "n_losses": len(r["losses"]),
                     "stopper_steps": stopper.seen, "early_stopping_info": r["early_stopping_info"],
                     "variants": [x["latents"] for x in variants]}
    lrs.clear()

    # ---- no warm-up, no variants, no stopper
    dit = OneParamDiT()
    torch.manual_seed(SEED)
    r = rl.finetune_lora_on_conditioning(dit, None, v["cond_latents"], v["train_latents"], v["prompt_embeds"], v["prompt_mask"],
                                         num_steps=4, lr=3e-4, warmup_steps=0, device="cpu", dtype=torch.float32,
                                         lora_param_fn=lambda: [dit.w])
    out["plain"] = {"calls": dit.calls, "lrs": list(lrs), "keys": sorted(r.keys()), "n_losses": len(r["losses"]),
                    "early_stopping_info": r["early_stopping_info"]}
    for k, e in out.items():
        print(k, len(e["calls"]), e["lrs"], e["keys"], e.get("stopper_steps"))
    torch.save(out, ROOT / "tests" / "golden" / "loop_schedule.pt")


if __name__ == "__main__":
    main()
