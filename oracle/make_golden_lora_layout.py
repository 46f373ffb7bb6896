"""Test infrastructure, not product code.  Generates tests/golden/lora_layout.pt: what the reference's own
``inject_lora_into_dit`` / ``inject_builtin_lora_into_dit`` / ``get_lora_parameters`` / ``count_lora_parameters`` /
``save_lora_weights`` / ``reset_lora_weights`` (lora_experiment/scripts/run_lora_tta.py:104-418, imported from
/root/reference through oracle/ref_bridge.py) produce on the tiny oracle DiT for a grid of target settings: which
linears are wrapped and in which order, the parameter list (shapes, in order), the counts, the checkpoint keys, and the
statistics of a reset.  The adapter ordering is what the checkpoint layout and the flat gradient buffer depend on.

Run here (needs /root/reference):  python oracle/make_golden_lora_layout.py"""
import os
import pathlib
import sys
import tempfile

import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_bridge  # noqa: E402
from oracle.dit_oracle import build_oracle_dit  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

GRID = [
    dict(rank=16, alpha=32.0, target_modules=["qkv", "proj"], target_ffn=False, target_blocks="all"),
    dict(rank=4, alpha=8.0, target_modules=["qkv"], target_ffn=False, target_blocks="all"),
    dict(rank=4, alpha=4.0, target_modules=["proj"], target_ffn=True, target_blocks="last_1"),
    dict(rank=8, alpha=16.0, target_modules=["qkv", "proj"], target_ffn=True, target_blocks="0"),
    dict(rank=2, alpha=1.0, target_modules=["qkv", "proj"], target_ffn=False, target_blocks="1,0"),
]


def wrapped_sites(dit, wrapper_types):
    """dotted names of the wrapped leaves, in module-tree order"""
    return [n for n, m in dit.named_modules() if isinstance(m, wrapper_types)]


def describe(rl, kw):
    dit = build_oracle_dit("tiny", seed=0)
    torch.manual_seed(3)
    mods = rl.inject_lora_into_dit(dit, rank=kw["rank"], alpha=kw["alpha"], target_modules=kw["target_modules"],
                                   target_ffn=kw["target_ffn"], target_blocks=kw["target_blocks"])
    params = rl.get_lora_parameters(mods)
    with tempfile.TemporaryDirectory() as d:
        rl.save_lora_weights(mods, os.path.join(d, "w.pt"))
        saved = torch.load(os.path.join(d, "w.pt"))
    out = {
        "kwargs": kw,
        "sites": wrapped_sites(dit, rl.LoRALinear),
        "module_io": [(m.original.in_features, m.original.out_features) for m in mods],
        "scaling": [float(m.scaling) for m in mods],
        "param_shapes": [tuple(p.shape) for p in params],
        "counts": rl.count_lora_parameters(mods),
        "checkpoint": [(k, tuple(v.shape)) for k, v in saved.items()],
        "up_is_zero_at_init": all(float(m.lora_up.weight.abs().max()) == 0.0 for m in mods),
        "down_bound_at_init": [float(m.lora_down.weight.abs().max()) for m in mods],   # kaiming_uniform(a=sqrt 5): < 1/sqrt(in)
    }
    for m in mods:
        m.lora_up.weight.data.fill_(1.0)
    rl.reset_lora_weights(mods)
    out["up_is_zero_after_reset"] = all(float(m.lora_up.weight.abs().max()) == 0.0 for m in mods)
    return out


def describe_builtin(rl, kw):
    dit = build_oracle_dit("tiny", seed=0)
    torch.manual_seed(3)
    mods = rl.inject_builtin_lora_into_dit(dit, rank=kw["rank"], alpha=kw["alpha"], target_modules=kw["target_modules"],
                                           target_ffn=kw["target_ffn"], target_blocks=kw["target_blocks"])
    params = rl.get_builtin_lora_parameters(mods)
    return {"kwargs": kw, "n_modules": len(mods), "names": [m.lora_name for m in mods],
            "param_shapes": [tuple(p.shape) for p in params]}


def main():
    rl = ref_bridge.load("run_lora_tta")
    out = {"custom": [describe(rl, kw) for kw in GRID], "builtin": [describe_builtin(rl, kw) for kw in GRID]}
    for c in out["custom"]:
        print(c["kwargs"]["target_blocks"], len(c["sites"]), c["counts"], c["checkpoint"][:2])
    for b in out["builtin"]:
        print("builtin", b["n_modules"], b["names"][:3], b["param_shapes"][:4])
    torch.save(out, GOLDEN / "lora_layout.pt")


if __name__ == "__main__":
    main()
