"""Test infrastructure, not product code.  Generates tests/golden/adapter_layout.pt: what the reference's own adapter
wrappers (delta_experiment/scripts/run_delta_{a,b,c}.py, run_norm_tune_tta.py, run_film_tta.py, imported from
/root/reference through oracle/ref_bridge.py) construct -- trainable tensor shapes and order, the block -> group maps,
target-block selection, zero padding of partial deltas, the expansion of partial FiLM corrections into the 6C adaLN
layout, and which norm parameters each ``norm_target`` collects (by name).  Pure host logic, no DiT arithmetic.

Run here (needs /root/reference):  python oracle/make_golden_adapter_layout.py"""
import pathlib
import sys

import torch
import torch.nn as nn

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_bridge  # noqa: E402
from oracle.dit_oracle import build_oracle_dit  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


class StubDiT(nn.Module):
    """only what the wrappers' constructors touch: .blocks (length), .parameters(), .config"""

    def __init__(self, n_blocks):
        super().__init__()
        self.blocks = nn.ModuleList([nn.Linear(1, 1) for _ in range(n_blocks)])
        self.config = None


DELTA_B_GRID = [
    dict(n_blocks=48, num_groups=4, delta_target="timestep", delta_dim=None, target_blocks="all"),
    dict(n_blocks=48, num_groups=1, delta_target="timestep", delta_dim=None, target_blocks="all"),
    dict(n_blocks=48, num_groups=5, delta_target="timestep", delta_dim=128, target_blocks="last_12"),
    dict(n_blocks=7, num_groups=3, delta_target="hidden", delta_dim=64, target_blocks="0,2,6"),
    dict(n_blocks=48, num_groups=48, delta_target="hidden", delta_dim=4096, target_blocks="all"),
    dict(n_blocks=2, num_groups=4, delta_target="timestep", delta_dim=None, target_blocks="all"),
]
FILM_GRID = [dict(n_blocks=nb, num_groups=g, hidden_size=4, film_mode=mode)
             for nb, g in ((48, 4), (48, 1), (7, 3), (2, 4)) for mode in ("full", "shift_scale", "scale_only")]


def main():
    da, db, dc = ref_bridge.load("run_delta_a"), ref_bridge.load("run_delta_b"), ref_bridge.load("run_delta_c")
    nt, fm = ref_bridge.load("run_norm_tune_tta"), ref_bridge.load("run_film_tta")
    out = {}
    w = da.DeltaAWrapper(StubDiT(3), adaln_tembed_dim=512)
    out["delta_a"] = {"shape": tuple(w.delta.shape), "dtype": str(w.delta.dtype), "trainable": [n for n, p in w.named_parameters() if p.requires_grad]}
    w = dc.DeltaCWrapper(StubDiT(3), mode="per_channel", out_channels=16)
    out["delta_c"] = {"shape": tuple(w.delta_out.shape), "dtype": str(w.delta_out.dtype), "trainable": [n for n, p in w.named_parameters() if p.requires_grad]}
    out["delta_b"] = []
    for kw in DELTA_B_GRID:
        w = db.DeltaBWrapper(StubDiT(kw["n_blocks"]), num_groups=kw["num_groups"], adaln_tembed_dim=512, hidden_size=4096,
                             delta_target=kw["delta_target"], delta_dim=kw["delta_dim"], target_blocks=kw["target_blocks"])
        probe = torch.arange(1, w.deltas[0].shape[0] + 1, dtype=torch.float32)
        out["delta_b"].append({
            "kwargs": kw, "block_to_group": list(w.block_to_group),
            "target_block_indices": None if w.target_block_indices is None else sorted(w.target_block_indices),
            "delta_shapes": [tuple(p.shape) for p in w.deltas],
            "delta_final": None if w.delta_final is None else tuple(w.delta_final.shape),
            "padded": w._pad_delta(probe).clone(),
            "trainable": [n for n, p in w.named_parameters() if p.requires_grad],
        })
    out["film"] = []
    for kw in FILM_GRID:
        w = fm.FiLMAdapterWrapper(StubDiT(kw["n_blocks"]), num_groups=kw["num_groups"], hidden_size=kw["hidden_size"],
                                  film_mode=kw["film_mode"])
        probe = torch.arange(1, w.correction_dim + 1, dtype=torch.float32)
        out["film"].append({
            "kwargs": kw, "correction_dim": w.correction_dim, "shapes": [tuple(p.shape) for p in w.corrections],
            "group_of_block": [w._get_group_idx(i) for i in range(kw["n_blocks"])],
            "expanded": w._expand_correction(probe).clone(),
        })
    dit = build_oracle_dit("tiny", seed=0)
    names = {id(p): n for n, p in dit.named_parameters()}
    out["norm"] = {t: [(names[id(p)], tuple(p.shape)) for p in nt.collect_norm_params(dit, t)]
                   for t in ("cross_attn_norm", "qk_norm", "all_norm")}
    for k in ("delta_a", "delta_c"):
        print(k, out[k])
    for e in out["delta_b"]:
        print("delta_b", e["kwargs"], e["block_to_group"][:8], e["target_block_indices"], e["delta_shapes"][:1], e["delta_final"])
    for e in out["film"][:3]:
        print("film", e["kwargs"], e["group_of_block"][:8], e["expanded"].tolist())
    print({k: len(v) for k, v in out["norm"].items()}, out["norm"]["all_norm"][:6])
    torch.save(out, GOLDEN / "adapter_layout.pt")


if __name__ == "__main__":
    main()
