"""Test infrastructure, not product code.  Generates tests/golden/split_budget.pt: the reference's own
``estimate_tta_split_budget`` (delta_experiment/scripts/common.py:1493-1517, imported from /root/reference through
oracle/ref_bridge.py) over a grid of pixel-frame budgets, hold-out fractions and VAE time scales.

Run here (needs /root/reference):  python oracle/make_golden_budget.py"""
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_bridge  # noqa: E402


def main():
    cm = ref_bridge.load("common")
    table = {}
    for total in list(range(1, 40)) + [49, 61, 77, 93, 117, 121, 200]:
        for ctx in (1, 2, 5, 13, 14, 17, 33, 93, 500):
            for hf in (0.0, 0.1, 0.25, 0.5, 0.9):
                for scale in (4, 8):
                    table[(total, ctx, hf, scale)] = cm.estimate_tta_split_budget(total, ctx, holdout_fraction=hf, vae_t_scale=scale)
    print(len(table), table[(117, 13, 0.25, 4)], table[(14, 14, 0.25, 4)])
    torch.save(table, ROOT / "tests" / "golden" / "split_budget.pt")


if __name__ == "__main__":
    main()
