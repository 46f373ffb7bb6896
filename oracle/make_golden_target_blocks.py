"""Test infrastructure, not product code.  Generates tests/golden/target_blocks.pt: the reference's own
``_parse_target_blocks`` (lora_experiment/scripts/run_lora_tta.py:263-283 and its copy in
delta_experiment/scripts/run_delta_b.py) over a grid of specs, including the invalid ones (recorded as the exception type).

Run here (needs /root/reference):  python oracle/make_golden_target_blocks.py"""
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_bridge  # noqa: E402

SPECS = ["all", "ALL", " all ", "last_1", "last_4", "last_48", "last_49", "last_0", "Last_2", "0", "47", "48", "-1",
         "0,1,2", "5, 7 ,9", "3,3,3", "1,0", "0,48", "last_x", "a,b", "", "last_", "2,", "last_-1"]


def record(fn, spec, n):
    try:
        r = fn(spec, n)
        return None if r is None else sorted(r)
    except Exception as e:  # noqa: BLE001
        return f"raises {type(e).__name__}"


def main():
    rl, db = ref_bridge.load("run_lora_tta"), ref_bridge.load("run_delta_b")
    out = {"lora": {}, "delta_b": {}}
    for n in (1, 2, 48):
        for spec in SPECS:
            out["lora"][(spec, n)] = record(rl._parse_target_blocks, spec, n)
            out["delta_b"][(spec, n)] = record(db._parse_target_blocks, spec, n)
    diff = {k for k in out["lora"] if out["lora"][k] != out["delta_b"][k]}
    print(len(out["lora"]), "specs; lora vs delta_b copies differ on", sorted(diff))
    print({k: v for k, v in list(out["lora"].items()) if k[1] == 48 and isinstance(v, str)})
    torch.save(out, ROOT / "tests" / "golden" / "target_blocks.pt")


if __name__ == "__main__":
    main()
