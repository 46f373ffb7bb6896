"""Record the command-line surface of the reference's TTA scripts -> tests/golden/cli_flags.json.

TEST INFRASTRUCTURE ONLY (reads /root/reference; run in the build container, commit the output).

Every reference script builds its parser at the top of ``main()`` (run_lora_tta.py:654-741, run_delta_a.py:370-412,
run_delta_b.py:451-497, run_delta_c.py:253-286, run_norm_tune_tta.py:290-320, run_film_tta.py:346-374; run_full_tta.py
:324-388 is recorded for information, full-model TTA is out of scope).  The statements from ``parser = ArgumentParser``
up to ``parser.parse_args()`` are executed, unmodified, in the namespace of the bridged module (oracle/ref_bridge.py),
so the shared groups (common.py:1404-1706,2438-2450; early_stopping.py:33-51) are the reference's own functions.  For
every action the option strings, dest, default, type, choices, nargs, const, required flag and action class are kept.
"""
from __future__ import annotations

import argparse
import json
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import ref_bridge  # noqa: E402

SCRIPTS = {
    "lora": ("run_lora_tta", ref_bridge._LORA),
    "full": ("run_full_tta", ref_bridge._LORA),
    "delta_a": ("run_delta_a", ref_bridge._DELTA),
    "delta_b": ("run_delta_b", ref_bridge._DELTA),
    "delta_c": ("run_delta_c", ref_bridge._DELTA),
    "norm_tune": ("run_norm_tune_tta", ref_bridge._DELTA),
    "film": ("run_film_tta", ref_bridge._DELTA),
}


def reference_parser(method: str) -> argparse.ArgumentParser:
    name, folder = SCRIPTS[method]
    lines = (folder / f"{name}.py").read_text().splitlines()
    start = next(i for i, l in enumerate(lines) if l.strip().startswith("parser = argparse.ArgumentParser"))
    stop = next(i for i in range(start, len(lines)) if "parser.parse_args()" in lines[i])
    src = textwrap.dedent("\n".join(lines[start:stop]))
    if name == "run_full_tta":      # not bridged as a module (out of scope): only the shared groups are needed
        ns = dict(vars(ref_bridge.load("common")))
        ns.update(vars(ref_bridge.load("early_stopping")))
    else:
        ns = dict(vars(ref_bridge.load(name)))
    ns["argparse"] = argparse
    exec(compile(src, f"{name}.py:{start + 1}-{stop}", "exec"), ns)
    return ns["parser"]


def describe(parser: argparse.ArgumentParser):
    out = []
    for a in parser._actions:
        if isinstance(a, argparse._HelpAction):
            continue
        out.append({
            "flags": list(a.option_strings), "dest": a.dest, "default": a.default,
            "type": getattr(a.type, "__name__", None) if a.type is not None else None,
            "choices": list(a.choices) if a.choices is not None else None, "nargs": a.nargs, "const": a.const,
            "required": bool(a.required), "action": type(a).__name__,
        })
    return out


def main():
    table = {m: describe(reference_parser(m)) for m in SCRIPTS}
    path = ROOT / "tests" / "golden" / "cli_flags.json"
    path.write_text(json.dumps(table, indent=1, sort_keys=True) + "\n")
    for m, acts in table.items():
        print(f"{m}: {len(acts)} arguments")
    print("wrote", path)


if __name__ == "__main__":
    main()
