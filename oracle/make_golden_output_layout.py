"""Record the layout of the files the reference's method scripts write -> tests/golden/output_layout.json.

TEST INFRASTRUCTURE ONLY (reads /root/reference; run in the build container, commit the output).

``config.json`` (LoRA only: ``exp_config`` run_lora_tta.py:855-908) and ``summary.json`` (``summary`` literal near the end
of every ``main()``: run_lora_tta.py:1277-1322, run_delta_a.py:905-, run_delta_b.py:918-955, run_delta_c.py:674-,
run_norm_tune_tta.py:631-, run_film_tta.py:676-) are dict literals inside ``main()``, which cannot be executed here
(checkpoints, datasets).  Their KEYS -- what the reference's exporters and audit scripts read -- are taken from the source
text with a brace-depth scanner that survives the four snapshot-truncated scripts (``truncated: true`` marks a literal
whose tail is missing; the keys up to the cut are kept).  The constant ``"method"`` value is recorded as well, and the keys of
the per-video ``result`` record each main loop appends to ``checkpoint.json`` / ``summary.json``.
"""
from __future__ import annotations

import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle.make_golden_cli_flags import SCRIPTS  # noqa: E402


def dict_keys(src: str, opener: str):
    """Nested {key: sub-keys | None} of the dict literal that follows ``opener``; second value: was it cut short."""
    at = src.find(opener)
    if at < 0:
        return None, False
    i = at + len(opener) - 1          # on the '{'
    stack, root = [], None
    pending = None                    # last string constant seen at key position
    n = len(src)
    while i < n:
        c = src[i]
        if c in "\"'":
            q = c
            j = i + 1
            while j < n and src[j] != q:
                j += 2 if src[j] == "\\" else 1
            if j >= n:
                return root, True
            text = src[i + 1:j]
            k = j + 1
            while k < n and src[k] in " \t":
                k += 1
            if stack and stack[-1][1] == "{" and k < n and src[k] == ":" and stack[-1][2] == "key":
                pending = text
                stack[-1][0][text] = None
                stack[-1][2] = "value"
                i = k + 1
                continue
            i = j + 1
            continue
        if c == "#":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if c in "{[(":
            if c == "{":
                d = {}
                if root is None:
                    root = d
                elif stack and stack[-1][1] == "{" and stack[-1][2] == "value" and pending is not None \
                        and stack[-1][0].get(pending, 0) is None and src[:i].rstrip().endswith(":"):
                    stack[-1][0][pending] = d
                stack.append([d, "{", "key"])
            else:
                stack.append([None, c, None])
        elif c in "}])":
            stack.pop()
            if not stack:
                return root, False
        elif c == "," and stack and stack[-1][1] == "{":
            stack[-1][2] = "key"
        i += 1
    return root, True


def main():
    table = {}
    for method, (name, folder) in SCRIPTS.items():
        src = (folder / f"{name}.py").read_text()
        summary, cut = dict_keys(src, "    summary = {")
        config, _ = dict_keys(src, "    exp_config = {")
        m = re.search(r'    summary = \{\s*"method": "([^"]+)"', src)
        at = re.search(r"\n +result = \{", src)          # the per-video record of the main loop
        result, _ = dict_keys(src[at.start():], "result = {")
        table[method] = {"summary_keys": list(summary), "summary_truncated": cut, "summary_method": m.group(1),
                         "config": config, "result_keys": list(result),
                         "result_keys_later": sorted(set(re.findall(r'result\["([a-z_0-9]+)"\]\s*=', src)))}
        print(method, len(summary), "summary keys", "(truncated)" if cut else "", "| config:", list(config) if config else None)
    path = ROOT / "tests" / "golden" / "output_layout.json"
    path.write_text(json.dumps(table, indent=1) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
