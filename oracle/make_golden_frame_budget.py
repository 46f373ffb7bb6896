"""Record how the reference resolves the TTA frame budget and guards it -> tests/golden/frame_budget.json.gz.

TEST INFRASTRUCTURE ONLY (reads /root/reference; run in the build container, commit the output).

Every method script runs the same block right after ``parser.parse_args()`` (run_lora_tta.py:743-758, run_delta_a.py
:414-431, ... identical text in all six): default ``tta_total_frames`` / ``tta_context_frames``, clamp to
``gen_start_frame``, then ``validate_tta_feature_budget`` (common.py:1533-1598).  The block is cut out of
run_lora_tta.py's ``main()`` and executed, unmodified, on a Namespace for a grid of settings, with the reference's own
``validate_tta_feature_budget``; the resolved frames, the returned info, what was printed and the error raised are kept.
"""
from __future__ import annotations

import argparse
import contextlib
import gzip
import io
import itertools
import json
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import ref_bridge  # noqa: E402


def reference_block():
    lines = (ref_bridge._LORA / "run_lora_tta.py").read_text().splitlines()
    start = next(i for i, l in enumerate(lines) if l.strip() == "args = parser.parse_args()") + 1
    stop = next(i for i in range(start, len(lines)) if "validate_tta_feature_budget(" in lines[i]) + 1
    return compile(textwrap.dedent("\n".join(lines[start:stop])), f"run_lora_tta.py:{start + 1}-{stop}", "exec")


def main():
    common = ref_bridge.load("common")
    block = reference_block()
    rows = []
    clip_settings = [dict(clip_gate_enabled=False),
                     dict(clip_gate_enabled=True, clip_gate_sampling_mode="full_window", clip_gate_sample_frames=4),
                     dict(clip_gate_enabled=True, clip_gate_late_only=True, clip_gate_late_fraction=0.4,
                          clip_gate_backend="xclip")]
    grid = itertools.product([2, 14, 24], [None, 2, 13, 17, 33, 93, 117], [None, 5, 13, 14, 200], [14, 32, 93, 200],
                             [False, True], ["fail", "warn", "off"], range(len(clip_settings)), [0.25, 0.5])
    for cond, total, ctx, gen_start, es_disable, mode, ci, holdout in grid:
        if holdout == 0.5 and (mode != "fail" or ci != 0):
            continue
        given = dict(num_cond_frames=cond, tta_total_frames=total, tta_context_frames=ctx, gen_start_frame=gen_start,
                     es_disable=es_disable, es_holdout_fraction=holdout, feature_frame_guard_mode=mode,
                     clip_gate_sampling_mode="full_window", clip_gate_late_only=False, clip_gate_late_fraction=0.4,
                     clip_gate_sample_frames=4, clip_gate_backend="clip")
        given.update(clip_settings[ci])
        args = argparse.Namespace(**given)
        got = {}

        def validate(a, context=""):
            got["info"] = common.validate_tta_feature_budget(a, context=context)

        buf, err = io.StringIO(), None
        with contextlib.redirect_stdout(buf):
            try:
                exec(block, {"args": args, "validate_tta_feature_budget": validate})
            except RuntimeError as e:
                err = str(e)
        rows.append({"given": given, "total": args.tta_total_frames, "context": args.tta_context_frames,
                     "info": got.get("info"), "error": err, "printed": buf.getvalue().splitlines()})
    path = ROOT / "tests" / "golden" / "frame_budget.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(rows, sort_keys=True).encode())
    print(f"{len(rows)} settings ({sum(r['error'] is not None for r in rows)} refused), {path.stat().st_size} bytes -> {path}")


if __name__ == "__main__":
    main()
