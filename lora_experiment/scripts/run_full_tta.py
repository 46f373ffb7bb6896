#!/usr/bin/env python3
"""B200-native drop-in for the reference's `lora_experiment/scripts/run_full_tta.py` (same flags and output files; every DiT
parameter trains, the step runs on the sm_100a engine).  See longcat_video_tta_b200/cli.py, full.py and INTEGRATION.md."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from longcat_video_tta_b200.cli import run  # noqa: E402

if __name__ == "__main__":
    run("full")
