#!/usr/bin/env python3
"""B200-native drop-in for the reference's `sweep_experiment/scripts/run_sweep.py`: same YAML schema, config keys and options; rows run as local
processes on the box's GPUs instead of SLURM jobs.  See longcat_video_tta_b200/sweep.py and INTEGRATION.md."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from longcat_video_tta_b200.sweep import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
