"""Record what the reference's sweep layer turns its own YAML configs into -> tests/golden/sweep_rows.json.gz.

TEST INFRASTRUCTURE ONLY (reads and executes files of /root/reference in the build container; commit the output).

Two stages, both the reference's own code, unmodified:
  1. ``sweep_experiment/scripts/run_sweep.py``: ``load_config`` + ``build_env_vars`` (:150-209) for every row of every
     YAML under ``sweep_experiment/configs`` -> the environment one SLURM job would receive (``_KEY_TO_ENV`` :51-136).
  2. ``sweep_experiment/sbatch/run_sweep.sbatch``: the path / default assignments (:28-147) and the flag assembly +
     ``case "${METHOD}"`` dispatch (:196-631) are cut out of the script by their section markers and run under ``bash``
     with that environment and ``$PYTHON`` bound to a function that writes its argument vector (NUL-separated) to a scratch file.
     The cluster-only middle (module / conda / nvidia-smi / cd) is skipped.
The result pins config keys -> environment names -> the exact command line of the method scripts.
"""
from __future__ import annotations

import contextlib
import gzip
import importlib.util
import io
import json
import os
import subprocess
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/sweep_experiment")


def _load_run_sweep():
    spec = importlib.util.spec_from_file_location("ref_run_sweep", REF / "scripts" / "run_sweep.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _dispatch_script() -> str:
    lines = (REF / "sbatch" / "run_sweep.sbatch").read_text().splitlines()
    find = lambda pred, lo=0: next(i for i in range(lo, len(lines)) if pred(lines[i]))  # noqa: E731
    a0 = find(lambda l: l.startswith("SCRATCH_BASE="))
    a1 = find(lambda l: l.startswith("echo "), a0)                      # first banner after the defaults
    b0 = find(lambda l: l.startswith("ES_FLAGS="), a1)
    b1 = find(lambda l: l.strip() == "esac", b0)
    head = ["set -euo pipefail", "emit_argv() { printf '%s\\0' \"$@\" > \"$ARGV_OUT\"; }"]
    return "\n".join(head + lines[a0:a1] + ["PYTHON=emit_argv"] + lines[b0:b1 + 1]) + "\n"


def main():
    rs = _load_run_sweep()
    script = _dispatch_script()
    rows, skipped = [], []
    with tempfile.TemporaryDirectory() as tmp:
        sh = Path(tmp) / "dispatch.sh"
        sh.write_text(script)
        for cfg_path in sorted((REF / "configs").glob("*.yaml")):
            err = io.StringIO()
            try:
                with contextlib.redirect_stderr(err):
                    cfg = rs.load_config(str(cfg_path))
            except SystemExit:
                skipped.append({"config": cfg_path.name, "why": err.getvalue().strip()})
                continue
            for row in cfg["sweep"]:
                err = io.StringIO()
                with contextlib.redirect_stderr(err):
                    env = rs.build_env_vars(cfg["method"], cfg["series_name"], row["run_id"], cfg["fixed"], row)
                out = Path(tmp) / "argv.bin"
                clean = {"PATH": os.environ["PATH"], "ARGV_OUT": str(out), **env}
                r = subprocess.run(["bash", str(sh)], env=clean, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
                if r.returncode != 0:
                    raise RuntimeError(f"{cfg_path.name}/{row['run_id']}: {r.stderr.decode()}")
                argv = out.read_bytes().decode().split("\0")[:-1]
                rows.append({"config": cfg_path.name, "method": cfg["method"], "series_name": cfg["series_name"],
                             "run_id": row["run_id"], "fixed": cfg["fixed"], "row": row, "env": env, "argv": argv,
                             "time": rs.estimate_time(cfg["method"], row, cfg["fixed"]), "mem": rs.estimate_mem(cfg["method"]),
                             "warnings": err.getvalue().strip().splitlines()})
    table = {"key_to_env": rs._KEY_TO_ENV, "methods": sorted(rs._METHOD_MAP), "rows": rows, "skipped": skipped}
    path = ROOT / "tests" / "golden" / "sweep_rows.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(table, sort_keys=True, default=str).encode())
    print(f"{len(rows)} rows of {len({r['config'] for r in rows})} configs, {len(skipped)} skipped; "
          f"{path.stat().st_size} bytes -> {path}")
    by = {}
    for r in rows:
        by[r["method"]] = by.get(r["method"], 0) + 1
    print(by)


if __name__ == "__main__":
    main()
