"""Sweep layer of the drop-in: the reference's YAML configs drive the B200 method scripts.

The reference turns ``{method, series, series_name, fixed{}, sweep[{run_id, ...}]}`` into one SLURM job per row:
config keys -> environment names (``sweep_experiment/scripts/run_sweep.py:51-136``, ``build_env_vars`` :169-209) -> bash
defaults and flag assembly (``sweep_experiment/sbatch/run_sweep.sbatch:28-147``, ``:196-631``) -> the command line of
``run_lora_tta.py`` / ``run_delta_*.py`` / ``run_norm_tune_tta.py`` / ``run_film_tta.py``.  The config keys and the
resulting command lines are the contract; SLURM itself is not rebuilt.  Here the same two steps are host Python
(``build_env_vars`` -> ``command_line``), and rows run as local processes, optionally spread over the GPUs of the box.
``tests/test_sweep_cpu.py`` holds both steps to what the reference's own code produces for all 260 rows of its 64 configs
(``tests/golden/sweep_rows.json.gz`` <- ``oracle/make_golden_sweep.py``).
"""
from __future__ import annotations

import os
import shlex
import subprocess
import sys
from pathlib import Path
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

REPO = Path(__file__).resolve().parents[1]

METHODS = ("full", "lora", "delta_a", "delta_b", "delta_c", "norm_tune", "film")
SCRIPTS = {
    "full": "lora_experiment/scripts/run_full_tta.py", "lora": "lora_experiment/scripts/run_lora_tta.py",
    "delta_a": "delta_experiment/scripts/run_delta_a.py", "delta_b": "delta_experiment/scripts/run_delta_b.py",
    "delta_c": "delta_experiment/scripts/run_delta_c.py", "norm_tune": "delta_experiment/scripts/run_norm_tune_tta.py",
    "film": "delta_experiment/scripts/run_film_tta.py",
}

# Config keys a YAML may carry (anything else is warned about and skipped, run_sweep.py:192-196).  The environment name
# of a key is its upper-case form.
CONFIG_KEYS = frozenset("""
    num_cond_frames num_frames gen_start_frame tta_total_frames tta_context_frames num_inference_steps guidance_scale
    resolution seed max_videos batch_videos batch_method retrieval_pool_dir
    learning_rate num_steps warmup_steps weight_decay max_grad_norm
    lora_rank lora_alpha target_modules lora_target_blocks use_builtin_lora target_ffn
    delta_steps delta_lr num_groups delta_target delta_dim delta_target_blocks delta_mode
    norm_steps norm_lr norm_target also_tune_delta film_steps film_lr film_mode optimizer
    es_disable es_check_every es_patience es_anchor_sigmas es_noise_draws es_strategy es_holdout_fraction
    clip_gate_enabled clip_gate_threshold clip_gate_backend clip_gate_model clip_gate_sample_frames
    clip_gate_aggregation clip_gate_sampling_mode clip_gate_late_fraction clip_gate_late_only clip_gate_fail_open
    clip_gate_log_only
    caption_guard_mode caption_guard_min_nonempty_ratio caption_guard_min_unique_ratio caption_guard_max_top1_ratio
    caption_guard_max_generic_top1_ratio caption_guard_topk fixed_caption feature_frame_guard_mode
    compute_fvd compute_fid compute_vbench min_fvd_videos skip_generation
""".split())

# What a job sees for a variable the sweep did not set (run_sweep.sbatch:46-147), as the literal strings bash would pass on.
JOB_DEFAULTS = {
    "NUM_COND_FRAMES": "2", "NUM_FRAMES": "16", "GEN_START_FRAME": "32", "NUM_INFERENCE_STEPS": "50",
    "GUIDANCE_SCALE": "4.0", "RESOLUTION": "480p", "SEED": "42", "MAX_VIDEOS": "100",
    "ES_DISABLE": "", "ES_CHECK_EVERY": "5", "ES_PATIENCE": "3", "ES_ANCHOR_SIGMAS": "0.25,0.5,0.75",
    "ES_NOISE_DRAWS": "2", "ES_STRATEGY": "patience", "ES_HOLDOUT_FRACTION": "0.25", "NO_SAVE_VIDEOS": "1",
    "LEARNING_RATE": "2e-4", "NUM_STEPS": "20", "WARMUP_STEPS": "3", "WEIGHT_DECAY": "0.01", "MAX_GRAD_NORM": "1.0",
    "OPTIMIZER": "sgd", "LORA_RANK": "8", "LORA_ALPHA": "16", "TARGET_MODULES": "qkv,proj", "LORA_TARGET_BLOCKS": "all",
    "USE_BUILTIN_LORA": "", "TARGET_FFN": "", "SKIP_GENERATION": "", "BATCH_VIDEOS": "1", "BATCH_METHOD": "similarity",
    "RETRIEVAL_POOL_DIR": "", "DELTA_STEPS": "20", "DELTA_LR": "1e-3", "NUM_GROUPS": "4", "DELTA_TARGET": "timestep",
    "DELTA_DIM": "", "DELTA_TARGET_BLOCKS": "all", "DELTA_MODE": "per_channel",
    "CLIP_GATE_ENABLED": "", "CLIP_GATE_THRESHOLD": "0.0", "CLIP_GATE_BACKEND": "clip",
    "CLIP_GATE_MODEL": "openai/clip-vit-large-patch14", "CLIP_GATE_SAMPLE_FRAMES": "4", "CLIP_GATE_AGGREGATION": "mean",
    "CLIP_GATE_SAMPLING_MODE": "full_window", "CLIP_GATE_LATE_FRACTION": "0.4", "CLIP_GATE_LATE_ONLY": "",
    "CLIP_GATE_FAIL_OPEN": "1", "CLIP_GATE_LOG_ONLY": "",
    "CAPTION_GUARD_MODE": "fail", "CAPTION_GUARD_MIN_NONEMPTY_RATIO": "0.95", "CAPTION_GUARD_MIN_UNIQUE_RATIO": "0.10",
    "CAPTION_GUARD_MAX_TOP1_RATIO": "0.50", "CAPTION_GUARD_MAX_GENERIC_TOP1_RATIO": "0.20", "CAPTION_GUARD_TOPK": "5",
    "FIXED_CAPTION": "", "FEATURE_FRAME_GUARD_MODE": "fail",
    "COMPUTE_FVD": "0", "COMPUTE_FID": "0", "COMPUTE_VBENCH": "0", "MIN_FVD_VIDEOS": "256",
    "NORM_STEPS": "20", "NORM_LR": "1e-3", "NORM_TARGET": "all_norm", "ALSO_TUNE_DELTA": "",
    "FILM_STEPS": "20", "FILM_LR": "1e-3", "FILM_MODE": "full",
}


def load_config(path) -> dict:
    """A sweep YAML with the reference's schema (run_sweep.py:150-166); raises ValueError instead of exiting."""
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f)
    for key in ("method", "series", "series_name", "fixed", "sweep"):
        if key not in cfg:
            raise ValueError(f"missing required key '{key}' in {path}")
    if cfg["method"] not in METHODS:
        raise ValueError(f"unknown method '{cfg['method']}' in {path}; valid: {list(METHODS)}")
    return cfg


def build_env_vars(method: str, series_name: str, run_id: str, fixed: Mapping, run_overrides: Mapping,
                   data_dir: Optional[str] = None, output_base: Optional[str] = None,
                   warn=lambda msg: print(msg, file=sys.stderr)) -> Dict[str, str]:
    """Variables of one row (run_sweep.py:169-209): the row overrides ``fixed``; a true boolean becomes "1", a false one
    leaves the variable unset; everything else is ``str(value)``."""
    env = {"METHOD": method, "RUN_ID": run_id, "SERIES_NAME": series_name}
    if data_dir:
        env["DATA_DIR"] = data_dir
    if output_base:
        env["OUTPUT_DIR"] = f"{output_base}/{series_name}/{run_id}"
    merged = dict(fixed)
    merged.update({k: v for k, v in run_overrides.items() if k != "run_id"})
    for key, value in merged.items():
        if key not in CONFIG_KEYS:
            warn(f"WARNING: Unknown config key '{key}', skipping.")
        elif isinstance(value, bool):
            if value:
                env[key.upper()] = "1"
        else:
            env[key.upper()] = str(value)
    return env


def _pairs(v: Mapping[str, str], *names: str) -> List[str]:
    out: List[str] = []
    for n in names:
        out += ["--" + n.lower().replace("_", "-"), v[n]]
    return out


def command_line(env: Mapping[str, str], project_root: str, checkpoint_dir: str, data_dir: str) -> Tuple[str, List[str]]:
    """(script, arguments) a job of the reference would start for ``env`` (run_sweep.sbatch:196-631).  A variable that is
    unset OR empty takes the job default, as ``${X:-default}`` does; flag groups the template expands unquoted are
    whitespace-split the same way.  One deviation: a ``fixed_caption`` is handed over as a single argument (the
    template's unquoted expansion would split one containing spaces)."""
    v = dict(JOB_DEFAULTS)
    v.update({k: s for k, s in env.items() if s != ""})
    method = v["METHOD"]
    if method not in SCRIPTS:
        raise ValueError(f"unknown METHOD '{method}'; use {'|'.join(METHODS)}")
    for dep in ("TTA_TOTAL_FRAMES", "TTA_CONTEXT_FRAMES"):
        v.setdefault(dep, v["NUM_COND_FRAMES"])
    v.setdefault("SERIES_NAME", "sweep")
    out_dir = v.get("OUTPUT_DIR") or f"{project_root}/sweep_experiment/results/{v['SERIES_NAME']}/{v['RUN_ID']}"
    on = lambda name: v[name] != ""                                       # noqa: E731  ([ -n "${X}" ])
    split = lambda words: [w for s in words for w in s.split()]           # noqa: E731  (unquoted expansion)

    es = (["--es-disable"] if on("ES_DISABLE") else []) + split(_pairs(
        v, "ES_CHECK_EVERY", "ES_PATIENCE", "ES_ANCHOR_SIGMAS", "ES_NOISE_DRAWS", "ES_STRATEGY", "ES_HOLDOUT_FRACTION"))
    save = ["--no-save-videos"] if v["NO_SAVE_VIDEOS"] == "1" else []
    clip = (["--clip-gate-enabled"] if on("CLIP_GATE_ENABLED") else []) + split(_pairs(
        v, "CLIP_GATE_THRESHOLD", "CLIP_GATE_BACKEND", "CLIP_GATE_MODEL", "CLIP_GATE_SAMPLE_FRAMES", "CLIP_GATE_AGGREGATION",
        "CLIP_GATE_SAMPLING_MODE", "CLIP_GATE_LATE_FRACTION"))
    clip += ["--clip-gate-late-only"] if on("CLIP_GATE_LATE_ONLY") else []
    clip += ["--clip-gate-log-only"] if on("CLIP_GATE_LOG_ONLY") else []
    clip += ["--clip-gate-fail-open" if on("CLIP_GATE_FAIL_OPEN") else "--clip-gate-fail-closed"]
    caption = split(_pairs(v, "CAPTION_GUARD_MODE", "CAPTION_GUARD_MIN_NONEMPTY_RATIO", "CAPTION_GUARD_MIN_UNIQUE_RATIO",
                           "CAPTION_GUARD_MAX_TOP1_RATIO", "CAPTION_GUARD_MAX_GENERIC_TOP1_RATIO", "CAPTION_GUARD_TOPK"))
    caption += ["--fixed-caption", v["FIXED_CAPTION"]] if on("FIXED_CAPTION") else []
    feature = split(_pairs(v, "FEATURE_FRAME_GUARD_MODE"))
    evalf = [f"--compute-{m.lower()}" for m in ("FVD", "FID", "VBENCH") if v[f"COMPUTE_{m}"] == "1"]
    evalf += split(_pairs(v, "MIN_FVD_VIDEOS"))
    pool = split(_pairs(v, "RETRIEVAL_POOL_DIR")) if on("RETRIEVAL_POOL_DIR") else []
    skip = ["--skip-generation"] if on("SKIP_GENERATION") else []

    head = ["--checkpoint-dir", v.get("CHECKPOINT_DIR") or checkpoint_dir, "--data-dir", v.get("DATA_DIR") or data_dir,
            "--output-dir", out_dir] + _pairs(v, "MAX_VIDEOS")
    frames = _pairs(v, "NUM_COND_FRAMES", "NUM_FRAMES", "GEN_START_FRAME", "TTA_TOTAL_FRAMES", "TTA_CONTEXT_FRAMES",
                    "NUM_INFERENCE_STEPS", "GUIDANCE_SCALE", "RESOLUTION", "SEED")
    training = _pairs(v, "LEARNING_RATE", "NUM_STEPS", "WARMUP_STEPS", "WEIGHT_DECAY", "MAX_GRAD_NORM")
    after_frames: List[str] = []
    extra: List[str] = []
    gate = clip
    if method == "full":
        own = training + _pairs(v, "OPTIMIZER", "BATCH_VIDEOS") + pool
        extra = skip
    elif method == "lora":
        own = _pairs(v, "LORA_RANK", "LORA_ALPHA", "TARGET_MODULES", "LORA_TARGET_BLOCKS") + training \
            + _pairs(v, "BATCH_VIDEOS") + pool
        extra = (["--target-ffn"] if on("TARGET_FFN") else []) + skip \
            + (["--use-builtin-lora"] if on("USE_BUILTIN_LORA") else [])
    elif method == "delta_a":
        own = _pairs(v, "DELTA_STEPS", "DELTA_LR", "BATCH_VIDEOS", "BATCH_METHOD") + pool
    elif method == "delta_b":
        own = _pairs(v, "DELTA_STEPS", "DELTA_LR", "NUM_GROUPS", "DELTA_TARGET", "DELTA_TARGET_BLOCKS")
        after_frames = split(_pairs(v, "DELTA_DIM")) if on("DELTA_DIM") else []
    elif method == "delta_c":
        own = _pairs(v, "DELTA_STEPS", "DELTA_LR", "DELTA_MODE")
    elif method == "norm_tune":
        own = _pairs(v, "NORM_STEPS", "NORM_LR", "NORM_TARGET")
        extra = ["--also-tune-delta"] if on("ALSO_TUNE_DELTA") else []
        gate = []                                                          # no CLIP gate on norm-tune / FiLM jobs
    else:
        own = _pairs(v, "FILM_STEPS", "FILM_LR", "NUM_GROUPS", "FILM_MODE")
        gate = []
    return SCRIPTS[method], head + own + frames + after_frames + save + extra + caption + feature + gate + evalf + es


def rows_of(cfg: Mapping, run_ids: Optional[Sequence[str]] = None) -> List[Mapping]:
    rows = [r for r in cfg["sweep"] if not run_ids or r["run_id"] in run_ids]
    if not rows:
        raise ValueError(f"no matching run IDs; available: {[r['run_id'] for r in cfg['sweep']]}")
    return rows


def run_rows(commands: Sequence[Tuple[str, str, List[str]]], gpus: int = 1, python: str = sys.executable,
             dry_run: bool = False, log_dir: Optional[Path] = None, root: Path = REPO) -> List[Dict]:
    """Run (run_id, script, argv) rows as local processes; with ``gpus`` > 1 up to that many rows run at a time, row i on
    GPU ``i % gpus`` (``CUDA_VISIBLE_DEVICES``).  Returns [{run_id, returncode}] in row order."""
    done: List[Dict] = [None] * len(commands)     # type: ignore[list-item]
    running: Dict[int, Tuple[int, subprocess.Popen, object]] = {}
    nxt = 0
    while nxt < len(commands) or running:
        while nxt < len(commands) and len(running) < max(1, gpus):
            run_id, script, argv = commands[nxt]
            gpu = nxt % max(1, gpus) if dry_run else next(g for g in range(max(1, gpus)) if g not in running)
            cmd = [python, str(root / script), *argv]
            if dry_run:
                print(f"  [DRY-RUN] CUDA_VISIBLE_DEVICES={gpu} {shlex.join(cmd)}")
                done[nxt] = {"run_id": run_id, "returncode": None}
            else:
                if not (root / script).is_file():
                    raise NotImplementedError(f"{script} is not part of this build")
                env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(gpu)) if gpus > 1 else None
                log = open(log_dir / f"{run_id}.log", "w") if log_dir else None
                print(f"  Starting {run_id} on GPU {gpu}: {shlex.join(cmd)}")
                running[gpu] = (nxt, subprocess.Popen(cmd, env=env, stdout=log, stderr=subprocess.STDOUT if log else None), log)
            nxt += 1
        for gpu, (i, proc, log) in list(running.items()):
            try:
                rc = proc.wait(timeout=0.5)
            except subprocess.TimeoutExpired:
                continue
            if log is not None:
                log.close()
            done[i] = {"run_id": commands[i][0], "returncode": rc}
            del running[gpu]
    return done


def main(argv: Optional[Sequence[str]] = None) -> int:
    """``run_sweep.py --config X.yaml [--run-ids ...] [--dry-run]`` with the reference's options (run_sweep.py:354-376);
    the SLURM-only ones (--account, --time, --sbatch-template) are accepted and unused."""
    import argparse
    p = argparse.ArgumentParser(description="Run TTA sweep rows from a YAML config on the local B200s")
    p.add_argument("--config", type=str, required=True)
    p.add_argument("--account", type=str, default=None, help="(SLURM) accepted, unused")
    p.add_argument("--dry-run", action="store_true")
    p.add_argument("--run-ids", nargs="+", type=str, default=None)
    p.add_argument("--data-dir", type=str, default=None)
    p.add_argument("--output-base", type=str, default=None)
    p.add_argument("--time", type=str, default=None, help="(SLURM) accepted, unused")
    p.add_argument("--sbatch-template", type=str, default=None, help="(SLURM) accepted, unused")
    p.add_argument("--checkpoint-dir", type=str, default=os.environ.get("CHECKPOINT_DIR", ""))
    p.add_argument("--gpus", type=int, default=1, help="rows run concurrently, one per GPU")
    p.add_argument("--script-args", type=str, default="",
                   help="extra arguments for every method script, e.g. '--synthetic --model tiny --latent-hw 32,32'")
    a = p.parse_args(argv)
    cfg = load_config(a.config)
    rows = rows_of(cfg, a.run_ids)
    print(f"TTA sweep: {cfg.get('description', cfg['series_name'])}\n  Method : {cfg['method']}\n"
          f"  Series : {cfg['series']} ({cfg['series_name']})\n  Runs   : {len(rows)}")
    commands = []
    for row in rows:
        env = build_env_vars(cfg["method"], cfg["series_name"], row["run_id"], cfg["fixed"], row, a.data_dir, a.output_base)
        for passthrough in ("NO_SAVE_VIDEOS", "CHECKPOINT_DIR", "DATA_DIR", "OUTPUT_DIR"):   # sbatch --export=ALL
            if passthrough in os.environ:
                env.setdefault(passthrough, os.environ[passthrough])
        script, args = command_line(env, str(REPO), a.checkpoint_dir, a.data_dir or "")
        commands.append((row["run_id"], script, args + a.script_args.split()))
    log_dir = None
    if not a.dry_run:
        log_dir = Path(a.output_base or REPO / "sweep_experiment" / "results") / cfg["series_name"] / "logs"
        log_dir.mkdir(parents=True, exist_ok=True)
    done = run_rows(commands, gpus=a.gpus, dry_run=a.dry_run, log_dir=log_dir)
    print(f"Summary: {len(done)} rows {'would be ' if a.dry_run else ''}run")
    for d in done:
        print(f"  {d['run_id']:>8} -> {'(dry-run)' if d['returncode'] is None else 'exit ' + str(d['returncode'])}")
    return 0 if all(d["returncode"] in (None, 0) for d in done) else 1
