"""Command-line entry points with the reference's flags and output layout:

  lora_experiment/scripts/run_full_tta.py        --learning-rate --num-steps --warmup-steps --weight-decay --max-grad-norm
                                                 --optimizer {sgd,adamw}          (every DiT parameter trains)
  lora_experiment/scripts/run_lora_tta.py        --lora-rank --lora-alpha --target-ffn --target-modules
                                                 --lora-target-blocks --use-builtin-lora --save-lora-weights
                                                 --learning-rate --num-steps --warmup-steps --weight-decay --max-grad-norm
  delta_experiment/scripts/run_delta_{a,b,c}.py  --delta-steps --delta-lr [--num-groups --delta-target --delta-dim
                                                 --delta-target-blocks | --delta-mode]
  delta_experiment/scripts/run_norm_tune_tta.py  --norm-steps --norm-lr --norm-target
  delta_experiment/scripts/run_film_tta.py       --film-steps --film-lr --film-mode --num-groups

Outputs: ``config.json``, ``checkpoint.json`` ({next_idx, results}), ``summary.json``, ``lora_weights/<video>_lora.pt``.
Only the TTA step is in scope: video decoding, VAE / UMT5 encoding, generation and quality metrics belong to the upstream
LongCat-Video package, which is not vendored.  Without it (or with ``--synthetic``) the scripts run the identical loop on
seeded synthetic latents and a random-init DiT -- which is how the tests and the benchmark drive them.
"""
from __future__ import annotations

import argparse
import json
import os
import time
from pathlib import Path
from typing import Dict, List

import torch

from . import adapters as A
from . import dist as D
from . import full as F
from . import lora as L
from .common import resolve_tta_frames, split_tta_latents, validate_tta_feature_budget
from .dit import B200DiT
from .early_stopping import add_early_stopping_args, build_early_stopper_from_args

BF16 = torch.bfloat16
METHODS = ("lora", "full", "delta_a", "delta_b", "delta_c", "norm_tune", "film")


# ---- the reference's command-line surface, one row per flag: (flag, kind, default[, choices]) with kind in
# {int, float, str, "on" (store_true), "off:<dest>" (store_false into dest)}.  tests/test_cli_cpu.py holds every row to
# what the reference's own parsers declare (tests/golden/cli_flags.json <- oracle/make_golden_cli_flags.py).
_IO = [("--max-videos", int, 100), ("--seed", int, 42), ("--device", str, "cuda")]
_GENERATION = [("--num-cond-frames", int, 2), ("--num-frames", int, 16), ("--gen-start-frame", int, 32),
               ("--num-inference-steps", int, 50), ("--guidance-scale", float, 4.0), ("--resolution", str, "480p"),
               ("--skip-generation", "on", False), ("--no-save-videos", "on", False)]
_BATCH = [("--batch-videos", int, 1), ("--batch-method", str, "similarity", ["sequential", "similarity"]),
          ("--retrieval-pool-dir", str, None)]                                   # run_lora_tta.py:716-725
_AUGMENTATION = [("--aug-enabled", "on", False), ("--aug-flip", "on", False), ("--aug-rotate-deg", float, 10.0),
                 ("--aug-rotate-random-min", float, 5.0), ("--aug-rotate-random-max", float, 15.0),
                 ("--aug-rotate-random-count", int, 2), ("--aug-rotate-random-step", float, 1.0),
                 ("--no-aug-rotate-zoom", "off:aug_rotate_zoom", True), ("--aug-speed-factors", str, "")]   # common.py:1680-1706
_TTA_FRAMES = [("--tta-total-frames", int, None), ("--tta-context-frames", int, None)]                       # common.py:1404-1417
_GUARD_MODES = ["fail", "warn", "off"]
_CAPTION = [("--caption-guard-mode", str, "fail", _GUARD_MODES), ("--caption-guard-min-nonempty-ratio", float, 0.95),
            ("--caption-guard-min-unique-ratio", float, 0.10), ("--caption-guard-max-top1-ratio", float, 0.50),
            ("--caption-guard-max-generic-top1-ratio", float, 0.20), ("--caption-guard-topk", int, 5),
            ("--fixed-caption", str, None),                                      # common.py:1420-1472
            ("--feature-frame-guard-mode", str, "fail", _GUARD_MODES)]           # common.py:1475-1485
_ONLINE_EVAL = [("--compute-fvd", "on", False), ("--compute-fid", "on", False), ("--compute-vbench", "on", False),
                ("--min-fvd-videos", int, 256)]                                  # common.py:2438-2450
_CLIP_GATE = [("--clip-gate-enabled", "on", False), ("--clip-gate-threshold", float, 0.0),
              ("--clip-gate-backend", str, "clip", ["clip", "xclip"]),
              ("--clip-gate-model", str, "openai/clip-vit-large-patch14"), ("--clip-gate-sample-frames", int, 4),
              ("--clip-gate-aggregation", str, "mean", ["mean", "min", "max"]),
              ("--clip-gate-sampling-mode", str, "full_window", ["full_window", "late_only"]),
              ("--clip-gate-late-fraction", float, 0.4), ("--clip-gate-late-only", "on", False),
              ("--clip-gate-fail-open", "on:clip_gate_fail_open", True),
              ("--clip-gate-fail-closed", "off:clip_gate_fail_open", True),
              ("--clip-gate-log-only", "on", False)]                             # common.py:1601-1677
_METHOD_FLAGS = {
    "lora": [("--lora-rank", int, 8), ("--lora-alpha", float, 16.0), ("--lora-dropout", float, 0.0),
             ("--target-ffn", "on", False), ("--target-modules", str, "qkv,proj"), ("--lora-target-blocks", str, "all"),
             ("--use-builtin-lora", "on", False), ("--save-lora-weights", "on", False), ("--learning-rate", float, 2e-4),
             ("--num-steps", int, 20), ("--warmup-steps", int, 3), ("--weight-decay", float, 0.01),
             ("--max-grad-norm", float, 1.0)],                                   # run_lora_tta.py:671-698
    "full": [("--learning-rate", float, 1e-5), ("--num-steps", int, 10), ("--warmup-steps", int, 2),
             ("--weight-decay", float, 0.01), ("--max-grad-norm", float, 1.0),
             ("--optimizer", str, "sgd", ["sgd", "adamw"])],                    # run_full_tta.py:341-351
    "delta_a": [("--delta-steps", int, 20), ("--delta-lr", float, 1e-3)],       # run_delta_a.py:376-377
    "delta_b": [("--delta-steps", int, 20), ("--delta-lr", float, 1e-3), ("--num-groups", int, 4),
                ("--delta-target", str, "timestep", ["timestep", "hidden"]), ("--delta-dim", int, None),
                ("--delta-target-blocks", str, "all")],                          # run_delta_b.py:459-473
    "delta_c": [("--delta-steps", int, 20), ("--delta-lr", float, 1e-3),
                ("--delta-mode", str, "per_channel", ["per_channel"])],
    "norm_tune": [("--norm-steps", int, 20), ("--norm-lr", float, 1e-3),
                  ("--norm-target", str, "all_norm", ["cross_attn_norm", "qk_norm", "all_norm"]),
                  ("--also-tune-delta", "on", False)],                           # run_norm_tune_tta.py:296-312
    "film": [("--film-steps", int, 20), ("--film-lr", float, 1e-3),
             ("--film-mode", str, "full", ["full", "shift_scale", "scale_only"]), ("--num-groups", int, 4)],
}
# The reference registers the retrieval-batch flags for LoRA / delta-A only and the CLIP gate for everything except
# norm-tune / FiLM; the union is accepted everywhere here so that one sweep template drives all methods.


def _add_rows(target, rows):
    for row in rows:
        flag, kind, default = row[:3]
        if kind == "on":
            target.add_argument(flag, action="store_true", default=default)
        elif isinstance(kind, str) and kind.startswith(("on:", "off:")):
            target.add_argument(flag, action="store_true" if kind[1] == "n" else "store_false", dest=kind.split(":")[1],
                                default=default)
        else:
            target.add_argument(flag, type=kind, default=default, choices=row[3] if len(row) > 3 else None)


def build_parser(method: str) -> argparse.ArgumentParser:
    """Parser of ``run_<method>`` with the reference's flags, dests, types, defaults and choices.  Deviations, all
    supersets: ``--checkpoint-dir`` / ``--data-dir`` are required by the reference and only checked here once real-video
    mode is selected; ``--restart`` is accepted by every method (reference: LoRA / full only); ``--synthetic``,
    ``--model`` and ``--latent-hw`` are ours."""
    p = argparse.ArgumentParser(description=f"B200-native {method} TTA for LongCat-Video")
    p.add_argument("--checkpoint-dir", type=str, default=None)
    p.add_argument("--data-dir", type=str, default=None)
    p.add_argument("--output-dir", type=str, required=True)
    p.add_argument("--restart", action="store_true")
    _add_rows(p, _IO)
    _add_rows(p.add_argument_group(method), _METHOD_FLAGS[method])
    _add_rows(p.add_argument_group("Video continuation"), _GENERATION)
    _add_rows(p.add_argument_group("Retrieval-augmented batch TTA"), _BATCH)
    add_early_stopping_args(p)
    _add_rows(p.add_argument_group("Data augmentation"), _AUGMENTATION)
    _add_rows(p.add_argument_group("TTA frames"), _TTA_FRAMES)
    _add_rows(p.add_argument_group("Caption / feature-frame guards"), _CAPTION)
    _add_rows(p.add_argument_group("Online distributional metrics"), _ONLINE_EVAL)
    _add_rows(p.add_argument_group("CLIP gate"), _CLIP_GATE)
    ours = p.add_argument_group("B200 build")
    ours.add_argument("--synthetic", action="store_true", help="seeded synthetic latents + random-init DiT")
    ours.add_argument("--model", type=str, default="13.6b", choices=["13.6b", "tiny"])
    ours.add_argument("--latent-hw", type=str, default=None, help="synthetic latent H,W (default 60,104 = 480x832)")
    return p


def outside_the_step(args) -> Dict:
    """Flags that configure what sits either side of the TTA step in the reference (pixel-space augmentation, caption /
    feature guards, CLIP gate, online FVD / FID / VBench, retrieval pool).  They are parsed and written to
    ``config.json`` like the reference does (run_lora_tta.py:855-908) so that a sweep's records stay comparable; the
    subsystems themselves need the upstream package and are not part of this build."""
    v = vars(args)
    pick = lambda prefix: {k: v[k] for k in sorted(v) if k.startswith(prefix)}  # noqa: E731
    return {"augmentation": pick("aug_"), "caption_guard": {**pick("caption_guard_"), "fixed_caption": v["fixed_caption"]},
            "feature_frame_guard_mode": v["feature_frame_guard_mode"], "clip_gate": pick("clip_gate_"),
            "online_eval": {**pick("compute_"), "min_fvd_videos": v["min_fvd_videos"]},
            "batch": {"batch_videos": v["batch_videos"], "batch_method": v["batch_method"],
                      "retrieval_pool_dir": v["retrieval_pool_dir"]}}


def _clip_gate_flat(args) -> Dict:
    mode = "late_only" if args.clip_gate_late_only else args.clip_gate_sampling_mode
    return {"clip_gate_enabled": args.clip_gate_enabled, "clip_gate_threshold": args.clip_gate_threshold,
            "clip_gate_backend": args.clip_gate_backend, "clip_gate_model": args.clip_gate_model,
            "clip_gate_sample_frames": args.clip_gate_sample_frames, "clip_gate_aggregation": args.clip_gate_aggregation,
            "clip_gate_sampling_mode": mode, "clip_gate_late_fraction": args.clip_gate_late_fraction,
            "clip_gate_log_only": args.clip_gate_log_only, "clip_gate_fail_open": args.clip_gate_fail_open}


def experiment_config(method: str, args, adapter_cfg: Dict, frames: Dict) -> Dict:
    """``config.json``.  For LoRA the reference's layout (run_lora_tta.py:855-906: ``method`` = ``lora_tta_<impl>``,
    ``lora{}``, ``training{}``, ``generation{}``, seed, max_videos, the CLIP-gate settings flat and nested); the delta
    scripts write no config file in the reference, ours get the same frame.  Additions: ``tta_frames``, ``synthetic``,
    ``model`` and the recorded-only groups of ``outside_the_step``."""
    flat = _clip_gate_flat(args)
    impl = adapter_cfg.get("lora", {}).get("implementation")
    cfg = {"method": f"lora_tta_{impl}" if method == "lora" else ("full_tta" if method == "full" else method), **adapter_cfg,
           "generation": {"num_cond_frames": args.num_cond_frames, "num_frames": args.num_frames,
                          "gen_start_frame": args.gen_start_frame, "num_inference_steps": args.num_inference_steps,
                          "guidance_scale": args.guidance_scale, "resolution": args.resolution},
           "seed": args.seed, "max_videos": args.max_videos, **flat,
           "tta_frames": frames, "synthetic": True, "model": args.model}
    rest = outside_the_step(args)
    rest["clip_gate"] = {k[len("clip_gate_"):]: v for k, v in flat.items()}
    return {**cfg, **rest}


# summary.json: the constant the reference writes under "method" and the hyper-parameters it repeats at the top
# (run_lora_tta.py:1277-1322, run_delta_a.py:905-, run_delta_b.py:918-955, run_delta_c.py:674-, run_norm_tune_tta.py:631-,
# run_film_tta.py:676-)
_SUMMARY_HEAD = {
    "lora": ("lora_tta", ("lora_rank", "lora_alpha", "learning_rate", "num_steps")),
    "full": ("full_tta", ("learning_rate", "num_steps")),                       # run_full_tta.py:865-868
    "delta_a": ("delta_a", ("delta_steps", "delta_lr")),
    "delta_b": ("delta_b", ("delta_target", "delta_target_blocks", "num_groups", "delta_steps", "delta_lr")),
    "delta_c": ("delta_c", ("delta_mode", "delta_steps", "delta_lr")),
    "norm_tune": ("norm_tune", ("norm_target", "norm_steps", "norm_lr")),
    "film": ("film_adapter", ("film_mode", "num_groups", "film_steps", "film_lr")),
}


def summary_record(method: str, args, results: List[Dict]) -> Dict:
    """``summary.json`` with the reference's keys: averages run over the successful videos and are 0 without any; the
    generation / CLIP-gate columns exist (0 / empty statistics) because the exporters index them.  ``num_success`` is an
    alias kept from earlier builds."""
    ok = [r for r in results if r.get("success", False)]
    mean = lambda key: (sum(r.get(key) or 0.0 for r in ok) / len(ok)) if ok else 0      # noqa: E731
    losses = [r["final_loss"] for r in ok if r.get("final_loss") is not None]
    name, head = _SUMMARY_HEAD[method]
    gated = [r for r in ok if r.get("clip_gate_enabled")]
    extra = {"total_params": getattr(args, "_total_params", None)} if method == "full" else {}     # run_full_tta.py:874
    return {"method": name, **{k: getattr(args, k) for k in head}, **extra,
            "num_cond_frames": args.num_cond_frames, "num_frames": args.num_frames, "gen_start_frame": args.gen_start_frame,
            "batch_videos": args.batch_videos, "retrieval_pool_dir": args.retrieval_pool_dir,
            "num_videos": len(results), "num_successful": len(ok), "num_failed": len(results) - len(ok),
            "num_success": len(ok),
            "avg_train_time": mean("train_time"), "avg_clip_gate_eval_time": mean("clip_gate_eval_time"),
            "avg_es_check_time": mean("es_check_time"), "avg_gen_time": mean("gen_time"),
            "avg_total_time": mean("total_time"), "avg_final_loss": (sum(losses) / len(losses)) if losses else 0,
            **_clip_gate_flat(args),
            "clip_gate_stats": {"num_enabled": len(gated), "num_scored": 0, "num_skipped": 0, "skip_rate": 0.0},
            "results": results}


def training_record(method: str, args, r: Dict) -> Dict:
    """Per-video fields that come out of the optimisation loop, with the method-specific ones the reference records
    (run_delta_a.py:763-775 ``delta_norm``, run_delta_b.py:757-768 ``delta_norms`` + ``num_groups``, run_delta_c.py:531-541
    ``delta_out_norm``, run_film_tta.py:584-593 ``correction_norm``); ``losses`` and ``num_train_steps`` on every method
    are additions."""
    rec = {"train_time": r["train_time"], "es_check_time": r.get("es_check_time", 0.0),
           "final_loss": r["losses"][-1] if r["losses"] else None, "num_train_steps": len(r["losses"]), "losses": r["losses"]}
    for key in ("delta_norm", "delta_norms", "delta_out_norm", "correction_norm"):
        if key in r:
            rec[key] = r[key]
    if method == "delta_b":
        rec["num_groups"] = args.num_groups
    rec.update({"early_stopping_info": r.get("early_stopping_info"), "success": True})
    return rec


def frame_budget(args, context: str = ""):
    """Resolved TTA window (run_lora_tta.py:743-758, guard included) and what it is in latent frames: the VAE keeps the
    first pixel frame and then one latent per 4 (run_lora_tta.py:1088-1093)."""
    resolve_tta_frames(args)
    validate_tta_feature_budget(args, context=context)
    total, ctx = args.tta_total_frames, args.tta_context_frames
    n_lat = 1 + (total - 1) // 4
    n_ctx_lat = 1 + (ctx - 1) // 4
    return total, ctx, n_lat, n_ctx_lat


def synthetic_video(idx: int, n_lat: int, hw, cfg, device):
    g = torch.Generator().manual_seed(1000 + idx)
    lat = torch.randn(1, 16, n_lat, hw[0], hw[1], generator=g).to(BF16).to(device)
    prompt = torch.randn(1, 1, 512, cfg.caption_channels, generator=g).to(BF16).to(device)
    mask = torch.zeros(1, 512, dtype=torch.int64)
    mask[:, :128] = 1
    return dict(video_name=f"synthetic_{idx:04d}", video_path="", caption="synthetic", latents=lat, prompt_embeds=prompt,
                prompt_mask=mask.to(device))


def _save_json(path: Path, obj):
    tmp = path.with_suffix(path.suffix + ".tmp")
    tmp.write_text(json.dumps(obj, indent=2, default=float))
    os.replace(tmp, path)


def run(method: str, argv=None) -> Dict:
    args = build_parser(method).parse_args(argv)
    out = Path(args.output_dir)
    out.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(args.seed)
    device = args.device
    # Under torchrun the step runs data-parallel over noise draws (dist.py): one process per GPU, every rank
    # the same videos and adapters, its own (sigma, eps) stream, gradients all-reduced inside TTAStepper; rank 0 alone
    # writes the output files.  Without torchrun: world = 1 and nothing below differs from the single-GPU run.
    world = D.init_from_env()
    rank = D.rank()
    if world > 1 and str(device).startswith("cuda"):
        device = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
        torch.cuda.set_device(device)
    save_json = _save_json if rank == 0 else (lambda path, obj: None)
    try:
        import longcat_video  # noqa: F401  (upstream package: VAE / text encoder / pipeline)
        have_upstream = True
    except Exception:
        have_upstream = False
    if not args.synthetic and not have_upstream:
        print("[b200tta] upstream LongCat-Video package not importable -> running on synthetic latents (--synthetic)")
        args.synthetic = True
    if not args.synthetic:
        raise NotImplementedError(
            "real-video mode needs the upstream VAE / UMT5 / pipeline (out of scope here): encode with the reference's "
            "common.load_longcat_components / encode_video / encode_prompt, load the DiT weights into B200DiT "
            "(INTEGRATION.md 1a) and call longcat_video_tta_b200.lora.finetune_lora_on_conditioning")

    if args.batch_videos > 1 and method not in ("lora", "full", "delta_a"):
        raise NotImplementedError("--batch-videos > 1 exists for LoRA (finetune_lora_batch), full-model TTA "
                                  "(finetune_full_batch) and delta-A (_optimize_delta_a_batch) only, as in the reference")

    total, ctx, n_lat, n_ctx_lat = frame_budget(args, context={"lora": "lora_tta", "full": "full_tta"}.get(method, method))
    dit = B200DiT.random_init(args.model, seed=0, device=device)
    cfg = dit.config
    hw = tuple(int(x) for x in args.latent_hw.split(",")) if args.latent_hw else (60, 104)

    # ---- adapters
    mods = wrapper = norm_params = None
    adapter_cfg: Dict = {}
    if method == "lora":
        tm = [m.strip() for m in args.target_modules.split(",")]
        inject = L.inject_builtin_lora_into_dit if args.use_builtin_lora else L.inject_lora_into_dit
        kw = dict(rank=args.lora_rank, alpha=args.lora_alpha, target_modules=tm, target_ffn=args.target_ffn,
                  target_blocks=args.lora_target_blocks)
        if not args.use_builtin_lora:
            kw["dropout"] = args.lora_dropout
        mods = inject(dit, **kw)
        counts = L.count_lora_parameters(mods)
        adapter_cfg = {"lora": {"implementation": "builtin" if args.use_builtin_lora else "custom", "rank": args.lora_rank,
                                "alpha": args.lora_alpha, "dropout": args.lora_dropout, "target_modules": tm,
                                "target_blocks": args.lora_target_blocks, "target_ffn": args.target_ffn,
                                "num_modules": len(mods), "trainable_params": counts["trainable"]},
                       "training": {"learning_rate": args.learning_rate, "num_steps": args.num_steps,
                                    "warmup_steps": args.warmup_steps, "weight_decay": args.weight_decay,
                                    "max_grad_norm": args.max_grad_norm}}
    elif method == "full":
        # run_full_tta.py:449-462: every DiT parameter trains; the base state is kept on the host and restored per video
        for p in dit.parameters():
            p.requires_grad = True
        total_params = sum(p.numel() for p in dit.parameters())
        args._total_params = total_params
        base_state = {k: v.detach().cpu().clone() for k, v in dit.state_dict().items()}
        adapter_cfg = {"training": {"learning_rate": args.learning_rate, "num_steps": args.num_steps,
                                    "warmup_steps": args.warmup_steps, "weight_decay": args.weight_decay,
                                    "max_grad_norm": args.max_grad_norm, "optimizer": args.optimizer,
                                    "total_params": total_params, "trainable_params": total_params}}
    elif method == "delta_a":
        wrapper = A.DeltaAWrapper(dit, cfg.adaln_tembed_dim)
    elif method == "delta_b":
        wrapper = A.DeltaBWrapper(dit, num_groups=args.num_groups, adaln_tembed_dim=cfg.adaln_tembed_dim,
                                  hidden_size=cfg.hidden_size, delta_target=args.delta_target,
                                  delta_dim=args.delta_dim if args.delta_dim else (cfg.hidden_size if args.delta_target == "hidden" else None),
                                  target_blocks=args.delta_target_blocks)
    elif method == "delta_c":
        wrapper = A.DeltaCWrapper(dit, mode=args.delta_mode, out_channels=cfg.out_channels)
    elif method == "norm_tune":
        norm_params = A.collect_norm_params(dit, args.norm_target)
        for p in norm_params:
            p.requires_grad_(True)
        wrapper = A.NormTuneForward(dit, also_tune_delta=args.also_tune_delta, adaln_tembed_dim=cfg.adaln_tembed_dim)
        if wrapper.delta is not None:
            norm_params.append(wrapper.delta)       # run_norm_tune_tta.py:382-385: last in the optimizer's list
    elif method == "film":
        wrapper = A.FiLMAdapterWrapper(dit, num_groups=args.num_groups, hidden_size=cfg.hidden_size, film_mode=args.film_mode)
        wrapper.apply_to_dit()
    if method not in ("lora", "full"):
        n_train = sum(p.numel() for p in wrapper.trainable())
        adapter_cfg = {method: {k: v for k, v in vars(args).items() if k.startswith(("delta", "norm", "film", "num_groups", "also_tune"))},
                       "trainable_params": n_train}

    save_json(out / "config.json", experiment_config(method, args, adapter_cfg, {
        "total": total, "context": ctx, "latent_frames": n_lat, "context_latents": n_ctx_lat}))
    ckpt_path = out / "checkpoint.json"
    state = {"next_idx": 0, "results": []}
    if ckpt_path.exists() and not args.restart:
        state = json.loads(ckpt_path.read_text())
    early_stopper = build_early_stopper_from_args(args)
    init_norm = A.snapshot_params(norm_params) if norm_params else None

    for idx in range(state["next_idx"], args.max_videos):
        t_video = time.time()
        vid = synthetic_video(idx, n_lat, hw, cfg, device)
        result = {"idx": idx, "video_name": vid["video_name"], "video_path": vid["video_path"], "caption": vid["caption"],
                  "batch_size": 1, "num_neighbors": 0}
        try:
            cond, train, val = split_tta_latents(vid["latents"], n_ctx_lat, args.es_holdout_fraction)
            model = dit if method in ("lora", "full") else wrapper
            if world > 1:       # replicas re-initialise identically ...
                torch.manual_seed(args.seed + 1_000_003 * (idx + 1))
            # reset the adapter for every video (run_lora_tta.py:1127)
            if method == "lora":
                (L.reset_builtin_lora_weights if args.use_builtin_lora else L.reset_lora_weights)(mods)
                dit.engine.resolve_sites()
            elif method == "full":
                F.reset_dit_weights(dit, base_state)            # run_full_tta.py: per-video reset from the host copy
            elif method == "norm_tune":
                A.restore_params(norm_params, init_norm)
            else:
                for p in wrapper.trainable():
                    p.data.zero_()
            if world > 1:       # ... and then draw from their own stream (world = 1 keeps the reference's single stream)
                torch.manual_seed(D.draw_seed(args.seed + 1_000_003 * (idx + 1) + 7919, rank))
            batch_mode = args.batch_videos > 1      # the reference's batch loops run without the early stopper
            es = early_stopper if (early_stopper is not None and val is not None and not batch_mode) else None
            if es is not None:
                if method == "full":    # run_full_tta.py:722-741: no save_fn, the stopper's default snapshot covers the model
                    es.setup(model, cond, val, vid["prompt_embeds"], vid["prompt_mask"], device=device, dtype=BF16,
                             video_id=vid["video_name"])
                else:
                    save_fn = (lambda: [p.data.clone() for p in (L.get_lora_parameters(mods) if method == "lora" else wrapper.trainable())])
                    es.setup(model, cond, val, vid["prompt_embeds"], vid["prompt_mask"], device=device, dtype=BF16,
                             video_id=vid["video_name"], save_fn=save_fn)
            if batch_mode:
                # retrieval-augmented batch (run_lora_tta.py:1037-1062, run_delta_a.py:~640 build it from the pool; here:
                # the evaluation video + K-1 further synthetic videos), held on the host and visited round-robin
                batch = []
                for j in range(args.batch_videos):
                    nb = vid if j == 0 else synthetic_video(10_000 * j + idx, n_lat, hw, cfg, "cpu")
                    c_j, t_j, _ = split_tta_latents(nb["latents"], n_ctx_lat, args.es_holdout_fraction)
                    batch.append({"cond_latents": c_j.cpu(), "train_latents": t_j.cpu(),
                                  "prompt_embeds": nb["prompt_embeds"].cpu(), "prompt_mask": nb["prompt_mask"].cpu()})
                result.update({"batch_size": len(batch), "num_neighbors": len(batch) - 1})
            if batch_mode and method == "lora":
                r = L.finetune_lora_batch(dit, mods, batch, num_steps=args.num_steps, lr=args.learning_rate,
                                          warmup_steps=args.warmup_steps, weight_decay=args.weight_decay,
                                          max_grad_norm=args.max_grad_norm, device=device, dtype=BF16)
            elif batch_mode and method == "full":
                r = F.finetune_full_batch(dit, batch, num_steps=args.num_steps, lr=args.learning_rate,
                                          warmup_steps=args.warmup_steps, weight_decay=args.weight_decay,
                                          max_grad_norm=args.max_grad_norm, device=device, dtype=BF16,
                                          optimizer_type=args.optimizer)
            elif batch_mode:
                t0 = time.time()
                r = A._optimize_delta_a_batch(wrapper, batch, num_steps=args.delta_steps, lr=args.delta_lr, device=device)
                r["train_time"] = time.time() - t0
            elif method == "full":
                r = F.finetune_full_on_conditioning(dit, cond, train, vid["prompt_embeds"], vid["prompt_mask"],
                                                    num_steps=args.num_steps, lr=args.learning_rate,
                                                    warmup_steps=args.warmup_steps, weight_decay=args.weight_decay,
                                                    max_grad_norm=args.max_grad_norm, device=device, dtype=BF16,
                                                    early_stopper=es, optimizer_type=args.optimizer)
            elif method == "lora":
                variants = None
                if args.aug_enabled and args.aug_flip:
                    # the reference flips pixel frames before the VAE (common.py augmentation helpers); on synthetic
                    # latents the mirrored latent stands in so that the variant draw of the loop is exercised
                    variants = [{"latents": train, "name": "orig"}, {"latents": torch.flip(train, dims=[-1]), "name": "flip"}]
                r = L.finetune_lora_on_conditioning(dit, mods, cond, train, vid["prompt_embeds"], vid["prompt_mask"],
                                                    num_steps=args.num_steps, lr=args.learning_rate,
                                                    warmup_steps=args.warmup_steps, weight_decay=args.weight_decay,
                                                    max_grad_norm=args.max_grad_norm, device=device, dtype=BF16,
                                                    early_stopper=es, train_latents_variants=variants)
            else:
                fn = {"delta_a": A.optimize_delta_a, "delta_b": A.optimize_delta_b, "delta_c": A.optimize_delta_c,
                      "film": A.optimize_film_adapter}.get(method)
                steps = getattr(args, {"norm_tune": "norm_steps", "film": "film_steps"}.get(method, "delta_steps"))
                lr = getattr(args, {"norm_tune": "norm_lr", "film": "film_lr"}.get(method, "delta_lr"))
                t0 = time.time()
                if method == "norm_tune":
                    r = A.optimize_norm_params(wrapper, norm_params, cond, train, vid["prompt_embeds"], vid["prompt_mask"],
                                               num_steps=steps, lr=lr, device=device, early_stopper=es)
                else:
                    r = fn(wrapper, cond, train, vid["prompt_embeds"], vid["prompt_mask"], num_steps=steps, lr=lr,
                           device=device, early_stopper=es)
                r.setdefault("train_time", time.time() - t0)
                r.setdefault("es_check_time", 0.0)
            if method == "lora" and args.save_lora_weights and not args.use_builtin_lora and rank == 0:
                (out / "lora_weights").mkdir(exist_ok=True)                  # run_lora_tta.py:1250-1252
                L.save_lora_weights(mods, str(out / "lora_weights" / f"{vid['video_name']}_lora.pt"))
            result.update(training_record(method, args, r))
        except Exception as e:  # per-video failure is recorded and the run continues (run_lora_tta.py:1264-1271)
            if world > 1:       # ... unless other ranks are waiting in a collective: fail the whole job loudly
                raise
            result.update({"success": False, "error": f"{type(e).__name__}: {e}"})
        result["total_time"] = time.time() - t_video
        state["results"].append(result)
        state["next_idx"] = idx + 1
        save_json(ckpt_path, state)

    summary = summary_record(method, args, state["results"])
    save_json(out / "summary.json", summary)
    if world > 1:
        torch.distributed.barrier()
    return summary
