"""Command-line entry points with the reference's flags and output layout:

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
from . import lora as L
from .common import split_tta_latents
from .dit import B200DiT
from .early_stopping import add_early_stopping_args, build_early_stopper_from_args

BF16 = torch.bfloat16
METHODS = ("lora", "delta_a", "delta_b", "delta_c", "norm_tune", "film")


def _common_args(p: argparse.ArgumentParser):
    p.add_argument("--checkpoint-dir", type=str, default=None)
    p.add_argument("--data-dir", type=str, default=None)
    p.add_argument("--output-dir", type=str, required=True)
    p.add_argument("--max-videos", type=int, default=2)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--device", type=str, default="cuda")
    p.add_argument("--restart", action="store_true")
    p.add_argument("--num-cond-frames", type=int, default=14)
    p.add_argument("--tta-total-frames", type=int, default=None)
    p.add_argument("--tta-context-frames", type=int, default=None)
    p.add_argument("--num-frames", type=int, default=28)
    p.add_argument("--gen-start-frame", type=int, default=14)
    p.add_argument("--num-inference-steps", type=int, default=50)
    p.add_argument("--guidance-scale", type=float, default=4.0)
    p.add_argument("--resolution", type=str, default="480p")
    p.add_argument("--skip-generation", action="store_true")
    p.add_argument("--no-save-videos", action="store_true")
    p.add_argument("--synthetic", action="store_true", help="seeded synthetic latents + random-init DiT")
    p.add_argument("--model", type=str, default="13.6b", choices=["13.6b", "tiny"])
    p.add_argument("--latent-hw", type=str, default=None, help="synthetic latent H,W (default 60,104 = 480x832)")
    add_early_stopping_args(p)


def build_parser(method: str) -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description=f"B200-native {method} TTA for LongCat-Video")
    _common_args(p)
    if method == "lora":
        p.add_argument("--lora-rank", type=int, default=8)
        p.add_argument("--lora-alpha", type=float, default=16.0)
        p.add_argument("--lora-dropout", type=float, default=0.0)
        p.add_argument("--target-ffn", action="store_true")
        p.add_argument("--target-modules", type=str, default="qkv,proj")
        p.add_argument("--lora-target-blocks", type=str, default="all")
        p.add_argument("--use-builtin-lora", action="store_true")
        p.add_argument("--save-lora-weights", action="store_true")
        p.add_argument("--learning-rate", type=float, default=2e-4)
        p.add_argument("--num-steps", type=int, default=20)
        p.add_argument("--warmup-steps", type=int, default=3)
        p.add_argument("--weight-decay", type=float, default=0.01)
        p.add_argument("--max-grad-norm", type=float, default=1.0)
    elif method in ("delta_a", "delta_b", "delta_c"):
        p.add_argument("--delta-steps", type=int, default=20)
        p.add_argument("--delta-lr", type=float, default=1e-3)
        if method == "delta_b":
            p.add_argument("--num-groups", type=int, default=4)
            p.add_argument("--delta-target", type=str, default="timestep", choices=["timestep", "hidden"])
            p.add_argument("--delta-dim", type=int, default=None)
            p.add_argument("--delta-target-blocks", type=str, default="all")
        if method == "delta_c":
            p.add_argument("--delta-mode", type=str, default="per_channel")
    elif method == "norm_tune":
        p.add_argument("--norm-steps", type=int, default=20)
        p.add_argument("--norm-lr", type=float, default=1e-4)
        p.add_argument("--norm-target", type=str, default="cross_attn_norm", choices=["cross_attn_norm", "qk_norm", "all_norm"])
    elif method == "film":
        p.add_argument("--film-steps", type=int, default=20)
        p.add_argument("--film-lr", type=float, default=1e-3)
        p.add_argument("--film-mode", type=str, default="full", choices=["full", "shift_scale", "scale_only"])
        p.add_argument("--num-groups", type=int, default=4)
    return p


def frame_budget(args):
    """run_lora_tta.py:743-758: latent frames used for TTA and the context split."""
    total = args.tta_total_frames or args.num_cond_frames
    ctx = args.tta_context_frames or total
    ctx = min(ctx, total)
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

    dit = B200DiT.random_init(args.model, seed=0, device=device)
    cfg = dit.config
    hw = tuple(int(x) for x in args.latent_hw.split(",")) if args.latent_hw else (60, 104)
    total, ctx, n_lat, n_ctx_lat = frame_budget(args)

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
        wrapper = A.NormTuneForward(dit)
    elif method == "film":
        wrapper = A.FiLMAdapterWrapper(dit, num_groups=args.num_groups, hidden_size=cfg.hidden_size, film_mode=args.film_mode)
        wrapper.apply_to_dit()
    if method != "lora":
        n_train = sum(p.numel() for p in wrapper.trainable())
        adapter_cfg = {method: {k: v for k, v in vars(args).items() if k.startswith(("delta", "norm", "film", "num_groups"))},
                       "trainable_params": n_train}

    _save_json(out / "config.json", {"method": method, **adapter_cfg,
                                     "generation": {"num_cond_frames": args.num_cond_frames, "num_frames": args.num_frames,
                                                    "gen_start_frame": args.gen_start_frame,
                                                    "num_inference_steps": args.num_inference_steps,
                                                    "guidance_scale": args.guidance_scale, "resolution": args.resolution},
                                     "tta_frames": {"total": total, "context": ctx, "latent_frames": n_lat,
                                                    "context_latents": n_ctx_lat},
                                     "seed": args.seed, "max_videos": args.max_videos, "synthetic": True, "model": args.model})
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
            model = dit if method == "lora" else wrapper
            # reset the adapter for every video (run_lora_tta.py:1127)
            if method == "lora":
                (L.reset_builtin_lora_weights if args.use_builtin_lora else L.reset_lora_weights)(mods)
                dit.engine.resolve_sites()
            elif method == "norm_tune":
                A.restore_params(norm_params, init_norm)
            else:
                for p in wrapper.trainable():
                    p.data.zero_()
            es = early_stopper if (early_stopper is not None and val is not None) else None
            if es is not None:
                save_fn = (lambda: [p.data.clone() for p in (L.get_lora_parameters(mods) if method == "lora" else wrapper.trainable())])
                es.setup(model, cond, val, vid["prompt_embeds"], vid["prompt_mask"], device=device, dtype=BF16,
                         video_id=vid["video_name"], save_fn=save_fn)
            if method == "lora":
                r = L.finetune_lora_on_conditioning(dit, mods, cond, train, vid["prompt_embeds"], vid["prompt_mask"],
                                                    num_steps=args.num_steps, lr=args.learning_rate,
                                                    warmup_steps=args.warmup_steps, weight_decay=args.weight_decay,
                                                    max_grad_norm=args.max_grad_norm, device=device, dtype=BF16, early_stopper=es)
                if args.save_lora_weights and not args.use_builtin_lora:
                    (out / "lora_weights").mkdir(exist_ok=True)
                    L.save_lora_weights(mods, str(out / "lora_weights" / f"{vid['video_name']}_lora.pt"))
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
            result.update({"train_time": r["train_time"], "es_check_time": r.get("es_check_time", 0.0),
                           "final_loss": r["losses"][-1] if r["losses"] else None, "num_train_steps": len(r["losses"]),
                           "losses": r["losses"], "early_stopping_info": r.get("early_stopping_info"), "success": True})
        except Exception as e:  # per-video failure is recorded and the run continues (run_lora_tta.py:1264-1271)
            result.update({"success": False, "error": f"{type(e).__name__}: {e}"})
        result["total_time"] = time.time() - t_video
        state["results"].append(result)
        state["next_idx"] = idx + 1
        _save_json(ckpt_path, state)

    ok = [r for r in state["results"] if r.get("success")]
    summary = {"method": method, "num_videos": len(state["results"]), "num_success": len(ok),
               "avg_train_time": (sum(r["train_time"] for r in ok) / len(ok)) if ok else None,
               "avg_final_loss": (sum(r["final_loss"] for r in ok) / len(ok)) if ok else None,
               "results": state["results"]}
    _save_json(out / "summary.json", summary)
    return summary
