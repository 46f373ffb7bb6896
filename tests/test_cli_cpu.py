"""CPU: the drop-in scripts keep the reference's flags and frame-budget arithmetic."""
import pytest

from longcat_video_tta_b200 import cli, lora
from longcat_video_tta_b200.common import split_tta_latents


def test_lora_flags_match_reference():
    a = cli.build_parser("lora").parse_args(
        "--output-dir /tmp/x --lora-rank 16 --lora-alpha 32 --target-modules qkv,proj --lora-target-blocks last_4 "
        "--learning-rate 2e-4 --num-steps 20 --warmup-steps 3 --weight-decay 0.01 --max-grad-norm 1.0 --target-ffn "
        "--use-builtin-lora --save-lora-weights --num-cond-frames 14 --es-check-every 1 --es-patience 2 "
        "--es-anchor-sigmas 0.25,0.5,0.75 --es-noise-draws 2 --es-strategy patience --es-holdout-fraction 0.25 "
        "--skip-generation --seed 42 --max-videos 3".split())
    assert a.lora_rank == 16 and a.lora_alpha == 32 and a.target_ffn and a.use_builtin_lora and a.es_check_every == 1


@pytest.mark.parametrize("method,flags", [
    ("delta_a", "--delta-steps 5 --delta-lr 1e-3"),
    ("delta_b", "--delta-steps 5 --delta-lr 1e-3 --num-groups 4 --delta-target hidden --delta-dim 4096 --delta-target-blocks all"),
    ("delta_c", "--delta-steps 5 --delta-lr 1e-3 --delta-mode per_channel"),
    ("norm_tune", "--norm-steps 20 --norm-lr 1e-4 --norm-target all_norm"),
    ("film", "--film-steps 20 --film-lr 1e-3 --film-mode shift_scale --num-groups 4"),
])
def test_adapter_flags(method, flags):
    cli.build_parser(method).parse_args(("--output-dir /tmp/x " + flags).split())


def test_frame_budget_and_split_match_survey_3_1():
    a = cli.build_parser("lora").parse_args("--output-dir /tmp/x --num-cond-frames 14 --feature-frame-guard-mode warn".split())
    total, ctx, n_lat, n_ctx = cli.frame_budget(a)
    assert (total, ctx, n_lat, n_ctx) == (14, 14, 4, 4)
    import torch
    c, t, v = split_tta_latents(torch.zeros(1, 1, n_lat, 1, 1), n_ctx)
    assert (c.shape[2], t.shape[2], v) == (3, 1, None)      # early stopping silently inactive
    a = cli.build_parser("lora").parse_args("--output-dir /tmp/x --tta-total-frames 117 --tta-context-frames 13 --gen-start-frame 117".split())
    total, ctx, n_lat, n_ctx = cli.frame_budget(a)
    c, t, v = split_tta_latents(torch.zeros(1, 1, n_lat, 1, 1), n_ctx)
    assert (n_lat, n_ctx) == (30, 4) and (c.shape[2], t.shape[2], v.shape[2]) == (4, 20, 6)


def test_parse_target_blocks():
    assert lora._parse_target_blocks("all", 48) is None
    assert lora._parse_target_blocks("last_4", 48) == {44, 45, 46, 47}
    assert lora._parse_target_blocks("0, 5,10", 48) == {0, 5, 10}
    with pytest.raises(ValueError):
        lora._parse_target_blocks("48", 48)
    with pytest.raises(ValueError):
        lora._parse_target_blocks("last_49", 48)


def test_split_table_matches_reference_golden(golden_dir):
    import torch
    table = torch.load(golden_dir / "split_table.pt")
    for (T, ctx, hf), want in table.items():
        c, t, v = split_tta_latents(torch.zeros(1, 1, T, 1, 1), ctx, hf)
        assert (c.shape[2], t.shape[2], 0 if v is None else v.shape[2]) == want


def test_split_budget_matches_reference_golden(golden_dir):
    """common.py:1493-1517 over 4 140 (total, context, hold-out, VAE scale) settings recorded from the reference, and its
    agreement with what split_tta_latents actually produces."""
    import torch
    from longcat_video_tta_b200.common import estimate_tta_split_budget
    table = torch.load(golden_dir / "split_budget.pt", weights_only=False)
    assert len(table) == 4140
    for (total, ctx, hf, scale), want in table.items():
        got = estimate_tta_split_budget(total, ctx, holdout_fraction=hf, vae_t_scale=scale)
        assert got == want, (total, ctx, hf, scale, got, want)
        if got["total_latents"] >= 2:
            n_ctx_lat = 1 + (max(1, ctx) - 1) // scale
            c, t, v = split_tta_latents(torch.zeros(1, 1, got["total_latents"], 1, 1), n_ctx_lat, hf)
            assert (c.shape[2], t.shape[2], 0 if v is None else v.shape[2]) == \
                (got["cond_latents"], got["train_latents"], got["val_latents"])


def test_parse_target_blocks_matches_reference_golden(golden_dir):
    """run_lora_tta.py:263-283 (and its copy in run_delta_b.py) over 72 (spec, depth) pairs, invalid specs included."""
    import torch
    from longcat_video_tta_b200.adapters import _parse_target_blocks as parse_adapters
    table = torch.load(golden_dir / "target_blocks.pt", weights_only=False)
    for fn, key in ((lora._parse_target_blocks, "lora"), (parse_adapters, "delta_b")):
        for (spec, n), want in table[key].items():
            try:
                r = fn(spec, n)
                got = None if r is None else sorted(r)
            except Exception as e:  # noqa: BLE001
                got = f"raises {type(e).__name__}"
            assert got == want, (key, spec, n, got, want)


@pytest.mark.parametrize("method", ["lora", "full", "delta_a", "delta_b", "delta_c", "norm_tune", "film"])
def test_flag_surface_matches_reference_golden(golden_dir, method):
    """Every flag the reference's own parser declares (oracle/make_golden_cli_flags.py executes the parser section of
    each script's main()) exists here with the same dest, type, default, choices, nargs, const and action class.  The one
    stated deviation: --checkpoint-dir / --data-dir are required there and checked after parsing here (--synthetic)."""
    import json
    golden = json.loads((golden_dir / "cli_flags.json").read_text())[method]
    ours = {}
    for a in cli.build_parser(method)._actions:
        for f in a.option_strings:
            ours[f] = a
    assert len(golden) >= 48
    for g in golden:
        for f in g["flags"]:
            assert f in ours, f"{method}: reference flag {f} missing"
            a = ours[f]
            got = {"dest": a.dest, "default": a.default, "type": getattr(a.type, "__name__", None) if a.type else None,
                   "choices": list(a.choices) if a.choices is not None else None, "nargs": a.nargs, "const": a.const,
                   "action": type(a).__name__, "required": bool(a.required)}
            want = {k: g[k] for k in got}
            if f in ("--checkpoint-dir", "--data-dir"):
                want["required"] = False
            assert got == want, (method, f, got, want)


def test_flags_outside_the_step_are_recorded():
    a = cli.build_parser("delta_b").parse_args(
        "--output-dir /tmp/x --clip-gate-enabled --clip-gate-fail-closed --caption-guard-mode warn --compute-fvd "
        "--aug-enabled --no-aug-rotate-zoom --batch-videos 4 --fixed-caption hello".split())
    rec = cli.outside_the_step(a)
    assert rec["clip_gate"]["clip_gate_enabled"] and rec["clip_gate"]["clip_gate_fail_open"] is False
    assert rec["caption_guard"]["caption_guard_mode"] == "warn" and rec["caption_guard"]["fixed_caption"] == "hello"
    assert rec["online_eval"]["compute_fvd"] and rec["augmentation"]["aug_rotate_zoom"] is False
    assert rec["batch"]["batch_videos"] == 4


def _stub_engine(monkeypatch, calls):
    """Replace the GPU pieces of cli.run with recorders so the host control flow (batch assembly, variants, records)
    runs on the CPU."""
    import types
    import torch
    cfg = types.SimpleNamespace(caption_channels=32, adaln_tembed_dim=512, hidden_size=64, out_channels=16)
    dit = types.SimpleNamespace(config=cfg, engine=types.SimpleNamespace(resolve_sites=lambda: None))
    monkeypatch.setattr(cli.B200DiT, "random_init", staticmethod(lambda *a, **k: dit))
    monkeypatch.setattr(cli.L, "inject_lora_into_dit", lambda d, **kw: ["m0", "m1"])
    monkeypatch.setattr(cli.L, "count_lora_parameters", lambda mods: {"trainable": 7})
    monkeypatch.setattr(cli.L, "reset_lora_weights", lambda mods: None)

    def single(d, mods, cond, train, pe, pm, **kw):
        calls.append(("single", cond.shape[2], train.shape[2], kw.get("train_latents_variants")))
        return {"losses": [1.0, 0.5], "train_time": 0.1, "es_check_time": 0.0, "early_stopping_info": None}

    def batch(d, mods, batch_data, **kw):
        calls.append(("batch", batch_data, kw))
        return {"losses": [1.0] * kw["num_steps"], "train_time": 0.1, "es_check_time": 0.0, "early_stopping_info": None}

    monkeypatch.setattr(cli.L, "finetune_lora_on_conditioning", single)
    monkeypatch.setattr(cli.L, "finetune_lora_batch", batch)
    return torch


def test_run_lora_batch_and_flip_variant_host_flow(tmp_path, monkeypatch, golden_dir):
    import json
    calls = []
    torch = _stub_engine(monkeypatch, calls)
    base = (f"--synthetic --model tiny --device cpu --latent-hw 8,8 --tta-total-frames 17 --tta-context-frames 5 "
            f"--max-videos 1 --es-disable --num-steps 4")
    s = cli.run("lora", (f"--output-dir {tmp_path / 'b'} {base} --batch-videos 3").split())
    kind, data, kw = calls[-1]
    assert kind == "batch" and len(data) == 3 and kw["num_steps"] == 4
    assert all(d["cond_latents"].device.type == "cpu" and d["cond_latents"].shape[2] == 2 for d in data)
    assert not torch.equal(data[0]["train_latents"], data[1]["train_latents"])      # neighbours are other videos
    r = s["results"][0]
    assert r["success"] and r["batch_size"] == 3 and r["num_neighbors"] == 2 and r["num_train_steps"] == 4
    cfg = json.loads((tmp_path / "b" / "config.json").read_text())
    assert cfg["batch"]["batch_videos"] == 3 and cfg["clip_gate"]["fail_open"] is True and cfg["method"] == "lora_tta_custom"
    layout = json.loads((golden_dir / "output_layout.json").read_text())["lora"]
    _covers(cfg, layout["config"])                      # the file as written by run(), not only the builder
    written = json.loads((tmp_path / "b" / "summary.json").read_text())
    assert [k for k in layout["summary_keys"] if k not in written] == [] and written["num_successful"] == 1
    ck = json.loads((tmp_path / "b" / "checkpoint.json").read_text())
    assert set(ck) == {"next_idx", "results"} and ck["next_idx"] == 1

    s = cli.run("lora", (f"--output-dir {tmp_path / 'v'} {base} --aug-enabled --aug-flip").split())
    kind, t_cond, t_train, variants = calls[-1]
    assert kind == "single" and [v["name"] for v in variants] == ["orig", "flip"]
    assert torch.equal(variants[1]["latents"], torch.flip(variants[0]["latents"], dims=[-1]))
    s = cli.run("lora", (f"--output-dir {tmp_path / 'p'} {base}").split())
    assert calls[-1][3] is None and s["results"][0]["batch_size"] == 1


def test_run_full_host_flow(tmp_path, monkeypatch, golden_dir):
    """lora_experiment/scripts/run_full_tta.py through cli.run with the GPU pieces stubbed: every parameter is made trainable,
    the base state is restored before every video (run_full_tta.py:455-462), single and batch loops get the reference's
    arguments, the files carry the reference's keys"""
    import json
    import types
    import torch
    lin = torch.nn.Linear(4, 3)
    for p in lin.parameters():
        p.requires_grad = False
    dit = types.SimpleNamespace(config=types.SimpleNamespace(caption_channels=32, adaln_tembed_dim=512, hidden_size=64, out_channels=16),
                                parameters=lin.parameters, state_dict=lin.state_dict)
    monkeypatch.setattr(cli.B200DiT, "random_init", staticmethod(lambda *a, **k: dit))
    calls, resets = [], []

    def reset(d, base):
        resets.append({k: v.clone() for k, v in base.items()})

    def single(d, cond, train, pe, pm, **kw):
        assert all(p.requires_grad for p in lin.parameters())
        with torch.no_grad():
            lin.weight.add_(1.0)          # "training" changes the model; the next video must start from the base state
        calls.append(("single", cond.shape[2], train.shape[2], kw))
        return {"losses": [2.0, 1.0], "train_time": 0.2, "es_check_time": 0.0, "early_stopping_info": None}

    def batch(d, batch_data, **kw):
        calls.append(("batch", batch_data, kw))
        return {"losses": [1.0] * kw["num_steps"], "train_time": 0.2, "es_check_time": 0.0, "early_stopping_info": None}

    monkeypatch.setattr(cli.F, "reset_dit_weights", reset)
    monkeypatch.setattr(cli.F, "finetune_full_on_conditioning", single)
    monkeypatch.setattr(cli.F, "finetune_full_batch", batch)
    base_w = lin.weight.detach().clone()
    base = ("--synthetic --model tiny --device cpu --latent-hw 8,8 --tta-total-frames 17 --tta-context-frames 5 "
            "--max-videos 2 --es-disable --num-steps 3 --optimizer adamw --learning-rate 2e-5")
    s = cli.run("full", f"--output-dir {tmp_path / 'f'} {base}".split())
    assert [c[0] for c in calls] == ["single", "single"] and calls[0][1:3] == (2, 2)
    kw = calls[0][3]
    assert (kw["num_steps"], kw["lr"], kw["warmup_steps"], kw["optimizer_type"], kw["early_stopper"]) == (3, 2e-5, 2, "adamw", None)
    assert len(resets) == 2 and all(torch.equal(r["weight"], base_w) for r in resets)      # the host copy of the BASE state
    assert s["method"] == "full_tta" and s["total_params"] == 15 and s["num_successful"] == 2
    layout = json.loads((golden_dir / "output_layout.json").read_text())["full"]
    cfg = json.loads((tmp_path / "f" / "config.json").read_text())
    _covers(cfg, layout["config"])
    assert cfg["method"] == "full_tta" and cfg["training"]["trainable_params"] == 15 and cfg["training"]["optimizer"] == "adamw"
    written = json.loads((tmp_path / "f" / "summary.json").read_text())
    assert [k for k in layout["summary_keys"] if k not in written] == []
    rec = written["results"][0]
    want = set(layout["result_keys"]) | (set(layout["result_keys_later"]) - {"gen_time", "output_path"})
    assert want <= set(rec), want - set(rec)
    s = cli.run("full", f"--output-dir {tmp_path / 'fb'} {base} --batch-videos 3".split())
    kind, data, kw = calls[-1]
    assert kind == "batch" and len(data) == 3 and kw["optimizer_type"] == "adamw" and s["results"][0]["batch_size"] == 3


@pytest.mark.parametrize("method,flags", [("delta_b", "--batch-videos 4"), ("film", "--batch-videos 2")])
def test_unbuilt_combinations_fail_loudly(tmp_path, method, flags):
    with pytest.raises(NotImplementedError):
        cli.run(method, f"--output-dir {tmp_path} --synthetic --model tiny --device cpu {flags}".split())


def _covers(have, want, path=""):
    for k, sub in want.items():
        assert k in have, f"missing {path}{k}"
        if isinstance(sub, dict):
            _covers(have[k], sub, f"{path}{k}.")


@pytest.mark.parametrize("method", ["lora", "full", "delta_a", "delta_b", "delta_c", "norm_tune", "film"])
def test_output_files_carry_the_reference_keys(golden_dir, method):
    """config.json / summary.json hold every key of the reference's own dict literals (oracle/make_golden_output_layout.py
    reads them out of each main(); four scripts are cut short in the snapshot, their keys up to the cut are checked) and
    the same constant under "method"."""
    import json
    layout = json.loads((golden_dir / "output_layout.json").read_text())[method]
    args = cli.build_parser(method).parse_args(["--output-dir", "/tmp/x", "--clip-gate-late-only"])
    results = [{"idx": 0, "success": True, "train_time": 2.0, "es_check_time": 1.0, "final_loss": 0.5, "total_time": 4.0},
               {"idx": 1, "success": True, "train_time": 4.0, "es_check_time": 0.0, "final_loss": 1.5, "total_time": 6.0},
               {"idx": 2, "success": False, "error": "x"}]
    s = cli.summary_record(method, args, results)
    assert s["method"] == layout["summary_method"]
    assert [k for k in layout["summary_keys"] if k not in s] == []
    assert (s["num_videos"], s["num_successful"], s["num_failed"]) == (3, 2, 1)
    assert (s["avg_train_time"], s["avg_es_check_time"], s["avg_total_time"], s["avg_final_loss"], s["avg_gen_time"]) == \
        (3.0, 0.5, 5.0, 1.0, 0.0)
    assert s["clip_gate_sampling_mode"] == "late_only" and s["results"] is results
    empty = cli.summary_record(method, args, [])
    assert empty["avg_train_time"] == 0 and empty["avg_final_loss"] == 0 and empty["num_successful"] == 0
    json.dumps(s)
    if layout["config"] and method == "lora":
        adapter = {"lora": {k: 0 for k in layout["config"]["lora"]}, "training": {k: 0 for k in layout["config"]["training"]}}
        adapter["lora"]["implementation"] = "builtin"
        cfg = cli.experiment_config(method, args, adapter, {"total": 2})
        _covers(cfg, layout["config"])
        assert cfg["method"] == "lora_tta_builtin" and cfg["clip_gate"]["sampling_mode"] == "late_only"
    if method == "full":
        assert "total_params" in s
        cfg = cli.experiment_config(method, args, {"training": {k: 0 for k in layout["config"]["training"]}}, {"total": 2})
        _covers(cfg, layout["config"])
        assert cfg["method"] == "full_tta" and cfg["clip_gate"]["sampling_mode"] == "late_only"


@pytest.mark.parametrize("method,loop_out", [
    ("lora", {}), ("full", {}), ("delta_a", {"delta_norm": 0.1}), ("delta_b", {"delta_norms": [0.1, 0.2]}),
    ("delta_c", {"delta_out_norm": 0.1, "delta_out_values": [0.0] * 16}), ("norm_tune", {}), ("film", {"correction_norm": 0.3}),
])
def test_per_video_record_carries_the_reference_keys(golden_dir, method, loop_out):
    """The record appended per video = run()'s identification fields + training_record(...) + total_time; it holds every
    key of the reference's ``result`` literal for the method (generation outputs gen_time / output_path excepted: they
    only exist when a video is generated, there as here)."""
    import json
    layout = json.loads((golden_dir / "output_layout.json").read_text())[method]
    args = cli.build_parser(method).parse_args(["--output-dir", "/tmp/x"])
    r = {"losses": [1.0, 0.5], "train_time": 2.0, "es_check_time": 0.5, "early_stopping_info": {"total_checks": 2}, **loop_out}
    rec = {"idx": 0, "video_name": "v", "video_path": "", "caption": "c", "batch_size": 1, "num_neighbors": 0,
           **cli.training_record(method, args, r), "total_time": 3.0}
    want = set(layout["result_keys"]) | (set(layout["result_keys_later"]) - {"gen_time", "output_path"})
    assert want <= set(rec), want - set(rec)
    assert rec["final_loss"] == 0.5 and rec["num_train_steps"] == 2 and rec["success"] is True
    assert cli.training_record(method, args, {**r, "losses": []})["final_loss"] is None


def test_frame_budget_resolution_and_guard_match_reference_golden(golden_dir, capsys):
    """resolve_tta_frames + validate_tta_feature_budget against the reference's own post-parse block and guard, executed
    by oracle/make_golden_frame_budget.py over 8 400 settings: resolved window, returned info, printed lines, error text."""
    import argparse
    import gzip
    import json
    from longcat_video_tta_b200.common import resolve_tta_frames, validate_tta_feature_budget
    rows = json.loads(gzip.open(golden_dir / "frame_budget.json.gz").read())
    assert len(rows) == 8400 and sum(r["error"] is not None for r in rows) > 1000
    for r in rows:
        args = argparse.Namespace(**r["given"])
        capsys.readouterr()
        info = err = None
        try:
            resolve_tta_frames(args)
            info = validate_tta_feature_budget(args, context="lora_tta")
        except RuntimeError as e:
            err = str(e)
        assert (args.tta_total_frames, args.tta_context_frames) == (r["total"], r["context"]), r["given"]
        assert info == r["info"] and err == r["error"], r["given"]
        assert capsys.readouterr().out.splitlines() == r["printed"], r["given"]


def test_default_frames_with_early_stopping_are_refused_like_the_reference(tmp_path):
    """The reference's own defaults (2 conditioning frames = 1 latent frame, early stopping on) leave no held-out frame:
    its guard stops the run before the model is loaded, and so does ours."""
    with pytest.raises(RuntimeError, match="ES is enabled but estimated val_latents=0"):
        cli.run("lora", f"--output-dir {tmp_path} --synthetic --model tiny --device cpu".split())


def test_run_delta_a_batch_host_flow(tmp_path, monkeypatch):
    """--batch-videos K on delta-A: K host-resident videos handed to _optimize_delta_a_batch, no early stopper
    (run_delta_a.py batch branch), the reference's record fields."""
    import torch.nn as nn
    calls = []
    torch = _stub_engine(monkeypatch, calls)

    class Wrapper:
        def __init__(self, dit, dim):
            self.delta = nn.Parameter(torch.zeros(dim))

        def trainable(self):
            return [self.delta]

    def batch(wrapper, batch_data, **kw):
        calls.append(("delta_a_batch", batch_data, kw))
        return {"losses": [0.7] * kw["num_steps"], "delta_norm": 0.25, "es_check_time": 0.0, "early_stopping_info": None}

    monkeypatch.setattr(cli.A, "DeltaAWrapper", Wrapper)
    monkeypatch.setattr(cli.A, "_optimize_delta_a_batch", batch)
    s = cli.run("delta_a", (f"--output-dir {tmp_path} --synthetic --model tiny --device cpu --latent-hw 8,8 --tta-total-frames 17 "
                            f"--tta-context-frames 5 --max-videos 1 --delta-steps 5 --delta-lr 2e-3 --batch-videos 2").split())
    kind, data, kw = calls[-1]
    assert kind == "delta_a_batch" and len(data) == 2 and kw["num_steps"] == 5 and kw["lr"] == 2e-3
    r = s["results"][0]
    assert r["success"] and (r["batch_size"], r["num_neighbors"], r["delta_norm"], r["num_train_steps"]) == (2, 1, 0.25, 5)
    assert r["early_stopping_info"] is None and s["method"] == "delta_a" and s["batch_videos"] == 2


def test_run_norm_tune_with_delta_host_flow(tmp_path, monkeypatch):
    """--also-tune-delta: the wrapper's delta vector is appended to the list the loop optimises (run_norm_tune_tta.py
    :382-385), it is part of the per-video reset, and it is counted as trainable."""
    import json
    import torch.nn as nn
    calls = []
    torch = _stub_engine(monkeypatch, calls)
    norm = [nn.Parameter(torch.ones(4)), nn.Parameter(torch.ones(4))]

    class Wrapper:
        def __init__(self, dit, also_tune_delta=False, adaln_tembed_dim=512):
            self.delta = nn.Parameter(torch.zeros(adaln_tembed_dim)) if also_tune_delta else None

        def trainable(self):
            return norm + ([self.delta] if self.delta is not None else [])

    def loop(wrapper, params, cond, train, pe, pm, **kw):
        calls.append(("norm", [p.detach().clone() for p in params], wrapper))
        with torch.no_grad():
            for p in params:
                p.add_(1.0)                      # "training" moves everything, the next video must start from the init
        return {"losses": [0.9, 0.8], "early_stopping_info": None}

    monkeypatch.setattr(cli.A, "collect_norm_params", lambda dit, target: list(norm))
    monkeypatch.setattr(cli.A, "NormTuneForward", Wrapper)
    monkeypatch.setattr(cli.A, "optimize_norm_params", loop)
    s = cli.run("norm_tune", (f"--output-dir {tmp_path} --synthetic --model tiny --device cpu --latent-hw 8,8 "
                              f"--tta-total-frames 17 --tta-context-frames 5 --max-videos 2 --es-disable --also-tune-delta").split())
    assert [r["success"] for r in s["results"]] == [True, True]
    for kind, seen, wrapper in calls[-2:]:
        assert kind == "norm" and len(seen) == 3 and seen[-1].shape == (512,)
        assert float(seen[0][0]) == 1.0 and float(seen[-1].abs().sum()) == 0.0          # reset before every video
    cfg = json.loads((tmp_path / "config.json").read_text())
    assert cfg["trainable_params"] == 4 + 4 + 512 and cfg["norm_tune"]["norm_target"] == "all_norm"
