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
    a = cli.build_parser("lora").parse_args("--output-dir /tmp/x --num-cond-frames 14".split())
    total, ctx, n_lat, n_ctx = cli.frame_budget(a)
    assert (total, ctx, n_lat, n_ctx) == (14, 14, 4, 4)
    import torch
    c, t, v = split_tta_latents(torch.zeros(1, 1, n_lat, 1, 1), n_ctx)
    assert (c.shape[2], t.shape[2], v) == (3, 1, None)      # early stopping silently inactive
    a = cli.build_parser("lora").parse_args("--output-dir /tmp/x --tta-total-frames 117 --tta-context-frames 13".split())
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
