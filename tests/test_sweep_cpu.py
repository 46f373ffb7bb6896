"""CPU: the sweep layer (YAML config keys -> job environment -> command line of the method scripts) against what the
reference's own run_sweep.py and run_sweep.sbatch produce for every row of its 64 configs (oracle/make_golden_sweep.py),
and every one of those command lines against our parsers."""
import gzip
import json
import subprocess
import sys
from pathlib import Path

import pytest

from longcat_video_tta_b200 import cli, sweep

REF_ROOT, REF_CKPT = "/scratch/wc3013/longcat-video-tta", "/scratch/wc3013/longcat-video-checkpoints"
REF_DATA = "/scratch/wc3013/longcat-video-tta/datasets/panda_100_480p"      # run_sweep.sbatch:30-33


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.loads(gzip.open(golden_dir / "sweep_rows.json.gz").read())


def test_config_keys_are_the_reference_keys(golden):
    assert set(golden["key_to_env"]) == set(sweep.CONFIG_KEYS)
    assert all(env == key.upper() for key, env in golden["key_to_env"].items())
    assert set(golden["methods"]) == set(sweep.METHODS)


def test_environment_of_every_row(golden):
    assert len(golden["rows"]) == 260 and len({r["config"] for r in golden["rows"]}) == 64
    for r in golden["rows"]:
        warned = []
        env = sweep.build_env_vars(r["method"], r["series_name"], r["run_id"], r["fixed"], r["row"], warn=warned.append)
        assert env == r["env"], (r["config"], r["run_id"])
        assert warned == r["warnings"]


def test_command_line_of_every_row(golden):
    seen = set()
    for r in golden["rows"]:
        script, argv = sweep.command_line(r["env"], REF_ROOT, REF_CKPT, REF_DATA)
        assert [script] + argv == r["argv"], (r["config"], r["run_id"])
        seen.add(r["method"])
    assert seen == set(sweep.METHODS)


def test_every_reference_command_line_parses_here(golden):
    """What a reference sweep would launch is accepted by the drop-in scripts, with the values arriving where the
    reference's parser would put them (spot-checked on the method's own hyper-parameters)."""
    n = 0
    for r in golden["rows"]:
        args = cli.build_parser(r["method"]).parse_args(r["argv"][1:])
        merged = {**r["fixed"], **{k: v for k, v in r["row"].items() if k != "run_id"}}
        for key, val in merged.items():
            if isinstance(val, bool) or not hasattr(args, key):
                continue
            got = getattr(args, key)
            assert got == (type(got)(val) if got is not None else val), (r["config"], r["run_id"], key, got, val)
        assert args.no_save_videos and args.clip_gate_fail_open
        n += 1
    assert n == len(golden["rows"]) == 260


def test_unknown_key_unset_bool_and_overrides():
    warned = []
    env = sweep.build_env_vars("lora", "s", "R1", {"lora_rank": 4, "target_ffn": False, "es_disable": True, "bogus": 1},
                               {"run_id": "R1", "lora_rank": 16}, data_dir="/d", output_base="/o", warn=warned.append)
    assert env == {"METHOD": "lora", "RUN_ID": "R1", "SERIES_NAME": "s", "DATA_DIR": "/d", "OUTPUT_DIR": "/o/s/R1",
                   "LORA_RANK": "16", "ES_DISABLE": "1"}
    assert warned == ["WARNING: Unknown config key 'bogus', skipping."]
    script, argv = sweep.command_line(env, "/root", "/ckpt", "/unused")
    assert script.endswith("run_lora_tta.py") and argv[:6] == ["--checkpoint-dir", "/ckpt", "--data-dir", "/d",
                                                              "--output-dir", "/o/s/R1"]
    assert "--es-disable" in argv and "--target-ffn" not in argv
    # TTA frame budget follows the conditioning frames unless set (run_sweep.sbatch:49-50)
    assert argv[argv.index("--tta-total-frames") + 1] == "2"
    with pytest.raises(ValueError):
        sweep.command_line({"METHOD": "sgd", "RUN_ID": "x"}, "/r", "/c", "/d")


def test_dry_run_of_the_shipped_configs():
    root = Path(__file__).resolve().parents[1]
    configs = sorted((root / "sweep_experiment" / "configs").glob("*.yaml"))
    assert configs
    for c in configs:
        cfg = sweep.load_config(c)
        r = subprocess.run([sys.executable, str(root / "sweep_experiment" / "scripts" / "run_sweep.py"), "--config", str(c),
                            "--account", "unused", "--dry-run", "--gpus", "2"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout.count("[DRY-RUN]") == len(cfg["sweep"])
        for row in cfg["sweep"]:
            env = sweep.build_env_vars(cfg["method"], cfg["series_name"], row["run_id"], cfg["fixed"], row)
            script, argv = sweep.command_line(env, str(root), "", "")
            cli.build_parser(cfg["method"]).parse_args(argv)


def test_rows_run_as_local_processes_one_per_gpu(tmp_path):
    (tmp_path / "s").mkdir()
    (tmp_path / "s" / "row.py").write_text(
        "import os, sys, time\n"
        "time.sleep(0.3)\n"
        "print(os.environ.get('CUDA_VISIBLE_DEVICES'), *sys.argv[1:])\n"
        "sys.exit(int(sys.argv[2]))\n")
    commands = [(f"R{i}", "s/row.py", ["--code", "3" if i == 2 else "0"]) for i in range(5)]
    done = sweep.run_rows(commands, gpus=2, log_dir=tmp_path, root=tmp_path)
    assert [d["run_id"] for d in done] == [f"R{i}" for i in range(5)]
    assert [d["returncode"] for d in done] == [0, 0, 3, 0, 0]
    gpus = [(tmp_path / f"R{i}.log").read_text().split()[0] for i in range(5)]
    assert set(gpus) == {"0", "1"} and gpus[0] != gpus[1]
    with pytest.raises(NotImplementedError):
        sweep.run_rows([("F1", "baseline_experiment/scripts/run_open_sora.py", [])])     # a reference method outside this build
