"""Import the reference's *unmodified* Python for the TTA step in this container.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` exists only in the build container
(not on the GPU box): nothing under ``-m gpu`` tests, ``smoke()`` or ``bench.py``
may call this.  It is used by ``oracle/make_golden.py`` and by the CPU tests that
validate ``oracle/tta_oracle.py`` against the reference's own functions.

How: the five ``longcat_video.*`` modules the reference imports at module scope
(``delta_experiment/scripts/common.py:33-39``, ``lora_experiment/scripts/run_lora_tta.py:101``)
are un-vendored, so they are stubbed in ``sys.modules``; ``LoRAModule`` gets a small
restatement because the builtin-LoRA path instantiates it.
Four reference scripts are truncated mid-``main()`` in the snapshot
(``run_delta_a.py``, ``run_delta_c.py``, ``run_film_tta.py``, ``run_norm_tune_tta.py``);
their classes / functions are extracted with ``ast`` and exec'd without ``main``.
"""
from __future__ import annotations

import ast
import importlib
import math
import sys
import types
from pathlib import Path

import torch
import torch.nn as nn

REF_ROOT = Path("/root/reference")
_DELTA = REF_ROOT / "delta_experiment" / "scripts"
_LORA = REF_ROOT / "lora_experiment" / "scripts"


def available() -> bool:
    return (_DELTA / "common.py").is_file()


class _StubLoRAUp(nn.Module):
    """``lora_up`` container with ``.blocks[i]: Linear(r, out / n_sep)`` (A.10)."""

    def __init__(self, r, out, n_sep):
        super().__init__()
        self.blocks = nn.ModuleList([nn.Linear(r, out // n_sep, bias=False) for _ in range(n_sep)])
        self.r = r

    def forward(self, x):
        chunks = x.split(self.r, dim=-1)
        return torch.cat([b(c) for b, c in zip(self.blocks, chunks)], dim=-1)


class StubLoRAModule(nn.Module):
    """Restatement of upstream ``LoRAModule`` as the reference uses it
    (run_lora_tta.py:132-135, 175-181, 201-209)."""

    def __init__(self, name, org_module, multiplier=1.0, lora_dim=4, alpha=1.0, n_seperate=1):
        super().__init__()
        self.lora_name = name
        self.lora_dim = lora_dim
        in_dim, out_dim = org_module.in_features, org_module.out_features
        self.lora_down = nn.Linear(in_dim, n_seperate * lora_dim, bias=False)
        if n_seperate > 1:
            self.lora_up = _StubLoRAUp(lora_dim, out_dim, n_seperate)
        else:
            self.lora_up = nn.Linear(lora_dim, out_dim, bias=False)
        nn.init.kaiming_uniform_(self.lora_down.weight, a=math.sqrt(5))
        for p in self.lora_up.parameters():
            nn.init.zeros_(p)
        self.multiplier = multiplier
        self.alpha_scale = alpha / lora_dim
        self.use_lora = True


def _install_stubs():
    if "longcat_video" in sys.modules:
        return

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Absent:  # placeholder types: only referenced in annotations / loaders
        def __init__(self, *a, **k):
            raise RuntimeError("upstream LongCat-Video class is not available in this container")

        @classmethod
        def from_pretrained(cls, *a, **k):
            raise RuntimeError("upstream LongCat-Video checkpoints are not available in this container")

    mod("longcat_video")
    mod("longcat_video.modules")
    mod("longcat_video.modules.scheduling_flow_match_euler_discrete", FlowMatchEulerDiscreteScheduler=_Absent)
    mod("longcat_video.modules.autoencoder_kl_wan", AutoencoderKLWan=_Absent)
    mod("longcat_video.modules.longcat_video_dit", LongCatVideoTransformer3DModel=_Absent)
    mod("longcat_video.pipeline_longcat_video", LongCatVideoPipeline=_Absent, retrieve_latents=lambda *a, **k: None)
    mod("longcat_video.modules.lora_utils", LoRAModule=StubLoRAModule)


_CACHE = {}


def load(name: str):
    """Return a reference module by short name.

    ``common``, ``early_stopping``, ``run_lora_tta``, ``run_delta_b`` import whole;
    ``run_delta_a``, ``run_delta_c``, ``run_film_tta``, ``run_norm_tune_tta`` are
    ast-extracted (everything except ``main`` and the ``__main__`` guard).
    """
    if name in _CACHE:
        return _CACHE[name]
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box?)")
    _install_stubs()
    for p in (str(_DELTA), str(_LORA)):
        if p not in sys.path:
            sys.path.insert(0, p)
    if name in ("common", "early_stopping", "run_lora_tta", "run_delta_b"):
        m = importlib.import_module(name)
    else:
        path = _DELTA / f"{name}.py"
        src = path.read_text()
        try:
            tree = ast.parse(src)
        except SyntaxError as e:  # truncated file: cut at the last complete top-level stmt
            lines = src.splitlines()
            cut = e.lineno
            tree = None
            while cut > 0:
                try:
                    tree = ast.parse("\n".join(lines[:cut]))
                    break
                except SyntaxError:
                    cut -= 1
            # drop the (now half-parsed) trailing main()
        body = [n for n in tree.body
                if not (isinstance(n, ast.FunctionDef) and n.name == "main")
                and not (isinstance(n, ast.If) and "__main__" in ast.unparse(n.test))]
        tree.body = body
        m = types.ModuleType(name)
        m.__file__ = str(path)
        sys.modules[name] = m
        exec(compile(tree, str(path), "exec"), m.__dict__)
    _CACHE[name] = m
    return m
