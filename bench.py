#!/usr/bin/env python3
"""bench.py -- TTA steps/s of the LongCat-Video 13.6B LoRA r=16 step at 480p / 93 frames on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
  python bench.py --method {lora,delta_a,delta_b,delta_c,norm_tune,film}   (configs[2]: the non-LoRA adapter families)
  python bench.py --impl reference ...                     (CPU arm: the oracle port on the host cores, bounded sample)
  python bench.py --impl torch_gpu ...                     (library arm: the same step in eager bf16 PyTorch on the GPU --
                                                            cuDNN SDPA, cuBLAS linears, per-block checkpoint, foreach AdamW)

One "step" = one flow-matching TTA update on one (sigma, eps) draw: noise + patchify, DiT forward (48 blocks),
adapter-only backward with per-block recompute, clip, AdamW.  At N > 1 every rank processes its own draw of the same
video and the flat fp32 adapter-gradient buffer is all-reduced once per step (weak scaling over draws).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC, UNIT = "tta_steps_per_sec", "steps/s"
_OUT = sys.stdout
WORKLOAD = ("LongCat-Video 13.6B LoRA r=16 (qkv,proj; 48 blocks) TTA step, bf16, synthetic 480p 93-frame latent "
            "[16,24,60,104] = 4 context + 20 noised latent frames (37440 tokens), 512 text tokens")


METHOD_NAMES = {
    "lora": "LoRA r=16 (qkv,proj)", "delta_a": "delta-A (one [512] timestep-embedding offset)",
    "delta_b": "delta-B (4 groups of [512] timestep-embedding offsets, per-tensor clip)",
    "delta_c": "delta-C (16-channel output bias; forward only)",
    "norm_tune": "norm-tune all_norm (LN affine + q/k RMSNorm weights, 417 792 params)",
    "film": "FiLM full (4 groups x [6C] additive adaLN corrections)",
    "full": "full-model TTA (all 13.6 B parameters, SGD momentum 0; SURVEY 8f row 3)",
}


# ------------------------------------------------------------------------------------------------ work model
def f_alg(C, F, L, N, Nc, M, r, sites_per_block):
    """algorithmic FLOPs of one step (fwd + dX only for frozen GEMMs, attention bwd = 2.5x fwd, LoRA x3);
    no recompute counted."""
    Nn = N - Nc
    f_lin = 2 * L * (N * (3 * C * C + C * C + 3 * C * F) + Nn * 2 * C * C + M * 2 * C * C)
    f_attn = 4 * L * C * (Nc * Nc + Nn * N)
    f_x = 4 * L * C * Nn * M
    f_lora = 2 * r * L * sum(tok * (i + o) for tok, i, o in sites_per_block(N, Nn, M, C))
    return dict(f_lin=f_lin, f_attn=f_attn, f_x=f_x, f_lora=f_lora,
                total=2 * f_lin + 3.5 * (f_attn + f_x) + 3 * f_lora)


def lora_sites_qkv_proj(N, Nn, M, C):
    return [(N, C, 3 * C), (N, C, C), (Nn, C, C), (M, C, 2 * C), (Nn, C, C)]


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.2)
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def _mem_available_bytes():
    try:
        with open("/proc/meminfo") as f:
            return next(int(ln.split()[1]) for ln in f if ln.startswith("MemAvailable")) * 1024.0
    except Exception:
        return 32e9


def _cpu_block_sample(Tc, Tt, threads, steps=1):
    """ONE TTA step (noise draw, forward, adapter backward, clip, AdamW) of ONE block of the 13.6 B architecture
    (hidden 4096, 32 heads, FFN 11008, LoRA r=16 on qkv,proj) through the oracle port in fp32 on the host cores, on
    Tc context + Tt noised 480p latent frames and 512 text tokens.  Attention is evaluated query-chunk by query-chunk
    (same arithmetic, bounded memory).  Returns wall seconds of the step."""
    import torch
    import oracle.dit_oracle as D
    from oracle import tta_oracle as T
    from torch.utils.checkpoint import checkpoint
    torch.set_num_threads(threads)
    mem_budget = 0.1 * _mem_available_bytes()

    def sdpa_chunked(q, k, v):
        scale = q.shape[-1] ** -0.5

        def one(qc, k, v):
            return torch.matmul(torch.softmax(torch.matmul(qc, k.transpose(-1, -2)) * scale, dim=-1), v)
        # chunk (and re-compute the chunk in the backward) only when the score matrix would not fit the host memory:
        # S, P and dS of one attention call at 4 B each against a tenth of what is available
        c = max(128, int(mem_budget / (12.0 * q.shape[1] * k.shape[2])))
        if q.shape[2] <= c:
            return one(q, k, v)
        return torch.cat([checkpoint(one, q[:, :, i:i + c], k, v, use_reentrant=False)
                          for i in range(0, q.shape[2], c)], dim=2)

    keep = D._sdpa
    D._sdpa = sdpa_chunked
    try:
        Hl, Wl, M = 60, 104, 512
        dit = D.build_oracle_dit("13.6b", seed=0, init_std=0.02, depth=1)
        torch.manual_seed(7)
        mods = T.inject_lora(dit, rank=16, alpha=32.0)
        g = torch.Generator().manual_seed(1)
        cond = torch.randn(1, 16, Tc, Hl, Wl, generator=g)
        train = torch.randn(1, 16, Tt, Hl, Wl, generator=g)
        prompt = torch.randn(1, 1, M, 4096, generator=g)
        mask = torch.ones(1, M, dtype=torch.int64)
        stamps = [time.perf_counter()]
        T.lora_tta_loop(dit, mods, cond, train, prompt, mask, num_steps=steps, lr=2e-4, warmup_steps=3,
                        on_step=lambda **kw: stamps.append(time.perf_counter()))
        return stamps[-1] - stamps[-2]      # the LAST step (earlier ones warm the allocator / oneDNN primitives up)
    finally:
        D._sdpa = keep


def cpu_sample(budget_s: float, threads: int):
    """Bounded CPU sample of the headline workload for the reference arm / cpu_baseline.

    What is TIMED: one step of one block (of 48 identical ones) through the oracle port, first on a small probe geometry
    (1 + 1 latent frames), then on the largest prefix of the headline geometry (4 context + k noised frames, k <= 20)
    whose predicted time fits ``budget_s`` and whose fp32 activations fit the host memory.  What is EXTRAPOLATED (and
    flagged as such in the JSON line): x48 blocks (exact: the blocks are identical; embedders / final layer are < 0.1 %)
    and, only if the full 24 frames did not fit the budget, the algorithmic-FLOP ratio to 24 frames."""
    tpf = 30 * 52
    fa = lambda L, Tc, Tt: f_alg(4096, 11008, L, (Tc + Tt) * tpf, Tc * tpf, 512, 16, lora_sites_qkv_proj)["total"]
    t_probe = _cpu_block_sample(1, 1, threads, steps=2)
    rate = fa(1, 1, 1) / t_probe                       # algorithmic FLOP/s of the port on this host
    avail_gb = _mem_available_bytes() / 1e9
    Tt = 0
    for k in range(20, 0, -1):   # ~0.9 GB of fp32 activations per latent frame of one block + 9 GB of weights / grads
        if fa(1, 4, k) / rate * 1.3 <= budget_s and 12.0 + 1.0 * (4 + k) <= 0.8 * avail_gb:
            Tt = k
            break
    if Tt == 0:
        Tc_s, Tt_s, t_sample = 1, 1, t_probe
    else:
        Tc_s, Tt_s = 4, Tt
        t_sample = _cpu_block_sample(4, Tt, threads)
    scale_blocks = 48.0
    scale_tokens = fa(1, 4, 20) / fa(1, Tc_s, Tt_s)
    t_full = t_sample * scale_blocks * scale_tokens
    return dict(t_sample=t_sample, t_probe=t_probe, value=1.0 / t_full, scale=scale_blocks * scale_tokens,
                scale_blocks=scale_blocks, scale_tokens=scale_tokens, frames=(Tc_s, Tt_s), tokens=(Tc_s + Tt_s) * tpf,
                seconds_timed=t_probe + (t_sample if Tt else 0.0),
                sample=(f"oracle port (PyTorch fp32, {threads} threads): ONE timed step of 1 of 48 blocks of the 13.6B "
                        f"architecture on {Tc_s}+{Tt_s} of 4+20 latent frames ({(Tc_s + Tt_s) * tpf} of 37440 tokens), "
                        f"{t_sample:.1f} s; extrapolated x{scale_blocks:.0f} (identical blocks)"
                        + (f" x{scale_tokens:.2f} (algorithmic-FLOP ratio to 24 frames)" if scale_tokens > 1.0001 else "")
                        + f" = {t_full:.0f} s per whole step"))


def cpu_tiny_sample(threads: int, steps: int = 5, warmup: int = 1):
    """BASELINE.json configs[0] as the reference can run it on a CPU (SURVEY 8d "CPU baseline timed beside it"): the tiny
    DiT (2 blocks, hidden 512), LoRA r=16 on qkv,proj, 17-frame 256x256 latent split 2 / 2 / 1, fp32, the whole LoRA loop
    of the oracle port (noise draw, forward, adapter backward, clip, AdamW).  Median seconds per step, not scaled."""
    import torch
    from oracle.dit_oracle import build_oracle_dit
    from oracle.make_golden import tiny_inputs, tiny_split
    from oracle import tta_oracle as T
    torch.set_num_threads(threads)
    latents, prompt, mask = tiny_inputs()
    cond, train, _ = tiny_split(latents)
    dit = build_oracle_dit("tiny", seed=0)
    torch.manual_seed(7)
    mods = T.inject_lora(dit, rank=16, alpha=32.0)
    stamps = [time.perf_counter()]
    torch.manual_seed(42)
    T.lora_tta_loop(dit, mods, cond, train, prompt, mask, num_steps=warmup + steps, lr=2e-4, warmup_steps=3,
                    on_step=lambda **kw: stamps.append(time.perf_counter()))
    per = sorted([b - a for a, b in zip(stamps[:-1], stamps[1:])][warmup:])
    t = per[len(per) // 2]
    return {"value": 1.0 / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"configs[0]: tiny DiT LoRA r=16 TTA step on a 17-frame 256x256 latent (1024 tokens), fp32, oracle port on "
                      f"{threads} threads, {warmup} warm-up + {steps} timed steps, median {t * 1e3:.0f} ms/step (unscaled)"}


def run_reference(args):
    """CPU arm.  The reference is a Python tree without packaging and its DiT dependency is absent, so the
    arm times the oracle port (kind "port").  A whole 13.6 B step would take ~1 h of CPU time: the line carries what was
    actually timed (``timed``), the factor it was multiplied by (``scale_factor``) and ``"extrapolated": true``; ``steps``
    counts the sample steps really executed.  configs[0] (the tiny model, the case the reference can run on a CPU) is
    measured whole and unscaled and sits inside ``cpu_baseline`` as ``config0``."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    r = cpu_sample(args.ref_budget_s, threads)
    tiny = cpu_tiny_sample(threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
        "warmup": 1, "ms_per_step": 1000.0 / r["value"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
        "extrapolated": True, "scale_factor": r["scale"],
        "timed": {"sample_steps": 1, "probe_steps": 2, "seconds": r["seconds_timed"], "sample_seconds": r["t_sample"],
                  "sample_tokens": r["tokens"], "sample_blocks": 1,
                  "note": "value / ms_per_step / e2e are sample_seconds x scale_factor, NOT a measurement of a whole step"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port-extrapolated",
                         "sample": r["sample"], "config0": tiny},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------------ library arm (PyTorch on the GPU)
def torch_gpu_steps(dev, steps, warmup, Tc, Tt, Hl, Wl, M, depth=48, seed=0):
    """The same LoRA TTA step in eager bf16 PyTorch on this GPU -- what the reference executes:
    the DiT restated in oracle/dit_oracle.py with bf16 parameters, ``F.scaled_dot_product_attention`` (cuDNN / flash
    backends) for both attentions, cuBLAS(Lt) linears, per-block ``torch.utils.checkpoint(use_reentrant=False)``
    (run_lora_tta.py:806-811), ``clip_grad_norm_`` and foreach ``torch.optim.AdamW`` on the bf16 LoRA tensors
    (run_lora_tta.py:462-468,513-514).  Returns (ms per step, losses, attention backend)."""
    import torch
    import torch.nn.functional as F
    from torch.nn.attention import sdpa_kernel, SDPBackend
    import oracle.dit_oracle as D
    from oracle import tta_oracle as T
    BF16 = torch.bfloat16
    backends = [SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION, SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH]

    def sdpa(q, k, v):
        with sdpa_kernel(backends, set_priority=True):
            return F.scaled_dot_product_attention(q, k, v)

    keep = D._sdpa
    D._sdpa = sdpa
    try:
        cfg = D.make_config("13.6b", depth=depth)
        with torch.device("meta"):
            dit = D.OracleDiT(cfg)
        dit = dit.to(BF16).to_empty(device=dev)
        g = torch.Generator(device=dev).manual_seed(seed)
        with torch.no_grad():
            for n, p in dit.named_parameters():
                if p.dim() >= 2 or n.endswith("bias"):
                    p.copy_(torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32) * 0.02)
                else:
                    p.copy_(1.0 + torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32) * 0.02)
        dit.requires_grad_(False)
        dit.gradient_checkpointing = True
        torch.manual_seed(7)
        mods = T.inject_lora(dit, rank=16, alpha=32.0, target_modules=("qkv", "proj"))
        params = T.lora_parameters(mods)
        for p in params:
            p.requires_grad_(True)
        opt = torch.optim.AdamW(params, lr=2e-4, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8)
        gi = torch.Generator().manual_seed(1)
        cond = torch.randn(1, 16, Tc, Hl, Wl, generator=gi).to(BF16).to(dev)
        train = torch.randn(1, 16, Tt, Hl, Wl, generator=gi).to(BF16).to(dev)
        prompt = torch.randn(1, 1, M, cfg.caption_channels, generator=gi).to(BF16).to(dev)
        mask = torch.ones(1, M, dtype=torch.int64, device=dev)
        losses = []

        def one_step(i):
            for gp in opt.param_groups:
                gp["lr"] = T.warmup_lr(2e-4, i, 3)
            sigma = torch.rand(1, device=dev, dtype=torch.float32) * 0.999 + 0.001
            noise = torch.randn_like(train)
            loss = T.fm_loss_given(dit, cond, train, prompt, mask, sigma, noise, BF16)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
            losses.append(loss.detach())

        torch.cuda.reset_peak_memory_stats(dev)
        for i in range(warmup):
            one_step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one_step(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        vals = [float(v) for v in torch.stack(losses).tolist()]
        peak_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30
        return ms, vals, peak_gb
    finally:
        D._sdpa = keep


def run_torch_gpu(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    sampler = ClockSampler(dev.index)
    sampler.start()
    ms, losses, peak_gb = torch_gpu_steps(dev, args.steps, args.warmup, args.cond_frames, args.train_frames, args.lat_h,
                                          args.lat_w, args.text_tokens)
    clocks = sampler.stop()
    tpf = (args.lat_h // 2) * (args.lat_w // 2)
    work = f_alg(4096, 11008, 48, (args.cond_frames + args.train_frames) * tpf, args.cond_frames * tpf, args.text_tokens, 16,
                 lora_sites_qkv_proj)
    value = 1000.0 / ms
    line = {
        "impl": "torch_gpu", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "library_stack": "eager PyTorch %s: cuDNN/flash SDPA, cuBLASLt linears, per-block "
                   "checkpoint, clip_grad_norm_, foreach AdamW (bf16 states)" % torch.__version__},
        "clocks": clocks, "achieved_tflops_per_gpu": work["total"] / (ms / 1e3) / 1e12,
        "frac_of_nominal_2250": work["total"] / (ms / 1e3) / 1e12 / 2250.0, "peak_alloc_gib": peak_gb,
        "loss_first_last": [losses[0], losses[-1]],
    }
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from longcat_video_tta_b200 import lora, ops
    from longcat_video_tta_b200.dit import B200DiT
    from longcat_video_tta_b200.stepper import TTAStepper
    BF16 = torch.bfloat16

    cfg_name = args.model
    Tc, Tt, Hl, Wl, M = args.cond_frames, args.train_frames, args.lat_h, args.lat_w, args.text_tokens
    bsa_over = {} if args.bsa_sparsity is None else dict(enable_bsa=True, bsa_params=dict(sparsity=args.bsa_sparsity))
    dit = B200DiT.random_init(cfg_name, seed=0, device=dev, **bsa_over)
    cfg = dit.config
    if args.bsa_sparsity is not None:   # block-sparse: only the attended 128 x 128 blocks count as algorithmic work
        from longcat_video_tta_b200 import bsa
        tpf_ = (Hl // 2) * (Wl // 2)
        nb, nctx = (Tc + Tt) * tpf_ // bsa.BLOCK, Tc * tpf_ // bsa.BLOCK
        keep, keep_ctx = bsa.row_counts(nb, args.bsa_sparsity, nctx)
        _BSA["pairs_per_head"] = float(nctx * keep_ctx + (nb - nctx) * keep) * bsa.BLOCK * bsa.BLOCK
    torch.manual_seed(7)
    import contextlib
    wrapper = None
    if args.method == "lora":
        with contextlib.redirect_stdout(sys.stderr):  # stdout carries exactly one JSON line
            mods = lora.inject_lora_into_dit(dit, rank=16, alpha=32.0, target_modules=["qkv", "proj"])
    elif args.method == "full":
        mods = None
        dit.requires_grad_(True)
    else:   # BASELINE.json configs[2]: the delta / norm / AdaLN(FiLM) / bias families (SURVEY 8d "Config 3")
        from longcat_video_tta_b200 import adapters
        C, Ct = cfg.hidden_size, cfg.adaln_tembed_dim
        wrapper = {
            "delta_a": lambda: adapters.DeltaAWrapper(dit, adaln_tembed_dim=Ct),
            "delta_b": lambda: adapters.DeltaBWrapper(dit, num_groups=4, adaln_tembed_dim=Ct, hidden_size=C),
            "delta_c": lambda: adapters.DeltaCWrapper(dit, mode="per_channel", out_channels=16),
            "norm_tune": lambda: adapters.NormTuneForward(dit),
            "film": lambda: adapters.FiLMAdapterWrapper(dit, num_groups=4, hidden_size=C, film_mode="full"),
        }[args.method]()
        if args.method == "norm_tune":
            for p in adapters.collect_norm_params(dit, "all_norm"):
                p.requires_grad_(True)
            wrapper = adapters.NormTuneForward(dit)
    g = torch.Generator().manual_seed(1)  # the same video on every rank; the (sigma, eps) draws differ per rank
    cond_h = torch.randn(1, 16, Tc, Hl, Wl, generator=g).to(BF16).pin_memory()
    train_h = torch.randn(1, 16, Tt, Hl, Wl, generator=g).to(BF16).pin_memory()
    prompt_h = torch.randn(1, 1, M, cfg.caption_channels, generator=g).to(BF16).pin_memory()
    mask_h = torch.ones(1, M, dtype=torch.int64).pin_memory()
    cond, train, prompt, mask = (t.to(dev) for t in (cond_h, train_h, prompt_h, mask_h))
    torch.manual_seed(42 + rank)
    if args.method == "full":   # run_full_tta.py defaults: SGD(momentum 0), lr 1e-5, warm-up 2, wd 0.01, clip 1.0
        stepper = TTAStepper(dit, full=True, optimizer="sgd", weight_decay=0.01, max_grad_norm=1.0)
    elif wrapper is None:
        stepper = TTAStepper(dit, eps=1e-8, weight_decay=0.01, max_grad_norm=1.0, master_weights=True)
    else:   # adapters._optimize: AdamW eps 1e-15, wd 0.01, constant lr 1e-3, clip 1.0 (per tensor for delta-B)
        stepper = TTAStepper(dit, adapter=wrapper, train_lora=False, eps=1e-15, weight_decay=0.01, max_grad_norm=1.0,
                             per_tensor_clip=wrapper.per_tensor_clip)

    def one_step(c, t, p, m, i):
        lr = lora._warmup_lr(1e-5, i, 2) if args.method == "full" else (lora._warmup_lr(2e-4, i, 3) if wrapper is None else 1e-3)
        sigma = torch.rand(1, device=dev, dtype=torch.float32) * 0.999 + 0.001
        noise = torch.randn_like(t)
        return stepper.step(c, t, p, m, sigma, noise, lr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    losses = []
    for i in range(args.warmup):
        losses.append(one_step(cond, train, prompt, mask, i))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    k0 = ops.kernel_launches()
    ms = timed(lambda i: losses.append(one_step(cond, train, prompt, mask, args.warmup + i)), args.steps)
    launches = ops.kernel_launches() - k0
    clocks = sampler.stop() if rank == 0 else None

    # end to end through the public API with HOST buffers: per step H2D of the latents / text from pinned memory and a
    # D2H read of the loss
    h2d = sum(t.numel() * t.element_size() for t in (cond_h, train_h, prompt_h, mask_h))
    e2e_losses = []

    def e2e_step(i):
        c, t, p, m = (x.to(dev, non_blocking=True) for x in (cond_h, train_h, prompt_h, mask_h))
        e2e_losses.append(one_step(c, t, p, m, args.warmup + args.steps + i).item())

    ms_e2e = timed(e2e_step, args.steps)

    # ---- per-kernel-family timing of one extra step (CUDA events on the launching stream around every ABI call)
    prof = {}
    if rank == 0:
        prof = profile_step(ops, lambda: one_step(cond, train, prompt, mask, 0))
    else:
        one_step(cond, train, prompt, mask, 0)   # every rank takes part in the step's all-reduce
        torch.cuda.synchronize()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    geo = dit.engine.geo
    work = f_alg(cfg.hidden_size, cfg.ffn_dim, cfg.depth, geo.N, geo.Nc, geo.M, 16 if args.method == "lora" else 0,
                 lora_sites_qkv_proj)
    if args.method == "full":       # + one weight-gradient GEMM per linear: forward, dX and dW
        work["total"] = 3 * work["f_lin"] + 3.5 * (work["f_attn"] + work["f_x"])
    if args.method == "delta_c":    # output-bias only: the gradient is a reduction of d loss / d pred, forward only
        work["total"] = work["f_lin"] + work["f_attn"] + work["f_x"]
    if args.bsa_sparsity is not None:
        f_attn = 4.0 * cfg.depth * cfg.hidden_size * _BSA["pairs_per_head"]
        work["total"] += 3.5 * (f_attn - work["f_attn"])
        work["f_attn"] = f_attn
    peaks = measured_peaks()
    t_step = ms / args.steps / 1000.0
    value = world / t_step
    tflops = work["total"] / t_step / 1e12
    loss_vals = [float(v) for v in torch.cat(losses).tolist()]
    assert all(v == v and v < 1e4 for v in loss_vals), "non-finite loss in the timed run"

    # roofline of the dominant kernel family: algorithmic FLOPs of its launches / their summed CUDA-event duration
    roof = None
    if prof:
        tot_ms = sum(v["ms"] for v in prof.values())
        fam = max((k for k in prof if prof[k]["flops"]), key=lambda k: prof[k]["ms"])
        d = prof[fam]
        ach = d["flops"] / (d["ms"] / 1000.0) / 1e12
        roof = {"kernel": fam, "bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["tflops_sustained"], "traffic": _ncu_traffic(fam),
                "peak_source": peaks["source"] + " cuBLAS bf16, sustained figure (kernel timed inside a long step)",
                "kernels": ("delta_kernel + dkvq_kernel (five products, fused) + dq_convert_kernel"
                            if os.environ.get("B200TTA_ATTN_BWD", "fused") != "split" else "delta_kernel + dq_kernel + dkv_kernel")
                if fam == "attn_bwd[self]" else fam,
                "launches_per_step": d["n"], "avg_launch_ms": d["ms"] / d["n"],
                "algorithmic_tflop_per_launch": d["flops"] / d["n"] / 1e12, "share_of_step": d["ms"] / tot_ms}
    cpu = cpu_tiny = None
    lib = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_sample(args.cpu_budget_s, os.cpu_count() or 1)
        cpu_tiny = cpu_tiny_sample(os.cpu_count() or 1)
        cpu = {"value": r["value"], "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port-extrapolated",
               "extrapolated": True, "scale_factor": r["scale"], "seconds_timed": r["seconds_timed"], "sample": r["sample"],
               "config0": cpu_tiny}
    if world == 1 and args.library_baseline and args.method == "lora" and args.bsa_sparsity is None and cfg_name == "13.6b":
        # same box, same process, after the engine's memory is returned: the step in eager bf16 PyTorch (library arm)
        n_params = stepper.n_params
        stash_txt = _stash_summary(stepper.eng)
        del stepper, mods
        dit._engine = None
        del dit
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        try:
            lms, llosses, lpeak = torch_gpu_steps(dev, 2, 2, Tc, Tt, Hl, Wl, M)
            lib = {"impl": "torch_gpu", "value": 1000.0 / lms, "unit": UNIT, "ms_per_step": lms, "steps": 2, "warmup": 2,
                   "stack": "eager PyTorch %s bf16: SDPA (cuDNN first), cuBLASLt linears, per-block checkpoint, clip_grad_norm_, "
                            "foreach AdamW" % torch.__version__, "peak_alloc_gib": lpeak, "loss_first_last": [llosses[0], llosses[-1]],
                   "speedup_of_this_repo": (ms / args.steps) and lms / (ms / args.steps)}
        except Exception as e:   # the library arm must never take the headline line down with it
            lib = {"impl": "torch_gpu", "unavailable": f"{type(e).__name__}: {str(e)[:300]}"}
    else:
        n_params = stepper.n_params
        stash_txt = _stash_summary(stepper.eng)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": (WORKLOAD if args.method == "lora" else WORKLOAD.replace(
                       "LoRA r=16 (qkv,proj; 48 blocks)", METHOD_NAMES[args.method]))
                   if cfg_name == "13.6b" and (Tc, Tt, Hl, Wl, M) == (4, 20, 60, 104, 512) else
                   f"{cfg_name} {METHOD_NAMES[args.method]} TTA step, latent [16,{Tc}+{Tt},{Hl},{Wl}], {M} text tokens"
                   + ("" if args.bsa_sparsity is None else
                      f", block-sparse self-attention (128-token chunks 4x4x8, sparsity {args.bsa_sparsity})"),
                   "method": args.method, "tokens": geo.N, "adapter_params": n_params, "parallelism": f"dp{world} over noise draws",
                   "recompute": "per-block forward re-run in the backward except self-attention (O, LSE kept for all blocks) "
                                "and whatever fits the spare-HBM activation stash (blocks covered: " + stash_txt + ")",
                   "l2": "inputs far exceed L2: 27 GB of frozen weights + 15 GB of block inputs are streamed every step"},
        "clocks": clocks,
        "e2e": {"value": world / (ms_e2e / args.steps / 1000.0), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": cpu,
        "library_baseline": lib,
        "algorithmic_tflop_per_step": work["total"] / 1e12,
        "achieved_tflops_per_gpu": tflops,
        "frac_of_nominal_2250": tflops / 2250.0,
        "frac_of_measured_sustained": tflops / peaks["tflops_sustained"],
        "kernel_ms_per_step": kernel_table(prof, peaks),
        "loss_first_last": [loss_vals[0], loss_vals[-1]],
    }
    print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_table(prof, peaks):
    """Per kernel family of one profiled step: CUDA-event time, launches, algorithmic TFLOP/s and -- the second half of
    BASELINE.json's metric, "attn/GEMM tensor-pipe % of peak" -- that rate as a fraction of the nominal 2 250 TFLOP/s and
    of the measured sustained cuBLAS figure (None for the HBM-bound families, which carry no FLOP count)."""
    table = {}
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        row = {"ms": round(v["ms"], 3), "n": v["n"], "tflops": None}
        if v["flops"] and v["ms"] > 0:
            t = v["flops"] / (v["ms"] / 1e3) / 1e12
            row.update({"tflops": round(t, 1), "frac_of_nominal_2250": round(t / 2250.0, 4),
                        "frac_of_measured_sustained": round(t / peaks["tflops_sustained"], 4)})
        table[k] = row
    return table


def _call_flops(name, a):
    """Algorithmic FLOPs of one C-ABI call, from its own arguments (None for memory-bound kernels)."""
    if name == "b200tta_lora_linear_fwd":      # (X, ldx, W, W_hi, A, B, XA, n_tok, in, out, r, ...)
        n_tok, fin, fout, r = a[7], a[8], a[9], a[10]
        return 2.0 * n_tok * fin * fout * (2 if a[3] else 1) + 2.0 * n_tok * r * (fin + fout)
    if name == "b200tta_lora_linear_bwd":      # (dY, lddy, X, ldx, W, A, B, XA, U, dA, dB, n_tok, in, out, r, scale, epi, ..)
        n_tok, fin, fout, r = a[11], a[12], a[13], a[14]
        has_dx = a[16] is not None
        return (2.0 * n_tok * fin * fout if has_dx else 0.0) + 2.0 * n_tok * r * (fin + fout) * (3 if has_dx else 2)
    if name == "b200tta_gemm":                 # (M, N, segs, nseg, epi, stream)
        return sum(2.0 * a[0] * a[1] * a[2][i].k for i in range(a[3]))
    if name in ("b200tta_attn_bsa_fwd", "b200tta_attn_bsa_bwd"):   # 128 x 128 token pairs actually attended (set by main)
        heads = a[10] if name.endswith("fwd") else a[19]
        return 4.0 * heads * 128 * _BSA["pairs_per_head"] * (1.0 if name.endswith("fwd") else 2.5)
    if name in ("b200tta_attn_fwd", "b200tta_attn_bwd", "b200tta_attn_bwd_fused"):   # the fused entry shares attn_bwd's leading arguments
        segs, n_seg, heads = (a[13], a[14], a[11]) if name.endswith("fwd") else (a[22], a[23], a[20])
        pairs = sum((segs[i].q_end - segs[i].q_begin) * segs[i].kv_len for i in range(n_seg))
        return 4.0 * heads * 128 * pairs * (1.0 if name.endswith("fwd") else 2.5)
    return None


_BSA = {"pairs_per_head": 0.0}


def _family(name, a):
    fam = name.replace("b200tta_", "")
    if fam == "attn_bwd_fused":     # same operation, same family: which kernels ran is recorded in roofline.kernels
        fam = "attn_bwd"
    if fam in ("attn_bsa_fwd", "attn_bsa_bwd"):
        return fam.replace("_bsa", "") + "[self, block-sparse]"
    if fam in ("attn_fwd", "attn_bwd"):
        n_q, n_kv = (a[9], a[10]) if fam == "attn_fwd" else (a[18], a[19])
        return fam + ("[self]" if n_q == n_kv else "[cross]")
    if fam == "lora_linear_fwd":
        return "linear_fwd (tcgen05 GEMM + lora_down)"
    if fam == "lora_linear_bwd":
        return "linear_bwd (tcgen05 GEMM dX + LoRA grads)"
    return fam


def _ncu_traffic(family):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel family from the
    committed `ncu --set full` captures (profiles/r2_ncu_kernels_full.json, headline shape, one launch per kernel); None if
    not captured.  attn_bwd[self] = dkvq_kernel (fused, the default) or dq_kernel + dkv_kernel (B200TTA_ATTN_BWD=split);
    the delta pre-pass and the fp32 -> bf16 dQ conversion (0.6 + 0.9 GB algorithmic) were not captured."""
    fused = os.environ.get("B200TTA_ATTN_BWD", "fused") != "split"
    kernels = {"attn_bwd[self]": ("dkvq_kernel",) if fused else ("dq_kernel", "dkv_kernel"),
               "attn_fwd[self]": ("attn_fwd_kernel",)}.get(family)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2_ncu_kernels_full.json")
    if kernels is None or not os.path.exists(path):
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        with open(path) as f:
            cap = json.load(f)
        total = 0.0
        for k in kernels:
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                val, u = cap[k][m].split()
                total += float(val) * unit[u]
        return total
    except (KeyError, ValueError):
        return None


def _stash_summary(eng):
    st = getattr(eng, "_stash", None)
    if not st:
        return "none"
    return ", ".join(f"{k[2:]} {st[k]}/{eng.L}" for k in ("k_x1", "k_x2", "k_qkv", "k_h"))


def profile_step(ops, fn):
    """Time every C-ABI call of one extra step with CUDA events on the launching stream.
    Returns {family: {ms, n, flops}} (flops = algorithmic FLOPs of those calls, None for memory-bound families)."""
    import torch
    events = []
    orig = ops._call

    def wrapped(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        events.append((name, _family(name, a), _call_flops(name, a), e0, e1))

    ops._call = wrapped
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        ops._call = orig
    out = {}
    for name, fam, fl, e0, e1 in events:
        d = out.setdefault(fam, {"ms": 0.0, "n": 0, "flops": 0.0 if fl is not None else None})
        d["ms"] += e0.elapsed_time(e1)
        d["n"] += 1
        if fl is not None and d["flops"] is not None:
            d["flops"] += fl
    return out


# ------------------------------------------------------------------------------------------------ text-encode workload
UMT5_XXL = dict(vocab_size=256384, d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64,
                relative_attention_num_buckets=32, relative_attention_max_distance=128, layer_norm_epsilon=1e-6)


def _umt5_state(cfg, device, seed=0):
    """random-init UMT5 encoder weights under transformers' parameter names, bf16, transformers' own initialisation
    scales (scores of order one, so a bf16 model stays close to fp32 over 24 layers)"""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    d, inner, ff = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"]

    def rn(*shape, std, mean=0.0):
        return (torch.randn(*shape, generator=g, device=device) * std + mean).to(torch.bfloat16)

    st = {"shared.weight": rn(cfg["vocab_size"], d, std=1.0)}
    st["encoder.embed_tokens.weight"] = st["shared.weight"]
    for l in range(cfg["num_layers"]):
        a, f = f"encoder.block.{l}.layer.0.", f"encoder.block.{l}.layer.1."
        st[a + "SelfAttention.q.weight"] = rn(inner, d, std=(d * cfg["d_kv"]) ** -0.5)
        st[a + "SelfAttention.k.weight"] = rn(inner, d, std=d ** -0.5)
        st[a + "SelfAttention.v.weight"] = rn(inner, d, std=d ** -0.5)
        st[a + "SelfAttention.o.weight"] = rn(d, inner, std=inner ** -0.5)
        st[a + "SelfAttention.relative_attention_bias.weight"] = rn(cfg["relative_attention_num_buckets"],
                                                                    cfg["num_heads"], std=0.5)
        st[a + "layer_norm.weight"] = rn(d, std=0.1, mean=1.0)
        st[f + "DenseReluDense.wi_0.weight"] = rn(ff, d, std=d ** -0.5)
        st[f + "DenseReluDense.wi_1.weight"] = rn(ff, d, std=d ** -0.5)
        st[f + "DenseReluDense.wo.weight"] = rn(d, ff, std=ff ** -0.5)
        st[f + "layer_norm.weight"] = rn(d, std=0.1, mean=1.0)
    st["encoder.final_layer_norm.weight"] = rn(d, std=0.1, mean=1.0)
    return st


def _hf_umt5(cfg, state, dtype, device):
    """the third-party implementation behind the reference's encode_prompt (common.py:33,62-64,250)"""
    import torch
    from transformers import UMT5Config, UMT5EncoderModel
    c = UMT5Config(vocab_size=cfg["vocab_size"], d_model=cfg["d_model"], d_kv=cfg["d_kv"], d_ff=cfg["d_ff"],
                   num_layers=cfg["num_layers"], num_heads=cfg["num_heads"],
                   relative_attention_num_buckets=cfg["relative_attention_num_buckets"],
                   relative_attention_max_distance=cfg["relative_attention_max_distance"],
                   layer_norm_epsilon=cfg["layer_norm_epsilon"], feed_forward_proj="gated-gelu", dropout_rate=0.0)
    with torch.device(device):
        m = UMT5EncoderModel(c).to(dtype).eval()
    m.load_state_dict({k: v.to(dtype) for k, v in state.items()}, strict=False)
    return m


def _umt5_work(cfg, B, N):
    d, inner, ff = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"]
    gemm = cfg["num_layers"] * 2 * B * N * (4 * inner * d + 3 * d * ff)
    attn = cfg["num_layers"] * 4 * B * cfg["num_heads"] * N * N * cfg["d_kv"]
    return gemm, attn


def _text_encode_cpu(budget_layers: int, N: int, n_valid: int, threads: int):
    """the reference's own implementation (transformers UMT5EncoderModel, fp32) on the host cores: a bounded sample of
    `budget_layers` of the 24 identical layers at full width, extrapolated by the layer count"""
    import torch
    torch.set_num_threads(threads)
    cfg = dict(UMT5_XXL, num_layers=budget_layers, vocab_size=1024)
    st = _umt5_state(cfg, "cpu")
    try:
        m, kind = _hf_umt5(cfg, st, torch.float32, "cpu"), "reference"
        fn = lambda ids, mask: m(ids, mask).last_hidden_state     # noqa: E731
    except Exception:                                             # transformers missing: the oracle port
        from oracle import umt5_oracle
        kind = "port"
        fn = lambda ids, mask: umt5_oracle.umt5_encode(st, cfg, ids, mask)     # noqa: E731
    ids = torch.randint(2, 1024, (1, N))
    mask = torch.zeros(1, N, dtype=torch.long)
    mask[:, :n_valid] = 1
    with torch.no_grad():
        fn(ids, mask)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            fn(ids, mask)
        dt = (time.perf_counter() - t0) / reps
    scale = UMT5_XXL["num_layers"] / budget_layers
    return {"value": 1.0 / (dt * scale), "unit": "encodes/s", "cores": threads, "kind": kind + "-extrapolated",
            "extrapolated": True, "scale_factor": scale, "seconds_timed": dt * reps,
            "sample": f"transformers UMT5EncoderModel fp32 on {threads} threads: {budget_layers} of 24 identical xxl-width "
                      f"layers at {N} tokens, {dt:.2f} s per pass, x{scale:.0f} layers = {dt * scale:.1f} s per encode"}


def run_text_encode(args):
    """SURVEY 8(f) row 4, text half: one 512-token prompt through the UMT5-xxl encoder (encode_prompt, common.py:228-255).
    A "step" is one encode; N > 1 = independent replicas (prompts do not shard), no collective."""
    import torch
    import torch.distributed as dist
    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N, n_valid = args.text_tokens, 77
    if args.impl == "reference":
        if rank == 0:
            r = _text_encode_cpu(2, N, n_valid, os.cpu_count() or 1)
            line = {"impl": "reference", "metric": "prompt_encodes_per_sec", "value": r["value"], "unit": "encodes/s",
                    "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / r["value"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": f"UMT5-xxl encoder, one {N}-token prompt ({n_valid} real tokens)"},
                    "cpu_baseline": r, "e2e": {"value": r["value"], "unit": "encodes/s", "h2d_bytes_per_step": 0,
                                               "d2h_bytes_per_step": 0}, "extrapolated": True}
            _OUT.write(json.dumps(line) + "\n")
            _OUT.flush()
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from longcat_video_tta_b200 import ops
    from longcat_video_tta_b200.text_encoder import B200UMT5Encoder
    cfg = UMT5_XXL
    st = _umt5_state(cfg, dev)
    enc = B200UMT5Encoder(st, device=dev, **{k: v for k, v in cfg.items() if k != "vocab_size"})
    g = torch.Generator().manual_seed(1 + rank)
    ids_h = torch.randint(2, cfg["vocab_size"], (1, N), generator=g)
    mask_h = torch.zeros(1, N, dtype=torch.long)
    mask_h[:, :n_valid] = 1
    ids_h[mask_h == 0] = 0
    ids_h, mask_h = ids_h.pin_memory(), mask_h.pin_memory()
    ids, mask = ids_h.to(dev), mask_h.to(dev)
    out_h = torch.empty(1, N, cfg["d_model"], dtype=torch.bfloat16).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    steps = max(args.steps, 20)          # an encode is ~7 ms: time at least 20 of them
    for _ in range(max(args.warmup, 3)):
        enc(ids, mask)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    k0 = ops.kernel_launches()
    ms = timed(lambda: enc(ids, mask), steps)          # 11 GB of weights per encode: far beyond L2
    launches = ops.kernel_launches() - k0
    clocks = sampler.stop() if rank == 0 else None

    def e2e():
        o = enc(ids_h.to(dev, non_blocking=True), mask_h.to(dev, non_blocking=True)).last_hidden_state
        out_h.copy_(o, non_blocking=True)
        torch.cuda.synchronize()

    ms_e2e = timed(e2e, steps)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # per-call device time of one encode (CUDA events on the launching stream around every ABI call)
    fam, orig = {}, ops._call

    def probe(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        fam.setdefault(name, []).append((e0, e1))

    ops._call = probe
    enc(ids, mask)
    torch.cuda.synchronize()
    ops._call = orig
    per = {k: {"calls": len(v), "ms": sum(a.elapsed_time(b) for a, b in v)} for k, v in fam.items()}
    peaks = measured_peaks()
    f_gemm, f_attn = _umt5_work(cfg, 1, N)
    t = ms / steps / 1000.0
    gm = per["b200tta_gemm"]
    ach = f_gemm / gm["calls"] / (gm["ms"] / gm["calls"] / 1e3) / 1e12
    line = {"metric": "prompt_encodes_per_sec", "value": world / t, "unit": "encodes/s", "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"UMT5-xxl encoder (24 layers, d_model 4096, 64 heads x 64, d_ff 10240), one {N}-token "
                                   f"prompt ({n_valid} real tokens, right-padded) per step, random-init bf16 weights",
                       "parallelism": f"{world} independent replica(s), no collective",
                       "l2": "9.3 GB of layer weights streamed per encode: far beyond L2"},
            "clocks": clocks,
            "e2e": {"value": world / (ms_e2e / steps / 1000.0), "unit": "encodes/s",
                    "h2d_bytes_per_step": ids_h.numel() * 8 + mask_h.numel() * 8,
                    "d2h_bytes_per_step": out_h.numel() * 2},
            "gpu_launches": launches,
            "roofline": {"kernel": "gemm2_kernel (q|k|v, o, wi_0|wi_1 + GEGLU, wo; 96 launches per encode)", "bound": "tensor",
                         "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                         "frac": ach / peaks["tflops_sustained"], "traffic": None,
                         "peak_source": peaks["source"] + " cuBLAS bf16, sustained figure",
                         "launches_per_step": gm["calls"], "avg_launch_ms": gm["ms"] / gm["calls"],
                         "algorithmic_tflop_per_launch": f_gemm / gm["calls"] / 1e12,
                         "share_of_step": gm["ms"] / sum(v["ms"] for v in per.values()),
                         "note": "512 rows: a 4096-wide layer is 32 CTA-pair tiles for 74 pairs (DESIGN 6.5)"},
            "kernel_ms_per_encode": per,
            "algorithmic_tflop_per_step": (f_gemm + f_attn) / 1e12,
            "achieved_tflops_per_gpu": (f_gemm + f_attn) / t / 1e12}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = _text_encode_cpu(2, N, n_valid, os.cpu_count() or 1)
    if world == 1 and args.library_baseline:
        try:
            with torch.no_grad():
                m16 = _hf_umt5(cfg, st, torch.bfloat16, dev)
                y16 = m16(ids, mask).last_hidden_state.float()
                ours = enc(ids, mask).last_hidden_state.float()
                for _ in range(3):
                    m16(ids, mask)
                ms_lib = timed(lambda: m16(ids, mask), steps)
                del m16
                torch.backends.cuda.matmul.allow_tf32 = False
                m32 = _hf_umt5(cfg, st, torch.float32, dev)
                y32 = m32(ids, mask).last_hidden_state
                del m32
            rl2 = lambda a, b: float((a - b).norm() / b.norm())     # noqa: E731
            import transformers
            line["library_baseline"] = {"impl": f"transformers {transformers.__version__} UMT5EncoderModel, eager bf16, same B200",
                                        "value": 1000.0 / (ms_lib / steps), "unit": "encodes/s", "ms_per_step": ms_lib / steps,
                                        "speedup_of_this_repo": ms_lib / ms}
            line["parity_at_size"] = {"reference": "transformers UMT5EncoderModel fp32 (TF32 off), same bf16-valued weights",
                                      "ours_rel_l2": rl2(ours, y32), "library_bf16_rel_l2": rl2(y16, y32),
                                      "ours_cosine": float(torch.dot(ours.flatten(), y32.flatten()) / (ours.norm() * y32.norm()))}
        except Exception as e:  # noqa: BLE001
            line["library_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()
    if world > 1:
        dist.destroy_process_group()


def _protect_stdout():
    """stdout must carry exactly ONE JSON line: route everything else (NCCL banners, library prints) to stderr and
    return a file object bound to the real stdout."""
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    return real


def main():
    global _OUT
    _OUT = _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch_gpu"])
    ap.add_argument("--method", default="lora", choices=list(METHOD_NAMES),
                    help="adapter family: lora = BASELINE.json configs[1]; the others = configs[2]")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="--impl reference: CPU seconds for the block sample")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0, help="cpu_baseline leg of the default run")
    ap.add_argument("--no-library-baseline", dest="library_baseline", action="store_false",
                    help="skip the same-box eager-PyTorch step (library_baseline key; N=1 only)")
    ap.add_argument("--model", default="13.6b", choices=["13.6b", "tiny"])
    ap.add_argument("--cond-frames", type=int, default=4)
    ap.add_argument("--train-frames", type=int, default=20)
    ap.add_argument("--lat-h", type=int, default=60)
    ap.add_argument("--lat-w", type=int, default=104)
    ap.add_argument("--text-tokens", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bsa-sparsity", type=float, default=None,
                    help="block-sparse self-attention (BASELINE.json configs[4], e.g. 0.9375 with --lat-h 96 --lat-w 160)")
    ap.add_argument("--workload", default="tta_step", choices=["tta_step", "text_encode"],
                    help="tta_step = BASELINE.json's metric (default); text_encode = SURVEY 8(f) row 4, one prompt per step")
    args = ap.parse_args()
    if args.workload == "text_encode":
        run_text_encode(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch_gpu":
        run_torch_gpu(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
