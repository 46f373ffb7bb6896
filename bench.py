#!/usr/bin/env python3
"""bench.py -- TTA steps/s of the LongCat-Video 13.6B LoRA r=16 step at 480p / 93 frames on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (CPU arm: the oracle port on the host cores)

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
def cpu_sample(steps: int, warmup: int, threads: int):
    """Bounded CPU sample of the SAME workload through the oracle port: ONE block of the 13.6 B architecture
    (hidden 4096, 32 heads, FFN 11008, LoRA r=16 on qkv,proj) on 1 context + 1 noised 480p latent frame (3120 tokens),
    512 text tokens, fp32; scaled to whole steps by the algorithmic-FLOP ratio of the two geometries."""
    import torch
    from oracle.dit_oracle import build_oracle_dit
    from oracle import tta_oracle as T
    torch.set_num_threads(threads)
    L_s, Tc, Tt, Hl, Wl, M = 1, 1, 1, 60, 104, 512
    dit = build_oracle_dit("13.6b", seed=0, init_std=0.02, depth=L_s)
    torch.manual_seed(7)
    mods = T.inject_lora(dit, rank=16, alpha=32.0)
    g = torch.Generator().manual_seed(1)
    cond = torch.randn(1, 16, Tc, Hl, Wl, generator=g)
    train = torch.randn(1, 16, Tt, Hl, Wl, generator=g)
    prompt = torch.randn(1, 1, M, 4096, generator=g)
    mask = torch.ones(1, M, dtype=torch.int64)
    times = []

    def on_step(**kw):
        times.append(time.perf_counter())

    t0 = time.perf_counter()
    T.lora_tta_loop(dit, mods, cond, train, prompt, mask, num_steps=warmup + steps, lr=2e-4, warmup_steps=3, on_step=on_step)
    stamps = [t0] + times
    per = [b - a for a, b in zip(stamps[:-1], stamps[1:])][warmup:]
    t_sample = sorted(per)[len(per) // 2]
    tpf = (Hl // 2) * (Wl // 2)
    fa_s = f_alg(4096, 11008, L_s, (Tc + Tt) * tpf, Tc * tpf, M, 16, lora_sites_qkv_proj)["total"]
    fa_f = f_alg(4096, 11008, 48, 24 * tpf, 4 * tpf, 512, 16, lora_sites_qkv_proj)["total"]
    scale = fa_f / fa_s
    return dict(t_sample=t_sample, scale=scale, value=1.0 / (t_sample * scale),
                sample=(f"oracle port (PyTorch fp32, {threads} threads): 1 of 48 blocks of the 13.6B architecture on 2 of 24 latent "
                        f"frames (3120 tokens), {steps} timed steps, median {t_sample:.2f} s/sample-step, scaled by the "
                        f"algorithmic-FLOP ratio {scale:.0f}x to whole steps"))


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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 3))
    r = cpu_sample(steps, min(args.warmup, 1), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 / r["value"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    with contextlib.redirect_stdout(sys.stderr):  # stdout carries exactly one JSON line
        mods = lora.inject_lora_into_dit(dit, rank=16, alpha=32.0, target_modules=["qkv", "proj"])
    g = torch.Generator().manual_seed(1)  # the same video on every rank; the (sigma, eps) draws differ per rank
    cond_h = torch.randn(1, 16, Tc, Hl, Wl, generator=g).to(BF16).pin_memory()
    train_h = torch.randn(1, 16, Tt, Hl, Wl, generator=g).to(BF16).pin_memory()
    prompt_h = torch.randn(1, 1, M, cfg.caption_channels, generator=g).to(BF16).pin_memory()
    mask_h = torch.ones(1, M, dtype=torch.int64).pin_memory()
    cond, train, prompt, mask = (t.to(dev) for t in (cond_h, train_h, prompt_h, mask_h))
    torch.manual_seed(42 + rank)
    stepper = TTAStepper(dit, eps=1e-8, weight_decay=0.01, max_grad_norm=1.0, master_weights=True)

    def one_step(c, t, p, m, i):
        lr = lora._warmup_lr(2e-4, i, 3)
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
    work = f_alg(cfg.hidden_size, cfg.ffn_dim, cfg.depth, geo.N, geo.Nc, geo.M, 16, lora_sites_qkv_proj)
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
                "launches_per_step": d["n"], "avg_launch_ms": d["ms"] / d["n"],
                "algorithmic_tflop_per_launch": d["flops"] / d["n"] / 1e12, "share_of_step": d["ms"] / tot_ms}
    cpu = cpu_tiny = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_sample(2, 1, os.cpu_count() or 1)
        cpu = {"value": r["value"], "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": r["sample"]}
        cpu_tiny = cpu_tiny_sample(os.cpu_count() or 1)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD if cfg_name == "13.6b" and (Tc, Tt, Hl, Wl, M) == (4, 20, 60, 104, 512) else
                   f"{cfg_name} LoRA r=16 TTA step, latent [16,{Tc}+{Tt},{Hl},{Wl}], {M} text tokens"
                   + ("" if args.bsa_sparsity is None else
                      f", block-sparse self-attention (128-token chunks 4x4x8, sparsity {args.bsa_sparsity})"),
                   "tokens": geo.N, "adapter_params": stepper.n_params, "parallelism": f"dp{world} over noise draws",
                   "recompute": "per-block forward re-run in the backward except self-attention (O, LSE kept for all blocks) "
                                "and whatever fits the spare-HBM activation stash (blocks covered: " + _stash_summary(stepper.eng) + ")",
                   "l2": "inputs far exceed L2: 27 GB of frozen weights + 15 GB of block inputs are streamed every step"},
        "clocks": clocks,
        "e2e": {"value": world / (ms_e2e / args.steps / 1000.0), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": cpu,
        "cpu_baseline_config0": cpu_tiny,
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
    if name in ("b200tta_attn_fwd", "b200tta_attn_bwd"):
        segs, n_seg, heads = (a[13], a[14], a[11]) if name.endswith("fwd") else (a[22], a[23], a[20])
        pairs = sum((segs[i].q_end - segs[i].q_begin) * segs[i].kv_len for i in range(n_seg))
        return 4.0 * heads * 128 * pairs * (1.0 if name.endswith("fwd") else 2.5)
    return None


_BSA = {"pairs_per_head": 0.0}


def _family(name, a):
    fam = name.replace("b200tta_", "")
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
    committed `ncu --set full` capture (profiles/r1_ncu_attention_full.json, headline shape); None if not captured.
    attn_bwd[self] = dq_kernel + dkv_kernel (the delta pre-pass, 0.6 GB algorithmic, was not captured)."""
    kernels = {"attn_bwd[self]": ("dq_kernel", "dkv_kernel"), "attn_fwd[self]": ("attn_fwd",)}.get(family)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_ncu_attention_full.json")
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
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="13.6b", choices=["13.6b", "tiny"])
    ap.add_argument("--cond-frames", type=int, default=4)
    ap.add_argument("--train-frames", type=int, default=20)
    ap.add_argument("--lat-h", type=int, default=60)
    ap.add_argument("--lat-w", type=int, default=104)
    ap.add_argument("--text-tokens", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bsa-sparsity", type=float, default=None,
                    help="block-sparse self-attention (BASELINE.json configs[4], e.g. 0.9375 with --lat-h 96 --lat-w 160)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
