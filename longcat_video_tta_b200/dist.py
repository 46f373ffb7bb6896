"""Data-parallel TTA over noise / timestep draws (SURVEY 8e -- our extension; the reference is single-GPU, the only
``torch.distributed`` use in it is a world-size-1 NCCL group in baseline_experiment/scripts/run_baseline.py:63-79).

One process per GPU (torchrun).  Every rank holds the frozen bf16 backbone and a replica of the adapters and processes
its own (sigma, eps) draw of the same video; the ONLY data-path collective is one all-reduce (sum) of the flat fp32
adapter-gradient buffer per step (<= 164 MB for LoRA r=16, 2 KB for delta-A) over NCCL / NVLink.  The 1/world factor is
folded into the clip-coefficient and AdamW kernels (``grad_scale``), so replicas apply bit-identical updates.
No activation or sequence exchange exists on this path, hence no fused compute+collective kernel.
"""
from __future__ import annotations

import os
from typing import Iterable, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> int:
    """Initialise the default process group from torchrun's environment; returns the world size (1 if not launched
    under torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend, **kw)
    return world


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None) -> int:
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def draw_seed(base_seed: int, rank_: int) -> int:
    """Generator seed of rank k's (sigma, eps) stream: base + k, so that world = 1 reproduces the single-GPU run."""
    return int(base_seed) + int(rank_)


def rank0_value(value: float, device="cpu", group=None) -> float:
    """Rank 0's value of a host-side scalar that feeds a control decision (the early stopper's anchor loss): replicas
    hold identical adapters, but the kernels' summation order is timing dependent, so two ranks may see anchor losses
    that differ in the last bits -- one would stop and the others would wait in the next all-reduce.  One 8-byte
    broadcast per check keeps every rank on rank 0's decision."""
    if world_size(group) == 1:
        return value
    dev = device if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return float(t.item())


def all_reduce_grads(buffers: Iterable[torch.Tensor], group=None) -> float:
    """Sum the gradient buffers over the ranks in place; returns the factor (1 / world) the optimizer kernels must
    apply to turn the sum into the mean over draws."""
    w = world_size(group)
    if w > 1:
        for b in buffers:
            dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / w


def assert_replicas_in_sync(tensors: Iterable[torch.Tensor], group=None, atol: float = 0.0) -> None:
    """Debug helper: every rank must hold the same adapter values (checked with a max/min all-reduce of a checksum)."""
    if world_size(group) == 1:
        return
    s = torch.stack([t.detach().double().sum() for t in tensors])
    hi, lo = s.clone(), s.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    if float((hi - lo).abs().max()) > atol:
        raise RuntimeError("adapter replicas diverged across ranks")
