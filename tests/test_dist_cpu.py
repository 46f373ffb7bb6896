"""CPU, world_size 2, gloo: the host-side logic of the data-parallel path -- per-rank draw seeds, the
single all-reduce of the flat adapter-gradient buffer, the 1/world factor folded into clip + AdamW, and replicas that
stay bit-identical.  (The kernels themselves only run on a B200; here the optimizer arithmetic is the oracle's.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from longcat_video_tta_b200 import dist as D
    from oracle import tta_oracle as T
    assert D.init_from_env("gloo") == world and D.rank() == rank and D.world_size() == world
    # replicas start identical; every rank draws its own (sigma, eps) from seed base + rank
    torch.manual_seed(0)
    params = [torch.randn(16, 32), torch.randn(48, 16)]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    gen = torch.Generator().manual_seed(D.draw_seed(42, rank))
    flat = torch.randn(sum(p.numel() for p in params), generator=gen)          # this rank's flat gradient buffer
    local = flat.clone()
    scale = D.all_reduce_grads([flat])                                          # ONE collective per step
    assert scale == 1.0 / world
    grads, off = [], 0
    for p in params:
        grads.append((flat[off:off + p.numel()] * scale).view_as(p).clone())
        off += p.numel()
    total = T.clip_grad_norm(grads, 1.0)
    T.adamw_step(params, grads, m, v, 1, 2e-4, eps=1e-8, wd=0.01)
    D.assert_replicas_in_sync(params)
    out[rank] = dict(local=local, params=[p.clone() for p in params], total=float(total))
    dist.destroy_process_group()


def test_two_rank_gradient_mean_and_identical_updates():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    from longcat_video_tta_b200 import dist as D
    from oracle import tta_oracle as T
    r0, r1 = out[0], out[1]
    assert not torch.equal(r0["local"], r1["local"])                            # different draws per rank
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)                                                # replicas bit-identical
    # == single-process update with the mean gradient over the two draws
    torch.manual_seed(0)
    params = [torch.randn(16, 32), torch.randn(48, 16)]
    mean = (r0["local"] + r1["local"]) / 2
    grads, off = [], 0
    for p in params:
        grads.append(mean[off:off + p.numel()].view_as(p).clone())
        off += p.numel()
    total = T.clip_grad_norm(grads, 1.0)
    T.adamw_step(params, grads, [torch.zeros_like(p) for p in params], [torch.zeros_like(p) for p in params], 1, 2e-4,
                 eps=1e-8, wd=0.01)
    assert abs(float(total) - r0["total"]) < 1e-6
    for a, b in zip(params, r0["params"]):
        assert torch.allclose(a, b, rtol=0, atol=1e-7)
    # world = 1 degenerates to the single-GPU run (same seed, no collective, scale 1)
    assert D.draw_seed(42, 0) == 42 and D.all_reduce_grads([torch.ones(3)]) == 1.0
