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


# ---------------------------------------------------------------------------------------------------------------------
# The method scripts under torchrun: host control flow of cli.run with the GPU pieces replaced by recorders, and the
# early stopper's decision input agreed across ranks.
def _script_worker(rank, world, port, out, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import types
    from longcat_video_tta_b200 import cli, early_stopping
    cfg = types.SimpleNamespace(caption_channels=32, adaln_tembed_dim=512, hidden_size=64, out_channels=16)
    dit = types.SimpleNamespace(config=cfg, engine=types.SimpleNamespace(resolve_sites=lambda: None), training=False,
                                eval=lambda: None, train=lambda: None)
    seen = {"init": [], "draw": [], "anchor": []}
    cli.B200DiT.random_init = staticmethod(lambda *a, **k: dit)
    cli.L.inject_lora_into_dit = lambda d, **kw: ["m0"]
    cli.L.count_lora_parameters = lambda mods: {"trainable": 1}
    cli.L.get_lora_parameters = lambda mods: []
    cli.L.reset_lora_weights = lambda mods: seen["init"].append(float(torch.rand(1)))     # stands for the kaiming re-init

    def loop(d, mods, cond, train, pe, pm, early_stopper=None, **kw):
        seen["draw"].append(float(torch.rand(1)))                                         # stands for (sigma, eps)
        stops = []
        for step in range(1, 5):
            stop, info = early_stopper.step(step, save_fn=lambda: step)
            stops.append(stop)
            if stop:
                break
        early_stopper.restore(restore_fn=lambda s: None)
        return {"losses": [1.0] * len(stops), "train_time": 0.0, "es_check_time": 0.0,
                "early_stopping_info": early_stopper.state}

    cli.L.finetune_lora_on_conditioning = loop
    # anchor losses that differ between the ranks in the last digits, around the "improved" threshold
    series = iter([1.0, 0.9, 0.9 + (1e-9 if rank else -1e-9), 0.9 + (-2e-9 if rank else 2e-9), 0.95, 0.96] * 4)

    def anchor(**kw):
        v = next(series)
        seen["anchor"].append(v)
        return v

    early_stopping.compute_flow_matching_loss_conditioned_fixed = anchor
    s = cli.run("lora", (f"--output-dir {out_dir} --synthetic --model tiny --device cpu --latent-hw 8,8 --tta-total-frames 17 "
                         f"--tta-context-frames 5 --max-videos 2 --es-check-every 1 --es-patience 2 --num-steps 4").split())
    out[rank] = dict(seen=seen, results=s["results"])
    dist.destroy_process_group()


def test_scripts_under_torchrun_host_flow(tmp_path):
    import json
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_script_worker, args=(world, port, out, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["seen"]["init"] == r1["seen"]["init"] and len(r0["seen"]["init"]) == 2     # identical re-initialisation
    assert all(a != b for a, b in zip(r0["seen"]["draw"], r1["seen"]["draw"]))             # own draw stream per rank
    assert r0["seen"]["anchor"] != r1["seen"]["anchor"]                                    # ranks did see different bits
    for a, b in zip(r0["results"], r1["results"]):                                         # ... and decided alike
        assert a["early_stopping_info"] == b["early_stopping_info"] and a["num_train_steps"] == b["num_train_steps"]
    h = r0["results"][0]["early_stopping_info"]["loss_history"]
    assert [v for _, v in h] == r0["seen"]["anchor"][:len(h)]                              # rank 0's values rule
    # rank 0 alone wrote the files, once
    ck = json.loads((tmp_path / "checkpoint.json").read_text())
    assert ck["next_idx"] == 2 and len(ck["results"]) == 2
    assert (tmp_path / "summary.json").exists() and (tmp_path / "config.json").exists()


def _overlap_worker(rank, world, port, out):
    """host logic of the overlapped full-model gradient all-reduce (stepper._reduce_block_grads /
    _sync_full_grads_overlapped) on gloo, with a stand-in for the engine's flat buffer"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from types import SimpleNamespace

    from longcat_video_tta_b200 import dist as D
    from longcat_video_tta_b200.stepper import TTAStepper
    D.init_from_env("gloo")
    # layout: [embedders 7 | block 0: w 10, norm 3 | block 1: w 10, norm 3 | final 5]
    names = ["emb.w", "blocks.0.w", "blocks.0.norm.weight", "blocks.1.w", "blocks.1.norm.weight", "final.w"]
    sizes = [7, 10, 3, 10, 3, 5]
    flat = torch.zeros(sum(sizes))
    views, off = {}, 0
    for n, k in zip(names, sizes):
        views[n] = flat[off: off + k]
        off += k
    gen = torch.Generator().manual_seed(100 + rank)
    local = torch.randn(flat.numel(), generator=gen)
    lv, off = {}, 0
    for n, k in zip(names, sizes):
        lv[n] = local[off: off + k]
        off += k
    full = SimpleNamespace(flat=flat, block_ranges=[(7, 20), (20, 33)], named=lambda: views)
    extras = SimpleNamespace(d_norm={"blocks.1.norm.weight": None, "blocks.0.norm.weight": None})
    st = SimpleNamespace(eng=SimpleNamespace(full=full), pg=None, _pending=[], _blocks_reduced=False, extras=extras)
    # the engine's order: final layer first, blocks from the last to the first (norm weights NOT yet written), then the
    # "late" call, then norm weights and embedders, then the optimizer's sync
    views["final.w"].copy_(lv["final.w"])
    for b in (1, 0):
        views[f"blocks.{b}.w"].copy_(lv[f"blocks.{b}.w"])
        TTAStepper._reduce_block_grads(st, b)
    assert len(st._pending) == 2
    TTAStepper._reduce_block_grads(st, None)
    for b in (1, 0):
        views[f"blocks.{b}.norm.weight"].copy_(lv[f"blocks.{b}.norm.weight"])
    views["emb.w"].copy_(lv["emb.w"])
    assert st._pending == [] and st._blocks_reduced
    TTAStepper._sync_full_grads_overlapped(st)
    assert not st._blocks_reduced
    out[rank] = dict(local=local, reduced=flat.clone())
    dist.destroy_process_group()


def test_overlapped_full_gradient_allreduce_covers_every_element_once():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_overlap_worker, args=(world, port, out), nprocs=world, join=True)
    want = out[0]["local"] + out[1]["local"]
    for r in range(world):
        assert torch.equal(out[r]["reduced"], want)      # each element summed over the ranks exactly once
