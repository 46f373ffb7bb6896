"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the data-parallel step keeps adapter replicas identical and
equals the single-GPU update with the mean gradient of the two draws."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    import torch.distributed as dist
    from longcat_video_tta_b200 import dist as D, lora
    from longcat_video_tta_b200.dit import B200DiT
    from longcat_video_tta_b200.stepper import TTAStepper
    D.init_from_env("nccl")
    BF16 = torch.bfloat16
    dit = B200DiT.random_init("tiny", seed=0, device=f"cuda:{rank}")
    torch.manual_seed(7)
    mods = lora.inject_lora_into_dit(dit, rank=16, alpha=32.0)
    g = torch.Generator().manual_seed(1)
    cond = torch.randn(1, 16, 2, 32, 32, generator=g).to(BF16).cuda()
    train = torch.randn(1, 16, 2, 32, 32, generator=g).to(BF16).cuda()
    prompt = torch.randn(1, 1, 512, 512, generator=g).to(BF16).cuda()
    mask = torch.ones(1, 512, dtype=torch.int64).cuda()
    st = TTAStepper(dit)
    assert st.world == world
    gen = torch.Generator(device="cuda").manual_seed(D.draw_seed(42, rank))
    for i in range(2):
        sigma = torch.rand(1, device="cuda", generator=gen) * 0.999 + 0.001
        eps = torch.randn(train.shape, device="cuda", generator=gen).to(BF16)
        st.step(cond, train, prompt, mask, sigma, eps, 2e-4)
    params = [e["master"] for e in st.group.entries]
    D.assert_replicas_in_sync(params)
    out[rank] = float(torch.stack([p.double().sum() for p in params]).sum())
    dist.destroy_process_group()


def test_two_rank_nccl_replicas_stay_identical():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == out[1]


def _full_worker(rank, world, port, overlap, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      B200TTA_OVERLAP_ALLREDUCE="1" if overlap else "0",
                      B200TTA_DETERMINISTIC="1", B200TTA_ATTN_BWD="split")      # fixed summation order where there is a switch
    torch.cuda.set_device(rank)
    import torch.distributed as dist
    from longcat_video_tta_b200 import dist as D
    from longcat_video_tta_b200.dit import B200DiT
    from longcat_video_tta_b200.stepper import TTAStepper
    D.init_from_env("nccl")
    BF16 = torch.bfloat16
    dit = B200DiT.random_init("tiny", seed=0, device=f"cuda:{rank}")
    dit.requires_grad_(True)
    g = torch.Generator().manual_seed(1)
    cond = torch.randn(1, 16, 2, 32, 32, generator=g).to(BF16).cuda()
    train = torch.randn(1, 16, 2, 32, 32, generator=g).to(BF16).cuda()
    prompt = torch.randn(1, 1, 512, 512, generator=g).to(BF16).cuda()
    mask = torch.ones(1, 512, dtype=torch.int64).cuda()
    st = TTAStepper(dit, full=True, optimizer="sgd", weight_decay=0.01, max_grad_norm=1.0)
    assert st.world == world and (st.eng.on_block_grads is not None) == overlap
    gen = torch.Generator(device="cuda").manual_seed(D.draw_seed(42, rank))
    for i in range(2):
        sigma = torch.rand(1, device="cuda", generator=gen) * 0.999 + 0.001
        eps = torch.randn(train.shape, device="cuda", generator=gen).to(BF16)
        st.step(cond, train, prompt, mask, sigma, eps, 1e-3)
    D.assert_replicas_in_sync([p.data for p in dit.parameters()])
    torch.save({"grads": st.eng.full.flat.cpu(), "names": st.eng.full.names,
                "params": torch.cat([p.data.float().reshape(-1) for p in dit.parameters()]).cpu()},
               os.path.join(out_dir, f"full_{int(overlap)}_{rank}.pt"))
    dist.destroy_process_group()


def test_full_model_gradient_allreduce_overlapped_equals_blocking(tmp_path):
    """full-model TTA on 2 ranks: the per-block all-reduces issued during the backward (default) must leave the same
    mean gradients -- every parameter, including the norm weights written after the block loop and the embedders --
    and the same updated parameters as one blocking all-reduce of the flat buffer; replicas stay identical either way."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    res = {}
    for overlap in (True, False):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        mp.spawn(_full_worker, args=(2, port, overlap, str(tmp_path)), nprocs=2, join=True)
        res[overlap] = [torch.load(tmp_path / f"full_{int(overlap)}_{r}.pt") for r in range(2)]
    for overlap in (True, False):
        assert torch.equal(res[overlap][0]["grads"], res[overlap][1]["grads"])      # both ranks hold the same sums
        assert torch.equal(res[overlap][0]["params"], res[overlap][1]["params"])
    a, b = res[True][0], res[False][0]
    assert float(a["grads"].abs().max()) > 0
    # kernels sum in a timing-dependent order: equal up to fp32 rounding, per parameter tensor
    off = 0
    from longcat_video_tta_b200.dit import B200DiT
    shapes = [p.numel() for p in B200DiT.random_init("tiny", seed=0, device="cpu").parameters()]
    for name, n in zip(a["names"], shapes):
        ga, gb = a["grads"][off: off + n].double(), b["grads"][off: off + n].double()
        off += n
        denom = gb.norm().item()
        assert denom > 0, f"{name}: zero gradient"
        assert (ga - gb).norm().item() <= 1e-3 * denom, f"{name}: {(ga - gb).norm().item() / denom}"
    # bf16 parameters: an update that differs in the last fp32 bits can round to the neighbouring bf16 value
    assert (a["params"] != b["params"]).float().mean().item() < 1e-3


def test_method_script_under_torchrun(tmp_path):
    """The LoRA script started the way a user would (torchrun, 2 ranks): both ranks finish, rank 0 writes the
    reference's files once, early stopping ran on agreed anchor losses (tests/test_dist_cpu.py covers the host flow on
    gloo; this is the same path on NCCL with the real engine)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(root / "lora_experiment" / "scripts" / "run_lora_tta.py"), "--output-dir", str(tmp_path),
           "--synthetic", "--model", "tiny", "--latent-hw", "32,32", "--tta-total-frames", "17", "--tta-context-frames", "5",
           "--lora-rank", "16", "--lora-alpha", "32", "--num-steps", "4", "--max-videos", "2", "--es-check-every", "1",
           "--skip-generation"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    summary = json.loads((tmp_path / "summary.json").read_text())
    assert summary["num_successful"] == 2 and summary["method"] == "lora_tta"
    for row in summary["results"]:
        assert row["early_stopping_info"]["total_checks"] >= 2 and row["final_loss"] > 0
    assert json.loads((tmp_path / "checkpoint.json").read_text())["next_idx"] == 2
