"""Shared helpers of the GPU parity tests: matched (fp32 oracle, B200DiT) pairs with identical bf16-valued weights and
LoRA factors, the engine run with its debug taps, seeded inputs.  TEST INFRASTRUCTURE."""
import os

import torch

BF16, F32 = torch.bfloat16, torch.float32
COS_BAR = 0.999       # north_star
NORM_RTOL = 2e-2      # north_star


def expect(ok, msg=""):
    """assert, or (PARITY_REPORT_ONLY=1: exploration runs) print the violation and carry on"""
    if os.environ.get("PARITY_REPORT_ONLY"):
        if not ok:
            print("VIOLATION:", msg)
        return
    assert ok, msg


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-300)).item()


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-300)).item()


def chunked_exact_sdpa(q, k, v, chunk=2048):
    """softmax(q k^T / sqrt(D)) v in fp32, query-chunked so that 37 440 keys fit; autograd-friendly (checkpointed)."""
    from torch.utils.checkpoint import checkpoint
    scale = q.shape[-1] ** -0.5

    def one(qc, k, v):
        s = torch.matmul(qc.float(), k.float().transpose(-1, -2)) * scale
        return torch.matmul(torch.softmax(s, dim=-1), v.float())
    outs = [checkpoint(one, q[:, :, i:i + chunk], k, v, use_reentrant=False) for i in range(0, q.shape[2], chunk)]
    return torch.cat(outs, dim=2).to(q.dtype)


def build_pair(name, seed, *, init_std=None, sharpen=1.0, rank=16, alpha=32.0, target_ffn=True, **overrides):
    """(fp32 oracle with bf16-rounded weights on the GPU, B200DiT with the same weights, LoRA in both, B != 0)."""
    from oracle.dit_oracle import build_oracle_dit
    from oracle import tta_oracle as T
    from longcat_video_tta_b200 import lora
    from longcat_video_tta_b200.dit import B200DiT
    oracle = build_oracle_dit(name, seed=seed, init_std=init_std, **overrides)
    with torch.no_grad():
        for p in oracle.parameters():
            p.copy_(p.to(BF16).float())
        if sharpen != 1.0:   # low-entropy attention: logits scale with the product of the q / k RMSNorm gains
            for blk in oracle.blocks:
                for nrm in (blk.attn.q_norm, blk.attn.k_norm):
                    nrm.weight.mul_(sharpen)
                    nrm.weight.copy_(nrm.weight.to(BF16).float())
    dit = B200DiT.from_oracle(oracle)
    oracle = oracle.cuda()
    mods = lora.inject_lora_into_dit(dit, rank=rank, alpha=alpha, target_modules=["qkv", "proj"], target_ffn=target_ffn)
    omods = T.inject_lora(oracle, rank=rank, alpha=alpha, target_modules=("qkv", "proj"), target_ffn=target_ffn)
    assert len(mods) == len(omods)
    g = torch.Generator(device="cuda").manual_seed(seed + 100)
    with torch.no_grad():
        for m, om in zip(mods, omods):
            for w, ow in ((m.lora_down.weight, om.lora_down.weight), (m.lora_up.weight, om.lora_up.weight)):
                # both factors non-zero (B = 0 at init would make every dA vanish and leave half the backward untested)
                val = (torch.randn(w.shape, generator=g, device="cuda") * (w.shape[1] ** -0.5)).to(BF16)
                w.copy_(val)
                ow.copy_(val.float())
    return oracle, omods, dit, mods


def run_engine(dit, cond, train, prompt, mask, sigma, eps, stash_gb=None, taps=True):
    from longcat_video_tta_b200.stepper import TTAStepper
    old = os.environ.get("B200TTA_STASH_GB")
    if stash_gb is not None:
        os.environ["B200TTA_STASH_GB"] = str(stash_gb)
    try:
        eng = dit.engine
        eng._release_stash()
        eng.debug = {} if taps else None
        st = TTAStepper(dit, build_optimizer=False)
        loss = st.forward_backward(cond, train, prompt, mask, sigma, eps)
        torch.cuda.synchronize()
        dbg, eng.debug = eng.debug, None
        return loss.item(), dbg
    finally:
        if old is None:
            os.environ.pop("B200TTA_STASH_GB", None)
        else:
            os.environ["B200TTA_STASH_GB"] = old


def tiny_case():
    from oracle.make_golden import tiny_inputs, tiny_split
    latents, prompt, mask = tiny_inputs()
    cond, train, _ = tiny_split(latents)
    torch.manual_seed(42)
    sigma = torch.rand(1) * 0.999 + 0.001
    eps = torch.randn_like(train)
    to = lambda t: t.to(BF16).cuda()
    return to(cond), to(train), to(prompt), mask.cuda(), sigma.cuda(), to(eps)


def wide_inputs(T_c, T_t, Hl, Wl, Cc, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    cond = torch.randn(1, 16, T_c, Hl, Wl, generator=g, device="cuda").to(BF16)
    train = torch.randn(1, 16, T_t, Hl, Wl, generator=g, device="cuda").to(BF16)
    prompt = torch.randn(1, 1, 512, Cc, generator=g, device="cuda").to(BF16)
    mask = torch.zeros(1, 512, dtype=torch.int64, device="cuda")
    mask[:, :128] = 1
    sigma = torch.tensor([0.37], device="cuda")
    eps = torch.randn(1, 16, T_t, Hl, Wl, generator=g, device="cuda").to(BF16)
    return cond, train, prompt, mask, sigma, eps


