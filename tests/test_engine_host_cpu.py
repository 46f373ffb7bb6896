"""CPU: host-side logic of the engine that can silently go wrong -- caches keyed on recyclable addresses
and the forward-hook seam on adaLN_modulation (refusal of anything that is not additive)."""
import gc

import pytest
import torch

from oracle.dit_oracle import build_oracle_dit
from longcat_video_tta_b200.dit import B200DiT
from longcat_video_tta_b200.engine import HookedModulationAdapter

BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def dit():
    return B200DiT.from_oracle(build_oracle_dit("tiny", seed=0), device="cpu")


def test_pack_text_cache_survives_freed_and_reallocated_prompts(dit):
    eng = dit.engine
    mask = torch.zeros(1, 512, dtype=torch.int64)
    mask[:, :100] = 1
    seen = []
    for video in range(6):   # per-video scope: the previous prompt is dropped before the next one is allocated
        prompt = torch.full((1, 1, 512, 512), float(video), dtype=BF16)
        rows = eng.pack_text(prompt, mask)
        assert rows.shape == (512, 512) and float(rows[0, 0]) == float(video), "stale text rows from a previous video"
        assert eng.pack_text(prompt, mask) is rows                      # same tensors, unchanged: cache hit
        seen.append(prompt.data_ptr())
        del prompt, rows
        gc.collect()
    # in-place edits are seen too
    prompt = torch.zeros(1, 1, 512, 512, dtype=BF16)
    a = float(eng.pack_text(prompt, mask)[0, 0])    # (on the CPU the packed rows may alias the prompt: read now)
    prompt.add_(1)
    b = float(eng.pack_text(prompt, mask)[0, 0])
    assert a == 0.0 and b == 1.0


def test_site_signature_holds_modules_not_ids(dit):
    from longcat_video_tta_b200 import lora
    eng = dit.engine
    for trial in range(4):
        mods = lora.inject_lora_into_dit(dit, rank=4, alpha=8.0, target_modules=["proj"], target_blocks="last_1")
        eng.resolve_sites()
        sites = eng.lora_sites()
        assert len(sites) == 2 and all(s.A_param is m.lora_down.weight for s, m in zip(sites, mods))
        for m in mods:   # un-inject; the wrapper objects die here
            for parent in (dit.blocks[-1].attn, dit.blocks[-1].cross_attn):
                if getattr(parent, "proj", None) is m:
                    parent.proj = m.original
        del mods, sites
        gc.collect()
        eng.resolve_sites()
        assert eng.lora_sites() == []


def test_additive_hook_is_folded_and_other_hooks_are_refused(dit):
    C = dit.config.hidden_size
    corr = torch.nn.Parameter(torch.arange(6 * C, dtype=torch.float32) / C)

    def additive(_m, _i, out):
        return out + corr.unsqueeze(0).unsqueeze(0).to(out.dtype)

    h = dit.blocks[1].adaLN_modulation.register_forward_hook(additive)
    try:
        ad = HookedModulationAdapter(dit)
        assert ad.trainable() == [corr] or (len(ad.trainable()) == 1 and ad.trainable()[0] is corr)
        ex = ad.build_extras()
        assert ex.film[0] is None and torch.equal(ex.film[1], corr.detach())
        ex.d_mod = [None, torch.ones(2, 6 * C)]
        (g,) = ad.grads_from(ex)
        assert torch.equal(g, torch.full((6 * C,), 2.0))
    finally:
        h.remove()
    h = dit.blocks[0].adaLN_modulation.register_forward_hook(lambda _m, _i, out: out * 1.5)
    try:
        with pytest.raises(NotImplementedError, match="additive"):
            HookedModulationAdapter(dit).build_extras()
    finally:
        h.remove()


def test_parameter_list_must_be_the_injected_adapter_set():
    """the fused stepper trains every adapter the engine finds; a subset handed to the loop must be refused."""
    from longcat_video_tta_b200 import lora
    d = B200DiT.from_oracle(build_oracle_dit("tiny", seed=0), device="cpu")
    mods = lora.inject_lora_into_dit(d, rank=4, alpha=8.0, target_modules=["qkv", "proj"])
    params = lora.get_lora_parameters(mods)
    lora._check_param_list(d, params)                                   # the full list, in injection order: fine
    with pytest.raises(NotImplementedError, match="not the adapter set"):
        lora._check_param_list(d, params[:-2])
    with pytest.raises(NotImplementedError, match="not the adapter set"):
        lora._check_param_list(d, list(reversed(params)))
    params[0].requires_grad_(False)
    with pytest.raises(NotImplementedError, match="frozen"):
        lora._check_param_list(d, params)


def test_full_model_tta_refuses_partially_frozen_models():
    from longcat_video_tta_b200 import full
    d = B200DiT.from_oracle(build_oracle_dit("tiny", seed=0), device="cpu")
    with pytest.raises(ValueError, match="No trainable parameters"):       # run_full_tta.py:126-128
        full._check_all_trainable(d)
    d.blocks[0].attn.qkv.weight.requires_grad_(True)
    with pytest.raises(NotImplementedError, match="EVERY parameter"):
        full._check_all_trainable(d)
    d.requires_grad_(True)
    assert len(full._check_all_trainable(d)) == sum(1 for _ in d.parameters())


def test_full_grads_layout_is_one_flat_buffer_in_parameter_order():
    from longcat_video_tta_b200.engine import FullGrads
    d = B200DiT.from_oracle(build_oracle_dit("tiny", seed=0), device="cpu")
    fg = FullGrads(d)
    assert fg.flat.dtype == torch.float32 and fg.flat.numel() == sum(p.numel() for p in d.parameters())
    off = 0
    for (name, p) in d.named_parameters():
        g = fg.g(p)
        assert g.shape == p.shape and g.data_ptr() == fg.flat.data_ptr() + 4 * off and fg.named()[name] is g
        off += p.numel()
