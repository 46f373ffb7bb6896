"""CPU, build container only: the reference's OWN injection functions, unmodified (oracle/ref_bridge.py), applied to a
B200DiT -- the module-surgery seam of SURVEY 8b.  The engine must find every adapter they create, with the rank, scale,
parameter objects and parameter order the reference's optimizer would see.  Skipped where /root/reference is absent."""
import pytest
import torch

from oracle import ref_bridge
from oracle.dit_oracle import build_oracle_dit
from longcat_video_tta_b200 import lora as L
from longcat_video_tta_b200.dit import B200DiT

pytestmark = pytest.mark.skipif(not ref_bridge.available(), reason="/root/reference is not present")


def _dit():
    return B200DiT.from_oracle(build_oracle_dit("tiny", seed=0), device="cpu")


@pytest.mark.parametrize("kw", [
    dict(rank=16, alpha=32.0, target_modules=["qkv", "proj"], target_ffn=False, target_blocks="all"),
    dict(rank=4, alpha=8.0, target_modules=["qkv"], target_ffn=True, target_blocks="last_1"),
])
def test_reference_custom_injection_is_recognised(kw):
    ref = ref_bridge.load("run_lora_tta")
    dit = _dit()
    torch.manual_seed(3)
    mods = ref.inject_lora_into_dit(dit, dropout=0.0, **kw)
    dit.engine.resolve_sites(force=True)
    sites = dit.engine.lora_sites()
    assert len(sites) == len(mods)
    ref_params = ref.get_lora_parameters(mods)
    got = [p for s in sites for p in s.params]
    assert len(got) == len(ref_params) and all(a is b for a, b in zip(got, ref_params))     # same objects, same order
    for s, m in zip(sites, mods):
        assert s.r_total == kw["rank"] and s.scale == pytest.approx(m.scaling) and s.W is m.original.weight
    # and it is the layout our own injection produces on a second model
    ours = _dit()
    torch.manual_seed(3)
    L.inject_lora_into_dit(ours, dropout=0.0, **kw)
    ours.engine.resolve_sites(force=True)
    assert [s.name for s in ours.engine.lora_sites()] == [s.name for s in sites]
    for a, b in zip(ours.engine.lora_sites(), sites):
        assert torch.equal(a.A_param, b.A_param) and torch.equal(a.B_param, b.B_param)


def test_reference_builtin_injection_is_recognised_and_unhooked():
    ref = ref_bridge.load("run_lora_tta")
    dit = _dit()
    torch.manual_seed(5)
    mods = ref.inject_builtin_lora_into_dit(dit, rank=8, alpha=16.0, target_modules=("qkv", "proj"), target_ffn=False,
                                            target_blocks="all")
    dit.engine.resolve_sites()                         # picked up without force: the hooked forward changes the signature
    sites = dit.engine.lora_sites()
    assert [s.name for s in sites] == [m.lora_name for m in mods]
    ref_params = ref.get_builtin_lora_parameters(mods)
    got = [p for s in sites for p in s.params]
    assert len(got) == len(ref_params) and all(a is b for a, b in zip(got, ref_params))
    by = {s.name.split(".", 2)[2]: s for s in sites if s.name.startswith("blocks.0.")}
    assert by["attn.qkv"].r_total == 24 and len(by["attn.qkv"].blocks_up) == 3            # n_separate = 3 (:149)
    assert by["cross_attn.kv_linear"].r_total == 16 and len(by["cross_attn.kv_linear"].blocks_up) == 2
    assert by["attn.proj"].r_total == 8 and by["attn.proj"].blocks_up is None
    assert all(s.scale == pytest.approx(1.0 * 16.0 / 8) for s in sites)
    # staged block-diagonal B reproduces lora_up(lora_down(x)) of the reference module
    s, m = by["attn.qkv"], mods[0]
    with torch.no_grad():
        for b in m.lora_up.blocks:
            b.weight.normal_()
    s.refresh()
    x = torch.randn(5, s.in_features).to(torch.bfloat16)
    want = m.lora_up(m.lora_down(x)).float()
    have = (x.float() @ s.A.float().t()) @ s.B.float().t()
    assert torch.allclose(have, want, rtol=2e-2, atol=2e-2)
    ref.unhook_builtin_lora(dit)
    dit.engine.resolve_sites()
    assert dit.engine.lora_sites() == []


@pytest.mark.parametrize("mode", ["full", "shift_scale", "scale_only"])
def test_reference_film_hooks_are_folded_into_the_engine(mode):
    """run_film_tta.py:146-163: apply_to_dit() puts `output + expand(correction)` hooks on every adaLN_modulation.  The
    engine turns them into per-block additive terms and pushes d loss / d(adaLN output) back onto the corrections."""
    from longcat_video_tta_b200.engine import HookedModulationAdapter, Extras
    ref = ref_bridge.load("run_film_tta")
    dit = _dit()
    C = dit.config.hidden_size
    w = ref.FiLMAdapterWrapper(dit, num_groups=2, hidden_size=C, film_mode=mode)
    with torch.no_grad():
        for i, c in enumerate(w.corrections):
            c.copy_(torch.randn(c.shape, generator=torch.Generator().manual_seed(i)))
    assert not HookedModulationAdapter.present(dit)
    w.apply_to_dit()
    assert HookedModulationAdapter.present(dit)
    ad = HookedModulationAdapter(dit)
    assert len(ad.trainable()) == 2 and all(a is b for a, b in zip(ad.trainable(), w.corrections))
    ex = ad.build_extras()
    assert ex.need_dmod
    for b in range(len(dit.blocks)):
        want = w._expand_correction(w.corrections[w._get_group_idx(b)]).detach()
        assert torch.equal(ex.film[b], want.float())
    # backward: a made-up d(adaLN output) per block [T, 6C]; the reference's gradient is autograd through its own hook
    T = 3
    ex.d_mod = [torch.randn(T, 6 * C, generator=torch.Generator().manual_seed(10 + b)) for b in range(len(dit.blocks))]
    got = ad.grads_from(ex)
    outs = [torch.zeros(1, T, 6 * C, requires_grad=False) for _ in dit.blocks]
    total = 0
    for b, blk in enumerate(dit.blocks):
        y = outs[b]
        for h in blk.adaLN_modulation._forward_hooks.values():
            y = h(blk.adaLN_modulation, (None,), y)
        total = total + (y[0] * ex.d_mod[b]).sum()
    want = torch.autograd.grad(total, list(w.corrections))
    for g, wnt in zip(got, want):
        assert torch.allclose(g, wnt, rtol=1e-5, atol=1e-5)
    w.remove_from_dit()
    assert not HookedModulationAdapter.present(dit)
