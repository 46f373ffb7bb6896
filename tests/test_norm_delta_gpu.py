"""GPU: norm-tune + delta-A in ONE step (the reference's ``--also-tune-delta``, run_norm_tune_tta.py:380-390).

The two gradient paths are each covered on their own in tests/test_adapters_gpu.py; the combined wrapper asks the backward
for both at once (``Extras.norm_grads`` and ``need_dt`` / ``need_dmod``).  With the delta vector at its zero init the
forward is the one of plain norm-tune, so:
  * the norm gradients of the combined step == those of ``NormTuneForward`` alone,
  * the delta gradient of the combined step == that of ``DeltaAWrapper`` alone on the same (nudged) DiT,
  * one optimizer step moves both kinds of parameters.

Ran green on a B200 (driver GPUTEST_r01: x-passed; round 2 suite): a plain test now."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF16, F32 = torch.bfloat16, torch.float32


def _cos(a, b):
    a, b = a.float().flatten().cpu(), b.float().flatten().cpu()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()


def test_combined_step_equals_its_two_parts():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle.dit_oracle import build_oracle_dit
    from oracle.make_golden import tiny_inputs, tiny_split
    from longcat_video_tta_b200 import adapters as A
    from longcat_video_tta_b200.dit import B200DiT
    from longcat_video_tta_b200.stepper import TTAStepper

    latents, prompt, mask = tiny_inputs()
    cond, train, _ = tiny_split(latents)
    torch.manual_seed(42)
    sigma = torch.rand(1, dtype=F32) * 0.999 + 0.001
    eps = torch.randn_like(train)
    cond, train, prompt, eps = (x.to(BF16).cuda() for x in (cond, train, prompt, eps))
    mask, sigma = mask.cuda(), sigma.cuda()
    oracle = build_oracle_dit("tiny", seed=0)

    def nudged_dit():
        dit = B200DiT.from_oracle(oracle)
        params = A.collect_norm_params(dit, "all_norm")
        g = torch.Generator().manual_seed(5)
        with torch.no_grad():
            for p in params:
                p.add_((torch.randn(p.shape, generator=g) * 0.02).to(p.device, p.dtype))
        return dit, params

    def grads(wrapper):
        st = TTAStepper(wrapper.dit, adapter=wrapper, train_lora=False, eps=1e-15)
        loss = st.forward_backward(cond, train, prompt, mask, sigma, eps).item()
        return loss, [g.float().clone() for g in wrapper.grads_from(st.extras)], st

    dit, params = nudged_dit()
    for p in params:
        p.requires_grad_(True)
    loss_n, g_norm, _ = grads(A.NormTuneForward(dit))

    dit, _ = nudged_dit()
    loss_d, g_delta, _ = grads(A.DeltaAWrapper(dit, 512))

    dit, params = nudged_dit()
    for p in params:
        p.requires_grad_(True)
    both = A.NormTuneForward(dit, also_tune_delta=True, adaln_tembed_dim=512)
    loss_b, g_both, st = grads(both)

    assert abs(loss_b - loss_n) <= 1e-4 * abs(loss_n) and abs(loss_b - loss_d) <= 1e-4 * abs(loss_d)
    assert len(g_both) == len(g_norm) + 1
    flat = lambda gs: torch.cat([x.flatten() for x in gs])       # noqa: E731
    assert _cos(flat(g_both[:-1]), flat(g_norm)) > 0.9999
    assert abs((flat(g_both[:-1]).norm() / flat(g_norm).norm()).item() - 1) < 1e-3
    assert _cos(g_both[-1], g_delta[0]) > 0.9999
    assert abs((g_both[-1].norm() / g_delta[0].norm()).item() - 1) < 1e-3

    before = [p.detach().float().clone() for p in both.trainable()]
    st.optimizer_step(1e-3)
    torch.cuda.synchronize()
    moved = [float((p.detach().float() - b).abs().max()) for p, b in zip(both.trainable(), before)]
    assert moved[-1] > 0 and max(moved[:-1]) > 0                  # the delta vector and the norm parameters both moved
