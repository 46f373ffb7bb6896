"""BASELINE.json configs[1] at FULL size (13.6 B parameters, 4 + 20 latent frames at 480p = 37 440 tokens, 512 text tokens)
through size-independent properties -- the oracle cannot run this size, the invariants of SURVEY 8(c)(ii) can:

* a LoRA at initialisation (B = 0) IS the base model: forward-only losses with and without the injected adapters are the
  same number (to 1e-5 relative: exact zeros are added, only the order of a few fp32 sums differs);
* its first backward has dA == 0 exactly (dA = (dY B)^T X) and every dB finite and non-zero;
* the first AdamW step therefore only decays A (p * (1 - lr * wd)) and moves every B off zero;
* the second backward has dA != 0, and the loss of the step sequence is finite and goes down on the same draw.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
BF16, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def big():
    """the 13.6 B DiT (random init, 27 GB), the headline inputs and the base model's forward-only loss on one draw"""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if torch.cuda.mem_get_info()[1] < 150e9:
        pytest.skip("needs the 180 GB of a B200")
    from longcat_video_tta_b200.dit import B200DiT
    from longcat_video_tta_b200.stepper import TTAStepper
    dev = torch.device("cuda", 0)
    old = os.environ.get("B200TTA_DETERMINISTIC")
    os.environ["B200TTA_DETERMINISTIC"] = "1"             # fixed summation order in the GEMMs (read per call)
    dit = B200DiT.random_init("13.6b", seed=0, device=dev)
    g = torch.Generator().manual_seed(1)
    cond = torch.randn(1, 16, 4, 60, 104, generator=g).to(BF16).to(dev)
    train = torch.randn(1, 16, 20, 60, 104, generator=g).to(BF16).to(dev)
    prompt = torch.randn(1, 1, 512, dit.config.caption_channels, generator=g).to(BF16).to(dev)
    mask = torch.ones(1, 512, dtype=torch.int64, device=dev)
    sigma = torch.tensor([0.6], device=dev)
    noise = torch.randn(train.shape, generator=torch.Generator(device=dev).manual_seed(3), device=dev).to(BF16)
    inputs = (cond, train, prompt, mask, sigma, noise)
    loss_base = float(TTAStepper(dit, build_optimizer=False).eval_loss(*inputs))
    assert loss_base == loss_base and 0.1 < loss_base < 100.0
    yield dit, inputs, loss_base
    if old is None:
        os.environ.pop("B200TTA_DETERMINISTIC", None)
    else:
        os.environ["B200TTA_DETERMINISTIC"] = old


@pytest.mark.parametrize("method", ["delta_a", "delta_b", "delta_c", "film"])
def test_zero_init_adapters_are_the_base_model_at_headline_size(big, method):
    """configs[2] at full size: every delta / FiLM adapter starts at zero and must reproduce the base loss; the first
    backward of the two grouped families gives finite, non-zero gradients for every trainable tensor"""
    from longcat_video_tta_b200 import adapters
    from longcat_video_tta_b200.stepper import TTAStepper
    dit, inputs, loss_base = big
    C, Ct = dit.config.hidden_size, dit.config.adaln_tembed_dim
    wrapper = {"delta_a": lambda: adapters.DeltaAWrapper(dit, adaln_tembed_dim=Ct),
               "delta_b": lambda: adapters.DeltaBWrapper(dit, num_groups=4, adaln_tembed_dim=Ct, hidden_size=C),
               "delta_c": lambda: adapters.DeltaCWrapper(dit, mode="per_channel", out_channels=16),
               "film": lambda: adapters.FiLMAdapterWrapper(dit, num_groups=4, hidden_size=C, film_mode="full")}[method]()
    st = TTAStepper(dit, adapter=wrapper, train_lora=False, eps=1e-15, weight_decay=0.01, max_grad_norm=1.0,
                    per_tensor_clip=wrapper.per_tensor_clip)
    loss = float(st.eval_loss(*inputs))
    assert abs(loss - loss_base) <= 1e-5 * loss_base, (method, loss, loss_base)
    if method in ("delta_b", "film"):
        loss_fb = float(st.forward_backward(*inputs))
        assert abs(loss_fb - loss_base) <= 1e-5 * loss_base
        grads = wrapper.grads_from(st.extras)
        assert len(grads) == len(list(wrapper.trainable()))
        for i, gr in enumerate(grads):
            assert torch.isfinite(gr).all() and float(gr.abs().max()) > 0, f"{method}: gradient {i}"


def test_lora_at_init_is_the_base_model_at_headline_size(big):
    import contextlib
    import sys

    from longcat_video_tta_b200 import lora
    from longcat_video_tta_b200.stepper import TTAStepper
    dit, (cond, train, prompt, mask, sigma, noise), loss_base = big     # runs last: the injection changes the model

    torch.manual_seed(7)
    with contextlib.redirect_stdout(sys.stderr):
        lora.inject_lora_into_dit(dit, rank=16, alpha=32.0, target_modules=["qkv", "proj"])
    st = TTAStepper(dit, eps=1e-8, weight_decay=0.01, max_grad_norm=1.0, master_weights=True)
    assert st.n_params == 40_894_464                      # 851 968 per block x 48
    loss_init = float(st.eval_loss(cond, train, prompt, mask, sigma, noise))
    # exact zeros are added to the accumulators; what is left is the order of a few fp32 sums (3e-7 relative measured)
    assert abs(loss_init - loss_base) <= 1e-5 * loss_base, (loss_init, loss_base)

    loss0 = float(st.forward_backward(cond, train, prompt, mask, sigma, noise))
    assert abs(loss0 - loss_base) <= 1e-5 * loss_base
    ents = st.group.entries
    assert len(ents) == 2 * 5 * 48                        # (A, B) x {qkv, proj, q_linear, kv_linear, proj} x 48 blocks
    a_before = [e["master"].clone() for e in ents[0::2]]
    for i, (ea, eb) in enumerate(zip(ents[0::2], ents[1::2])):
        assert ea.get("grad_transposed") and not eb.get("grad_transposed")
        assert not ea["grad"].any(), f"site {i}: dA must be exactly zero while B == 0"
        gb = eb["grad"]
        assert torch.isfinite(gb).all() and float(gb.abs().max()) > 0, f"site {i}: dB"
        assert not eb["param"].any()
    lr, wd = 2e-4, 0.01
    st.optimizer_step(lr)
    for i, (ea, eb, a0) in enumerate(zip(ents[0::2], ents[1::2], a_before)):
        assert torch.allclose(ea["master"], a0 * (1.0 - lr * wd), rtol=1e-6, atol=0), f"site {i}: A only decays"
        assert eb["param"].any(), f"site {i}: B moved"
    loss1 = float(st.forward_backward(cond, train, prompt, mask, sigma, noise))
    assert loss1 == loss1 and loss1 < loss0               # same draw, one step along -dL/dB
    n_zero = sum(int(not e["grad"].any()) for e in ents[0::2])
    assert n_zero == 0, f"{n_zero} dA tensors still zero after B moved"
    print(f"headline size: base loss {loss_base:.6f} == LoRA-at-init {loss_init:.6f}; after one step {loss1:.6f}")
