"""Test infrastructure, not product code.  Generates tests/golden/loss_inputs.pt: the reference's own four flow-matching
loss functions (delta_experiment/scripts/common.py:274-559, imported from /root/reference through oracle/ref_bridge.py)
run on CPU with a RECORDING ``forward_fn`` that returns a fixed affine function of its input.  What is pinned is
everything those functions do around the DiT: the order of the RNG draws, the noising arithmetic and its dtypes, the
per-frame timestep tensor (zero on conditioning frames, bf16-quantised sigma * 1000 elsewhere), ``num_cond_latents``,
the slice the loss is taken on, and the loss value itself.

Run here (needs /root/reference):  python oracle/make_golden_loss_inputs.py"""
import pathlib
import sys
import types

import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_bridge  # noqa: E402

SEED = 1234


def stub_dit():
    return types.SimpleNamespace(config=types.SimpleNamespace(patch_size=(1, 2, 2)))


def inputs(dtype):
    g = torch.Generator().manual_seed(5)
    cond = torch.randn(1, 16, 2, 4, 6, generator=g).to(dtype)
    tgt = torch.randn(1, 16, 3, 4, 6, generator=g).to(dtype)
    prompt = torch.randn(1, 1, 8, 32, generator=g).to(dtype)
    mask = torch.ones(1, 8, dtype=torch.int64)
    return cond, tgt, prompt, mask


def recorder(calls):
    def fwd(hidden, timestep, n_cond=None):
        calls.append({"hidden": hidden.detach().clone(), "timestep": timestep.detach().clone(), "n_cond": n_cond})
        return hidden.to(torch.float32) * 0.5 + 0.125       # stands in for the DiT: fp32 [B,16,T,H,W]
    return fwd


def run_all(cm, dtype):
    """cm: a module exposing the four loss functions (the reference's common.py, or the host mirror in the test)"""
    cond, tgt, prompt, mask = inputs(dtype)
    full = torch.cat([cond, tgt], dim=2)
    out = {}
    calls = []
    torch.manual_seed(SEED)
    loss = cm.compute_flow_matching_loss_conditioned(stub_dit(), cond, tgt, prompt, mask, device="cpu", dtype=dtype,
                                                     forward_fn=recorder(calls))
    loss2 = cm.compute_flow_matching_loss_conditioned(stub_dit(), cond, tgt, prompt, mask, sigma_min=0.2, sigma_max=0.3,
                                                      device="cpu", dtype=dtype, forward_fn=recorder(calls))
    out["conditioned"] = {"calls": calls, "losses": [loss.detach().clone(), loss2.detach().clone()]}
    calls = []
    torch.manual_seed(SEED)
    f2 = recorder(calls)
    loss = cm.compute_flow_matching_loss(stub_dit(), full, prompt, mask, device="cpu", dtype=dtype,
                                         forward_fn=lambda h, t: f2(h, t))
    out["unconditioned"] = {"calls": calls, "losses": [loss.detach().clone()]}
    calls = []
    noises = [torch.randn(tgt.shape, generator=torch.Generator().manual_seed(77 + d)).to(tgt.dtype) for d in range(2)]
    v = cm.compute_flow_matching_loss_conditioned_fixed(stub_dit(), cond, tgt, prompt, mask, fixed_sigmas=[0.25, 0.5, 0.75],
                                                        fixed_noises=noises, device="cpu", dtype=dtype,
                                                        forward_fn=recorder(calls))
    out["conditioned_fixed"] = {"calls": calls, "value": float(v)}
    calls = []
    f4 = recorder(calls)
    v = cm.compute_flow_matching_loss_fixed(stub_dit(), full, prompt, mask, fixed_sigmas=[0.3, 0.9], noise_draws=2,
                                            device="cpu", dtype=dtype, forward_fn=lambda h, t: f4(h, t))
    out["fixed"] = {"calls": calls, "value": float(v)}
    return out


def main():
    cm = ref_bridge.load("common")
    out = {"float32": run_all(cm, torch.float32), "bfloat16": run_all(cm, torch.bfloat16)}
    for k, v in out.items():
        print(k, {n: (len(e["calls"]), e.get("value", None) if "value" in e else [float(x) for x in e["losses"]]) for n, e in v.items()})
        print("  timestep of the first conditioned call:", v["conditioned"]["calls"][0]["timestep"].tolist(), v["conditioned"]["calls"][0]["n_cond"])
    torch.save(out, ROOT / "tests" / "golden" / "loss_inputs.pt")


if __name__ == "__main__":
    main()
