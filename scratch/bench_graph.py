"""Launch-by-launch vs CUDA-graph replay of forward + adapter backward (TTAStepper(cuda_graph=...)).
usage: python scratch/bench_graph.py tiny|13.6b"""
import sys, time, torch
sys.path.insert(0, '.')
from longcat_video_tta_b200 import lora, ops
from longcat_video_tta_b200.dit import B200DiT
from longcat_video_tta_b200.stepper import TTAStepper
BF16 = torch.bfloat16
name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
dev = torch.device("cuda", 0)
if name == "tiny":
    dit = B200DiT.random_init("tiny", seed=0, device=dev)
    Tc, Tt, Hl, Wl, Cc, steps = 2, 2, 32, 32, 512, 20
else:
    dit = B200DiT.random_init("13.6b", seed=0, device=dev)
    Tc, Tt, Hl, Wl, Cc, steps = 4, 20, 60, 104, 4096, 3
torch.manual_seed(7)
mods = lora.inject_lora_into_dit(dit, rank=16, alpha=32.0, target_modules=["qkv", "proj"])
g = torch.Generator().manual_seed(1)
cond = torch.randn(1, 16, Tc, Hl, Wl, generator=g).to(BF16).to(dev)
train = torch.randn(1, 16, Tt, Hl, Wl, generator=g).to(BF16).to(dev)
prompt = torch.randn(1, 1, 512, Cc, generator=g).to(BF16).to(dev)
mask = torch.ones(1, 512, dtype=torch.int64, device=dev)
for mode in (False, True):
    torch.manual_seed(7)
    lora.reset_lora_weights(mods)        # both modes start from the same adapters
    st = TTAStepper(dit, cuda_graph=mode)
    sigma = torch.full((1,), 0.4, device=dev)
    noise = torch.randn_like(train)
    for i in range(3):
        st.step(cond, train, prompt, mask, sigma, noise, 1e-4)
    torch.cuda.synchronize()
    k0 = ops.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(steps):
        loss = st.step(cond, train, prompt, mask, sigma, noise, 1e-4)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{name} cuda_graph={mode}: {e0.elapsed_time(e1) / steps:.3f} ms/step on the device, {(t1 - t0) / steps * 1e3:.3f} ms wall, "
          f"{(ops.kernel_launches() - k0) / steps:.0f} launches/step issued from the host, loss {loss.item():.5f}", flush=True)
    del st
