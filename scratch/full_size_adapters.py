"""Full-size (13.6B, 480p, 37 440 tokens) sanity run of the non-LoRA adapter methods of BASELINE.json configs[2]:
two optimisation steps each through the drop-in optimize_* entry points, wall-clock per step and peak memory."""
import sys, time, torch
sys.path.insert(0, '.')
from longcat_video_tta_b200 import adapters as A
from longcat_video_tta_b200.dit import B200DiT
BF16 = torch.bfloat16
dev = torch.device("cuda", 0)
dit = B200DiT.random_init("13.6b", seed=0, device=dev)
g = torch.Generator().manual_seed(1)
cond = torch.randn(1, 16, 4, 60, 104, generator=g).to(BF16).to(dev)
train = torch.randn(1, 16, 20, 60, 104, generator=g).to(BF16).to(dev)
prompt = torch.randn(1, 1, 512, dit.config.caption_channels, generator=g).to(BF16).to(dev)
mask = torch.ones(1, 512, dtype=torch.int64, device=dev)
C = dit.config.hidden_size


def run(name, make, optimize):
    w = make()
    torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats()
    torch.manual_seed(42)
    t0 = time.time()
    out = optimize(w)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / len(out["losses"])
    print(f"{name:22s} losses {['%.4f' % l for l in out['losses']]}  {dt:6.2f} s/step  peak {torch.cuda.max_memory_allocated() / 2**30:6.1f} GiB", flush=True)
    if hasattr(w, "remove_from_dit"):
        w.remove_from_dit()


kw = dict(num_steps=2, lr=1e-3, device="cuda")
run("delta_a", lambda: A.DeltaAWrapper(dit, 512), lambda w: A.optimize_delta_a(w, cond, train, prompt, mask, **kw))
run("delta_b timestep G=4", lambda: A.DeltaBWrapper(dit, num_groups=4, adaln_tembed_dim=512, hidden_size=C, delta_target="timestep"),
    lambda w: A.optimize_delta_b(w, cond, train, prompt, mask, **kw))
run("delta_c", lambda: A.DeltaCWrapper(dit, mode="per_channel", out_channels=16),
    lambda w: A.optimize_delta_c(w, cond, train, prompt, mask, **kw))


def make_norm():
    params = A.collect_norm_params(dit, "all_norm")
    w = A.NormTuneForward(dit)
    w._params = params
    return w


run("norm_tune all_norm", make_norm, lambda w: A.optimize_norm_params(w, w._params, cond, train, prompt, mask, **kw))


def make_film():
    w = A.FiLMAdapterWrapper(dit, num_groups=4, hidden_size=C, film_mode="full")
    w.apply_to_dit()
    return w


run("film full G=4", make_film, lambda w: A.optimize_film_adapter(w, cond, train, prompt, mask, **kw))
