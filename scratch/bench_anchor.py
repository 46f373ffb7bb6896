"""Early-stopper anchor check at the headline geometry (13.6B, 480p, 4 context + 5 validation latent frames = 14 040
tokens, 3 sigmas x 2 draws): six full forwards vs one full + five noise-rows-only forwards on the context K/V cache."""
import os, sys, time, torch
sys.path.insert(0, '.')
from longcat_video_tta_b200.common import compute_flow_matching_loss_conditioned_fixed
from longcat_video_tta_b200.dit import B200DiT
BF16 = torch.bfloat16
dev = torch.device("cuda", 0)
dit = B200DiT.random_init("13.6b", seed=0, device=dev)
g = torch.Generator().manual_seed(1)
cond = torch.randn(1, 16, 4, 60, 104, generator=g).to(BF16).to(dev)
val = torch.randn(1, 16, 5, 60, 104, generator=g).to(BF16).to(dev)
prompt = torch.randn(1, 1, 512, dit.config.caption_channels, generator=g).to(BF16).to(dev)
mask = torch.ones(1, 512, dtype=torch.int64, device=dev)
noises = [torch.randn(val.shape, generator=torch.Generator().manual_seed(100 + d)).to(BF16).to(dev) for d in range(2)]
args = (dit, cond, val, prompt, mask, [0.25, 0.5, 0.75], noises)


def timed(label):
    compute_flow_matching_loss_conditioned_fixed(*args, device="cuda", dtype=BF16)   # warm-up (workspace, caches)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        v = compute_flow_matching_loss_conditioned_fixed(*args, device="cuda", dtype=BF16)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / 3
    print(f"{label:28s} anchor loss {v:.6f}   {dt * 1e3:8.1f} ms per check (6 forwards)", flush=True)
    return dt


a = timed("context K/V cache")
os.environ["B200TTA_NO_CTX_CACHE"] = "1"
b = timed("six full forwards")
print(f"speed-up {b / a:.2f}x")
