"""Micro-benchmark of the HBM-bound kernels at the headline shape (N=37440 rows, C=4096)."""
import sys, torch
sys.path.insert(0, '.')
from longcat_video_tta_b200 import ops
BF16, F32 = torch.bfloat16, torch.float32
N, C, T, tpf = 37440, 4096, 24, 1560
x = torch.randn(N, C, device="cuda").to(BF16); y = torch.empty_like(x); g = torch.randn(N, C, device="cuda").to(BF16)
dx = torch.randn(N, C, device="cuda").to(BF16)
mod = torch.randn(T, 6 * C, device="cuda", dtype=F32) * 0.1
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n
t = timeit(lambda: ops.ln_mod_fwd(y, x, mod[:, C:2 * C], mod[:, :C], tokens_per_frame=tpf))
print(f"ln_mod_fwd  {t*1e3:7.1f} us  {2 * N * C * 2 / t / 1e9:6.2f} TB/s (algorithmic 2 x N x C x 2 B)")
t = timeit(lambda: ops.ln_mod_bwd(dx, g, x, mod[:, C:2 * C], dx_resid=dx, tokens_per_frame=tpf))
print(f"ln_mod_bwd  {t*1e3:7.1f} us  {4 * N * C * 2 / t / 1e9:6.2f} TB/s (algorithmic 4 x N x C x 2 B)")
t = timeit(lambda: ops.gate_mul(y, dx, mod[:, 2 * C:3 * C], tokens_per_frame=tpf))
print(f"gate_mul    {t*1e3:7.1f} us  {2 * N * C * 2 / t / 1e9:6.2f} TB/s")
H, D = 32, 128
qkv = torch.randn(N, 3 * C, device="cuda").to(BF16); qk = torch.empty(N, 2 * C, dtype=BF16, device="cuda")
dqk = torch.randn(N, 2 * C, device="cuda").to(BF16); dqkv = torch.empty(N, 3 * C, dtype=BF16, device="cuda")
wq = torch.ones(D, dtype=BF16, device="cuda"); wk = torch.ones(D, dtype=BF16, device="cuda")
t = timeit(lambda: ops.qk_rmsnorm_rope_fwd(qk, qkv, wq, wk, H, H, grid_hw=(30, 52), rope_base=10000.0))
print(f"qk_rope_fwd {t*1e3:7.1f} us  {4 * N * C * 2 / t / 1e9:6.2f} TB/s (algorithmic 4 x N x C x 2 B: q, k in and out)")
t = timeit(lambda: ops.qk_rmsnorm_rope_bwd(dqkv, dqk, qkv, wq, wk, H, H, grid_hw=(30, 52), rope_base=10000.0))
print(f"qk_rope_bwd {t*1e3:7.1f} us  {6 * N * C * 2 / t / 1e9:6.2f} TB/s (algorithmic 6 x N x C x 2 B: dq, dk, q, k in; dq, dk out)")
t = timeit(lambda: y.copy_(x))
print(f"torch copy  {t*1e3:7.1f} us  {2 * N * C * 2 / t / 1e9:6.2f} TB/s")
