"""Micro-benchmark of the tcgen05 GEMM at the headline shapes (N = 37 440 token rows)."""
import sys, torch
sys.path.insert(0, '.')
from longcat_video_tta_b200 import ops
BF16 = torch.bfloat16
M = 37440
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s, scale=1.0): return (torch.randn(*s, generator=g, device="cuda") * scale).to(BF16)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n
for name, N, K, mn in (("qkv  fwd  [M,4096]x[12288,4096]^T", 12288, 4096, False), ("proj fwd  [M,4096]x[4096,4096]^T", 4096, 4096, False),
                       ("w2   fwd  [M,11008]x[4096,11008]^T", 4096, 11008, False), ("qkv  dX   [M,12288]x[12288,4096]", 4096, 12288, True),
                       ("w2   dX   [M,4096]x[4096,11008]", 11008, 4096, True)):
    a = rnd(M, K)
    w = rnd(K, N, scale=0.02) if mn else rnd(N, K, scale=0.02)
    out = torch.empty(M, N, dtype=BF16, device="cuda")
    t = timeit(lambda: ops.gemm(M, N, [(a, w, K, mn, None)], ops.epi(ops.EPI_STORE, out)))
    tc = timeit(lambda: torch.matmul(a, w if mn else w.t(), out=out))
    print(f"{name:40s} {t:7.3f} ms {2.0 * M * N * K / t / 1e9:7.0f} TFLOP/s | cuBLAS {tc:7.3f} ms {2.0 * M * N * K / tc / 1e9:7.0f} TFLOP/s")
