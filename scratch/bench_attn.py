"""Micro-benchmark of the attention kernels at the headline shape (N=37440, H=32, D=128, context 6240)."""
import sys, torch
sys.path.insert(0, '.')
from longcat_video_tta_b200 import ops
BF16, F32 = torch.bfloat16, torch.float32
N, H, D, Nc = 37440, 32, 128, 6240
if len(sys.argv) > 1: N, Nc = int(sys.argv[1]), int(sys.argv[2])
segs = [(0, Nc, Nc), (Nc, N, N)]
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(N, 3, H, D, generator=g, device="cuda")).to(BF16)
q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
o = torch.zeros(N, H, D, dtype=BF16, device="cuda"); lse = torch.zeros(H, N, dtype=F32, device="cuda")
do = torch.randn(N, H, D, generator=g, device="cuda").to(BF16)
dqkv = torch.zeros_like(qkv); delta = torch.zeros(H, N, dtype=F32, device="cuda")
flops = 4.0 * H * D * (Nc * Nc + (N - Nc) * N)
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
f = timeit(lambda: ops.attn_fwd(q, k, v, o, lse, segs, D ** -0.5))
b = timeit(lambda: ops.attn_bwd(dqkv[:, 0], dqkv[:, 1], dqkv[:, 2], do, o, lse, delta, q, k, v, segs, D ** -0.5))
print(f"attn_fwd {f:.2f} ms  {flops / f / 1e9:.0f} TFLOP/s | attn_bwd {b:.2f} ms  {2.5 * flops / b / 1e9:.0f} TFLOP/s (algorithmic)")
# quick correctness spot check on a slice (fp32 reference on 2 heads, first 512 noise queries)
with torch.no_grad():
    hs = slice(0, 2); qs = slice(Nc, Nc + 512)
    s = torch.einsum("qhd,khd->hqk", q[qs, hs].float(), k[:, hs].float()) * D ** -0.5
    ref = torch.einsum("hqk,khd->qhd", torch.softmax(s, -1), v[:, hs].float())
    err = (o[qs, hs].float() - ref).abs().max().item()
    print(f"fwd spot-check max err {err:.4g} (ref max {ref.abs().max().item():.3g})")
import ctypes as C
from longcat_video_tta_b200 import _lib
lib = _lib.load()
buf = (C.c_longlong * 32)()
if hasattr(lib, "b200tta_debug_read_fwd"):
    lib.b200tta_debug_read_fwd.argtypes = [C.c_void_p]
    if lib.b200tta_debug_read_fwd(buf) == 0:
        d = list(buf); n = max(d[3], 1)
        print(f"fwd CTA(0,0): n_blocks {d[3]}")
        for off, nm in ((8, "tile0"), (16, "tile1")):
            print(f"  softmax {nm}: {d[off]/n:.0f}/block: wait s_full {d[off+1]/n:.0f}, tmem ld {d[off+2]/n:.0f}, max(+rescale) {d[off+3]/n:.0f}, exp+st+arrive {d[off+4]/n:.0f}")
if not hasattr(lib, "b200tta_debug_read"):
    sys.exit(0)   # library built without -DB200TTA_ATTN_DEBUG=1
lib.b200tta_debug_read.argtypes = [C.c_void_p]
if lib.b200tta_debug_read(buf) == 0:
    d = list(buf)
    n = max(d[18], 1); m = n
    print(f"dkv CTA(0,0): issuer own-subs {d[18]}; issuer total {d[16]} clk ({d[16]/n:.0f}/own-sub), waiting p/ds_full {d[17]/n:.0f}")
    print(f"  compute warp 3: total {d[20]} ({d[20]/m:.0f}/own-sub): wait s_full {d[21]/m:.0f}, stats {d[22]/m:.0f}, exp+P^T {d[23]/m:.0f}, dP^T wait + dS^T {d[24]/m:.0f}")
import os
if not os.environ.get("B200TTA_DEBUG_NO_DQ") and not os.environ.get("B200TTA_DEBUG_NO_DKV"):
    for var in ("B200TTA_DEBUG_NO_DQ", "B200TTA_DEBUG_NO_DKV"):
        os.environ[var] = "1"
        t = timeit(lambda: ops.attn_bwd(dqkv[:, 0], dqkv[:, 1], dqkv[:, 2], do, o, lse, delta, q, k, v, segs, D ** -0.5))
        print(f"{var}: {t:.2f} ms")
        del os.environ[var]
