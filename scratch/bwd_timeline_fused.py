"""Timeline of one CTA of the fused backward kernel (library built with B200TTA_ATTN_DEBUG=1): 48 consecutive query tiles
of K/V block 100, head 0; SM-clock stamps at the hand-over points, relative to the first."""
import sys, ctypes as C, torch
sys.path.insert(0, '.')
from longcat_video_tta_b200 import ops, _lib
BF16, F32 = torch.bfloat16, torch.float32
N, H, D, Nc = 37440, 32, 128, 6240
segs = [(0, Nc, Nc), (Nc, N, N)]
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(N, 3, H, D, generator=g, device="cuda").to(BF16)
q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
o = torch.zeros(N, H, D, dtype=BF16, device="cuda"); lse = torch.zeros(H, N, dtype=F32, device="cuda")
do = torch.randn(N, H, D, generator=g, device="cuda").to(BF16)
dqkv = torch.zeros_like(qkv); delta = torch.zeros(H, N, dtype=F32, device="cuda")
ops.attn_fwd(q, k, v, o, lse, segs, D ** -0.5)
for _ in range(2):
    ops.attn_bwd(dqkv[:, 0], dqkv[:, 1], dqkv[:, 2], do, o, lse, delta, q, k, v, segs, D ** -0.5, fused=True)
torch.cuda.synchronize()
lib = _lib.load()
dq_buf, buf = (C.c_longlong * (24 * 16))(), (C.c_longlong * (48 * 16))()
lib.b200tta_debug_bwd_timeline.argtypes = [C.c_void_p, C.c_void_p]
assert lib.b200tta_debug_bwd_timeline(dq_buf, buf) == 0
names = {3: "dP(t) seen", 4: "dS(t) stored", 5: "stageB(t-1)", 0: "S(t+1) seen", 1: "exp1 done", 6: "dQ(t) seen", 7: "stageA done",
         2: "P(t+1) stored", 10: "i:P(t) seen", 11: "i:dV,S' iss", 12: "i:dS seen", 13: "i:dQ,dK iss", 8: "i:dq_free", 9: "i:dP' iss"}
t = [list(buf[i * 16:(i + 1) * 16]) for i in range(48)]
t0 = min(x for r in t for x in r if x)
print("tile  " + " ".join(f"{n:>12s}" for n in names.values()))
for i, r in enumerate(t):
    print(f"{100 + i:5d} " + " ".join(f"{(r[s] - t0) if r[s] else -1:12d}" for s in names))
per = [t[i + 1][3] - t[i][3] for i in range(47)]
print(f"period (dP^T seen -> next dP^T seen): mean {sum(per) / len(per):.0f} min {min(per)} max {max(per)}")
for nm, a, b in (("dS(t): load, math, smem store", 3, 4), ("   dP^T load", 3, 14), ("   math (+ wait dS tile free)", 14, 15), ("   shuffle + smem store", 15, 4), ("stage second half of dQ(t-1)", 4, 5), ("wait S^T(t+1)", 5, 0),
                 ("exponentials, first half", 0, 1), ("wait dQ^T(t)", 1, 6), ("dQ^T load + stage first half", 6, 7),
                 ("exponentials, second half + P^T store", 7, 2), ("issuer: P^T signalled(prev) -> dV,S' issued", 10, 11),
                 ("issuer: dS signalled -> seen", 4, 12), ("issuer: dQ,dK issue", 12, 13), ("issuer: wait dq_free/dO", 13, 8),
                 ("issuer: dP' issue", 8, 9)):
    d = [r[b] - r[a] for r in t if r[a] and r[b]]
    print(f"  {nm:44s} mean {sum(d) / len(d):7.0f}  min {min(d)}  max {max(d)}")
d = [t[i + 1][3] - t[i][2] for i in range(47) if t[i][2] and t[i + 1][3]]
print(f"  {'P^T(t+1) stored -> dP^T(t+1) seen':44s} mean {sum(d) / len(d):7.0f}  min {min(d)}  max {max(d)}")
