"""Timelines of one CTA of the dq and of the dkv kernel (library built with B200TTA_ATTN_DEBUG=1): SM-clock stamps of the
hand-over points between the compute warps and the MMA issuers, printed relative to the first stamp.
dq : 24 consecutive K/V blocks of query tile 100.   dkv: 48 consecutive 64-row query sub-tiles of K/V block 100
(even sub-tiles belong to compute group / issuer 0, odd ones to group / issuer 1)."""
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
    ops.attn_bwd(dqkv[:, 0], dqkv[:, 1], dqkv[:, 2], do, o, lse, delta, q, k, v, segs, D ** -0.5)
torch.cuda.synchronize()
lib = _lib.load()
dq_buf, dkv_buf = (C.c_longlong * (24 * 16))(), (C.c_longlong * (48 * 16))()
lib.b200tta_debug_bwd_timeline.argtypes = [C.c_void_p, C.c_void_p]
assert lib.b200tta_debug_bwd_timeline(dq_buf, dkv_buf) == 0


def show(title, buf, rows, first, names, deltas, period_slot, stride=1):
    t = [list(buf[i * 16:(i + 1) * 16]) for i in range(rows)]
    t0 = min(x for r in t for x in r if x)
    print(title)
    print("index " + " ".join(f"{n:>11s}" for n in names.values()))
    for i, r in enumerate(t):
        print(f"{first + i:5d} " + " ".join(f"{(r[s] - t0) if r[s] else -1:11d}" for s in names))
    per = [t[i + stride][period_slot] - t[i][period_slot] for i in range(rows - stride)]
    print(f"period ({names[period_slot]} to the same event {stride} row(s) later): mean {sum(per) / len(per):.0f}  min {min(per)}  max {max(per)}")
    for nm, a, b in deltas:
        d = [r[b] - r[a] for r in t if r[a] and r[b]]
        print(f"  {nm:44s} mean {sum(d) / len(d):7.0f}  min {min(d)}  max {max(d)}")
    print()


show("dq kernel, query tile 100 (tensor work per K/V block: 3 x 512 clk)", dq_buf, 24, 100,
     {0: "S seen", 1: "S loaded", 2: "exp done", 3: "dP seen", 4: "dS stored", 8: "i:S freed", 9: "i:S' issued", 10: "i:dS seen", 11: "i:dQ,dP' iss"},
     [("load S", 0, 1), ("exponentials", 1, 2), ("wait for dP", 2, 3), ("dS = P (dP - delta), store", 3, 4),
      ("issuer: dS signalled -> seen", 4, 10), ("issuer: dQ + dP' issue", 10, 11)], 0)
show("dkv kernel, K/V block 100 (tensor work per 64-row sub-tile: 2 x 384 SS + 2 x 256 TS = 1280 clk)", dkv_buf, 48, 200,
     {0: "stats", 1: "S^T seen", 2: "P^T stored", 3: "dP^T seen", 4: "dS^T stored", 8: "i:P^T seen", 9: "i:dV,S' iss", 10: "i:dS^T seen", 11: "i:dK,dP' iss"},
     [("exponentials + P^T store", 1, 2), ("wait for dP^T", 2, 3), ("dS^T + store", 3, 4),
      ("issuer: P^T signalled -> seen", 2, 8), ("issuer: dV + S^T' issue", 8, 9), ("issuer: dS^T signalled -> seen", 4, 10),
      ("issuer: dK + dP^T' issue", 10, 11)], 1, stride=2)
