"""Timeline of one CTA of the forward attention kernel (library built with B200TTA_ATTN_DEBUG=1): SM-clock stamps of the
softmax / issuer hand-over points for 24 consecutive K/V blocks, printed relative to the first stamp."""
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
for _ in range(3): ops.attn_fwd(q, k, v, o, lse, segs, D ** -0.5)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * (24 * 16))()
lib.b200tta_debug_fwd_timeline.argtypes = [C.c_void_p]
assert lib.b200tta_debug_fwd_timeline(buf) == 0
t = [list(buf[i * 16:(i + 1) * 16]) for i in range(24)]
t0 = min(x for r in t for x in r if x)
names = ["s0 S seen", "s0 loaded", "s0 max", "s0 P half", "s0 P full", "s1 S seen", "s1 loaded", "s1 max", "s1 P half", "s1 P full",
         "is0 half", "is0 full", "is0 S'iss", "is1 half", "is1 full", "is1 S'iss"]
print("block " + " ".join(f"{n:>10s}" for n in names))
for i, r in enumerate(t):
    print(f"{100 + i:5d} " + " ".join(f"{(x - t0) if x else -1:10d}" for x in r))
per = [(t[i + 1][0] - t[i][0]) for i in range(23)]
print("period per block (tile 0 S seen):", per)
for nm, a, b in (("tile0 softmax (S seen -> P full)", 0, 4), ("tile1 softmax", 5, 9), ("tile0 ld", 0, 1), ("tile0 max", 1, 2), ("tile0 exp 1st half", 2, 3),
                 ("tile0 exp 2nd half", 3, 4), ("tile1 ld", 5, 6), ("tile1 max", 6, 7), ("tile1 exp 1st half", 7, 8), ("tile1 exp 2nd half", 8, 9)):
    d = [r[b] - r[a] for r in t]
    print(f"{nm:34s} mean {sum(d) / len(d):7.0f}  min {min(d)}  max {max(d)}")
