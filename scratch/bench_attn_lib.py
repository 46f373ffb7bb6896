"""Library baselines for self-attention at the headline shape (N=37440, H=32, D=128, context 6240): the kernels the
reference's GPU step would run -- torch SDPA (cuDNN / flash backends) and flash_attn 2.8 -- on the same two-segment
problem (context queries x context keys; noised queries x all keys), forward and backward, CUDA events."""
import sys, torch, torch.nn.functional as F
from torch.nn.attention import sdpa_kernel, SDPBackend
BF16 = torch.bfloat16
N, H, D, Nc = 37440, 32, 128, 6240
flops = 4.0 * H * D * (Nc * Nc + (N - Nc) * N)
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(N, 3, H, D, generator=g, device="cuda").to(BF16)
do = torch.randn(N, H, D, generator=g, device="cuda").to(BF16)


def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def sdpa_pair(backend):
    # [1, H, N, D] views of the packed qkv (strided, like the engine's)
    q, k, v = (qkv[:, i].permute(1, 0, 2).unsqueeze(0).detach().requires_grad_(True) for i in range(3))
    dot = do.permute(1, 0, 2).unsqueeze(0)

    def fwd():
        with sdpa_kernel(backend):
            o1 = F.scaled_dot_product_attention(q[:, :, :Nc], k[:, :, :Nc], v[:, :, :Nc])
            o2 = F.scaled_dot_product_attention(q[:, :, Nc:], k, v)
        return o1, o2

    def fwd_bwd():
        o1, o2 = fwd()
        torch.autograd.grad([o1, o2], [q, k, v], [dot[:, :, :Nc], dot[:, :, Nc:]])
    with torch.no_grad():
        f = timeit(fwd)
    fb = timeit(fwd_bwd)
    return f, fb - f


for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash(torch)", SDPBackend.FLASH_ATTENTION),
                 ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
    try:
        f, b = sdpa_pair(be)
        print(f"sdpa[{name}] fwd {f:.2f} ms {flops / f / 1e9:.0f} TFLOP/s | bwd {b:.2f} ms {2.5 * flops / b / 1e9:.0f} TFLOP/s", flush=True)
    except Exception as e:  # backend not available for this shape / arch
        print(f"sdpa[{name}] unavailable: {type(e).__name__}: {str(e)[:200]}", flush=True)
try:
    from flash_attn import flash_attn_func
    q, k, v = (qkv[:, i].unsqueeze(0).detach().requires_grad_(True) for i in range(3))

    def ffwd():
        return flash_attn_func(q[:, :Nc], k[:, :Nc], v[:, :Nc]), flash_attn_func(q[:, Nc:], k, v)

    def ffb():
        o1, o2 = ffwd()
        torch.autograd.grad([o1, o2], [q, k, v], [do[None, :Nc], do[None, Nc:]])
    with torch.no_grad():
        f = timeit(ffwd)
    fb = timeit(ffb)
    b = fb - f
    print(f"flash_attn 2.8 fwd {f:.2f} ms {flops / f / 1e9:.0f} TFLOP/s | bwd {b:.2f} ms {2.5 * flops / b / 1e9:.0f} TFLOP/s")
except Exception as e:
    print(f"flash_attn unavailable: {type(e).__name__}: {str(e)[:200]}")
