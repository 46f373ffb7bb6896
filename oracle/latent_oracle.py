"""Oracle restatement of the reference's latent normalisation (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows delta_experiment/scripts/common.py:175-205 (``normalize_latents`` / ``denormalize_latents``): the per-channel mean
and 1 / std are built in the LATENT dtype, and every intermediate is rounded to that dtype.  Pinned: bit-exact against
the reference's own functions on the committed vectors of tests/golden/latent_norm.pt (tests/test_latents_cpu.py)."""
import torch


def channel_stats(mean, std, dtype):
    n = len(mean)
    m = torch.tensor(mean).view(1, n, 1, 1, 1).to(dtype)
    inv = 1.0 / torch.tensor(std).view(1, n, 1, 1, 1).to(dtype)      # the division itself happens in `dtype`
    return m, inv


def normalize(x, mean, std):
    m, inv = channel_stats(mean, std, x.dtype)
    return (x - m) * inv


def denormalize(x, mean, std):
    m, inv = channel_stats(mean, std, x.dtype)
    return x / inv + m
