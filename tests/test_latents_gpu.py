"""GPU: b200tta_latent_affine / latents.normalize_latents, denormalize_latents bit-exact against the reference's own
outputs (tests/golden/latent_norm.pt) and, at a full 93-frame 480p latent, against the oracle."""
from types import SimpleNamespace

import pytest
import torch

from oracle import latent_oracle as lo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden(golden_dir):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.load(golden_dir / "latent_norm.pt")


def fake_vae(g):
    return SimpleNamespace(config=SimpleNamespace(z_dim=16, latents_mean=g["mean"], latents_std=g["std"]))


def test_bit_exact_with_the_reference_vectors(golden):
    from longcat_video_tta_b200 import ops
    from longcat_video_tta_b200.latents import denormalize_latents, normalize_latents
    vae = fake_vae(golden)
    n0 = ops.kernel_launches()
    for c in golden["cases"]:
        y = normalize_latents(vae, c["x"].cuda())
        assert y.dtype == c["x"].dtype and torch.equal(y.cpu(), c["normalized"])
        z = denormalize_latents(vae, y)
        assert torch.equal(z.cpu(), c["denormalized"])
    assert ops.kernel_launches() - n0 == 2 * len(golden["cases"])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_full_clip_latent_against_the_oracle(golden, dtype):
    """[1, 16, 24, 60, 104]: the latent of a 93-frame 480p clip; strided input (a frame slice) is made dense first"""
    from longcat_video_tta_b200.latents import denormalize_latents, normalize_latents
    vae = fake_vae(golden)
    g = torch.Generator(device="cuda").manual_seed(3)
    full = (torch.randn(1, 16, 30, 60, 104, generator=g, device="cuda") * 3.0).to(dtype)
    x = full[:, :, 4:28]                      # not contiguous
    y = normalize_latents(vae, x)
    ref = lo.normalize(x.cpu(), golden["mean"], golden["std"])
    assert torch.equal(y.cpu(), ref)
    z = denormalize_latents(vae, y)
    assert torch.equal(z.cpu(), lo.denormalize(ref, golden["mean"], golden["std"]))
    # round trip: back to the input up to the dtype's rounding (bf16: 2-3 roundings)
    tol = 3e-2 if dtype == torch.bfloat16 else 1e-5
    assert torch.allclose(z.float(), x.float(), rtol=tol, atol=tol)
