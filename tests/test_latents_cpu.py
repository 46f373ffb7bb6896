"""CPU: the latent-normalisation oracle against the reference's own outputs (tests/golden/latent_norm.pt, made by
oracle/make_golden_latent_norm.py), the host side of longcat_video_tta_b200.latents (statistics, wrappers around the
caller's VAE), and its refusal to run without a B200."""
from types import SimpleNamespace

import pytest
import torch

from oracle import latent_oracle as lo


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(golden_dir / "latent_norm.pt")


def fake_vae(g):
    return SimpleNamespace(config=SimpleNamespace(z_dim=16, latents_mean=g["mean"], latents_std=g["std"]), dtype=torch.bfloat16)


def test_oracle_is_bit_exact_with_the_reference(golden):
    assert len(golden["cases"]) == 4
    for c in golden["cases"]:
        y = lo.normalize(c["x"], golden["mean"], golden["std"])
        assert y.dtype == c["x"].dtype and torch.equal(y, c["normalized"])
        assert torch.equal(lo.denormalize(y, golden["mean"], golden["std"]), c["denormalized"])


def test_channel_statistics_are_built_in_the_latent_dtype(golden):
    from longcat_video_tta_b200.latents import _channel_stats
    vae = fake_vae(golden)
    for dtype in (torch.bfloat16, torch.float32):
        x = torch.zeros(1, 16, 1, 2, 2, dtype=dtype)
        mean, inv = _channel_stats(vae, x)
        m_ref, inv_ref = lo.channel_stats(golden["mean"], golden["std"], dtype)
        assert mean.dtype == inv.dtype == torch.float32
        assert torch.equal(mean, m_ref.float().reshape(-1)) and torch.equal(inv, inv_ref.float().reshape(-1))
    with pytest.raises(ValueError):
        _channel_stats(vae, torch.zeros(1, 8, 1, 2, 2))


def test_no_cpu_path(golden):
    from longcat_video_tta_b200 import _lib
    from longcat_video_tta_b200.latents import normalize_latents
    with pytest.raises(_lib.B200TTAError):
        normalize_latents(fake_vae(golden), torch.zeros(1, 16, 1, 2, 2, dtype=torch.bfloat16))


def test_encode_and_decode_wrap_the_callers_vae(golden, monkeypatch):
    """common.py:158-172, 208-221: encode -> retrieve -> normalise; denormalise -> decode in the VAE's dtype -> [0, 1]"""
    from longcat_video_tta_b200 import latents as L
    calls = []
    monkeypatch.setattr(L, "normalize_latents", lambda vae, z: calls.append("norm") or z + 1)
    monkeypatch.setattr(L, "denormalize_latents", lambda vae, z: calls.append("denorm") or z - 1)
    vae = fake_vae(golden)
    vae.encode = lambda px: SimpleNamespace(latent_dist=SimpleNamespace(sample=lambda: px.mean(1, keepdim=True).expand(-1, 16, -1, -1, -1)))
    vae.decode = lambda z, return_dict=False: (z[:, :3].float() * 4.0,)
    px = torch.rand(1, 3, 2, 4, 4)
    z = L.encode_video(vae, px)
    assert calls == ["norm"] and z.shape == (1, 16, 2, 4, 4)
    assert torch.equal(L.encode_video(vae, px, normalize=False) + 1, z)
    assert torch.equal(L.encode_video(vae, px, retrieve_latents=lambda post: torch.zeros(1, 16, 2, 4, 4)), torch.ones(1, 16, 2, 4, 4))
    video = L.decode_latents(vae, torch.full((1, 16, 2, 4, 4), 1.1))
    assert calls[-1] == "denorm" and video.shape == (1, 3, 2, 4, 4)
    assert float(video.min()) >= 0.0 and float(video.max()) <= 1.0
    assert torch.allclose(video, torch.full_like(video, ((1.1 - 1) * 4 + 1) / 2), atol=2e-2)     # decoded in bf16
