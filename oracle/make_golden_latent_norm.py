"""Test infrastructure, not product code.  Generates tests/golden/latent_norm.pt: the outputs of the reference's own
``normalize_latents`` / ``denormalize_latents`` (delta_experiment/scripts/common.py:175-205, imported unmodified from
/root/reference through oracle/ref_bridge.py) on seeded bf16 and fp32 latents with 16-channel statistics.

Run here (needs /root/reference):  python oracle/make_golden_latent_norm.py"""
import pathlib
import sys
from types import SimpleNamespace

import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_bridge  # noqa: E402

MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632, -0.1922,
        -0.9497, 0.2503, -0.2921]
STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253,
       2.8251, 1.9160]


def main():
    common = ref_bridge.load("common")
    vae = SimpleNamespace(config=SimpleNamespace(z_dim=16, latents_mean=MEAN, latents_std=STD))
    out = {"mean": MEAN, "std": STD, "cases": []}
    for dtype in (torch.bfloat16, torch.float32):
        for shape, seed in (((1, 16, 3, 6, 10), 1), ((2, 16, 1, 5, 7), 2)):
            x = (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * 2.5).to(dtype)
            y = common.normalize_latents(vae, x)
            z = common.denormalize_latents(vae, y)
            out["cases"].append(dict(x=x, normalized=y, denormalized=z))
    torch.save(out, ROOT / "tests" / "golden" / "latent_norm.pt")
    print(len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
