timeout 300 python -m pytest tests/test_latents_gpu.py -m gpu -q 2>&1 | tail -8
