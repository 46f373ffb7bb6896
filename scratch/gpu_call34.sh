timeout 900 python -m pytest tests/test_headline_invariants_gpu.py -m gpu -q -s 2>&1 | grep -v "LoRA target" | tail -12
