timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn or bsa" 2>&1 | tail -5
timeout 300 python scratch/bench_attn.py 2>&1 | grep -v "^spot" | head -4
