timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attn_bwd and fused" 2>&1 | tail -2
for i in 1 2; do timeout 120 python scratch/bench_attn.py 2>&1 | head -1; done
