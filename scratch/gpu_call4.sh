timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attn_bwd" 2>&1 | tail -15
B200TTA_ATTN_BWD=split timeout 120 python scratch/bench_attn.py 2>&1 | head -1
B200TTA_ATTN_BWD=fused timeout 120 python scratch/bench_attn.py 2>&1 | head -1
