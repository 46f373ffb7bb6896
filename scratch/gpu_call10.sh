timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attn_bwd or weight_gradient or ln_mod or layernorm" 2>&1 | tail -4
B200TTA_ATTN_BWD=split timeout 120 python scratch/bench_attn.py 2>&1 | head -1
B200TTA_ATTN_BWD=fused timeout 120 python scratch/bench_attn.py 2>&1 | head -1
timeout 200 python scratch/bench_elem.py 2>&1 | tee gpurun_out/elem2.txt
timeout 300 python -m pytest tests/test_step_gpu.py -q -k cuda_graph -s 2>&1 | grep -v "LoRA target" | tail -6
timeout 200 python scratch/bench_graph.py tiny 2>&1 | grep -v "LoRA target" | tee gpurun_out/graph_tiny.txt
