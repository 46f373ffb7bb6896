timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attn_bwd and fused" 2>&1 | tail -3
B200TTA_ATTN_BWD=fused timeout 120 python scratch/bench_attn.py 2>&1 | head -1
B200TTA_LIB=longcat_video_tta_b200/libb200tta_debug.so timeout 150 python scratch/bwd_timeline_fused.py > gpurun_out/bwd_timeline_fused.txt 2>&1; tail -18 gpurun_out/bwd_timeline_fused.txt
