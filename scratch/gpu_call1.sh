set -x
L=longcat_video_tta_b200
B200TTA_LIB=$L/libb200tta_debug.so timeout 150 python scratch/bwd_timeline.py > gpurun_out/bwd_timeline.txt 2>&1; tail -30 gpurun_out/bwd_timeline.txt
timeout 120 python scratch/bench_attn.py 2>&1 | head -3 | tee gpurun_out/attn_release.txt
B200TTA_LIB=$L/libb200tta_lazy.so timeout 120 python scratch/bench_attn.py 2>&1 | head -3 | tee gpurun_out/attn_lazy.txt
B200TTA_LIB=$L/libb200tta_dqpoly.so timeout 120 python scratch/bench_attn.py 2>&1 | head -3 | tee gpurun_out/attn_dqpoly.txt
B200TTA_LIB=$L/libb200tta_lazy.so timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_bsa_gpu.py -q -k "attn or bsa" 2>&1 | tail -3 | tee gpurun_out/lazy_tests.txt
B200TTA_LIB=$L/libb200tta_dqpoly.so timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_bsa_gpu.py -q -k "attn or bsa" 2>&1 | tail -3 | tee gpurun_out/dqpoly_tests.txt
timeout 300 python scratch/bench_attn_lib.py 2>&1 | tee gpurun_out/attn_lib.txt
