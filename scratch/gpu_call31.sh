timeout 900 python -m pytest tests/test_text_encoder_gpu.py tests/test_kernels_gpu.py tests/test_step_gpu.py -m gpu -q 2>&1 | tail -4
timeout 900 python scratch/bench_text_encoder.py gpurun_out/text_encoder_xxl.json > gpurun_out/text_encoder_xxl.log 2>&1; grep -A3 '"ours"' gpurun_out/text_encoder_xxl.log; grep -B1 -A3 '"b200tta_gemm"' gpurun_out/text_encoder_xxl.log; grep "ours_rel_l2\"" gpurun_out/text_encoder_xxl.log
timeout 300 python scratch/prof_text_encoder.py
mkdir -p /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"t5_attn_kernel|t5_rmsnorm_kernel|gemm_kernel" --launch-skip 30 -c 17 -o /tmp/ncu/text_encoder python scratch/prof_text_encoder.py > gpurun_out/ncu_text_encoder.log 2>&1; tail -2 gpurun_out/ncu_text_encoder.log
ncu -i /tmp/ncu/text_encoder.ncu-rep --page raw --csv > gpurun_out/r2_ncu_text_encoder.csv 2>/dev/null; ls -la gpurun_out/r2_ncu_text_encoder.csv
