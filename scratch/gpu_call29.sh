timeout 600 python -m pytest tests/test_text_encoder_gpu.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -25
timeout 900 python scratch/bench_text_encoder.py gpurun_out/text_encoder_xxl.json > gpurun_out/text_encoder_xxl.log 2>&1; tail -60 gpurun_out/text_encoder_xxl.log
