timeout 600 python -m pytest tests/test_text_encoder_gpu.py -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py --workload text_encode > gpurun_out/bench_text_encode_n1.json 2> gpurun_out/bench_text_encode_n1.err; tail -3 gpurun_out/bench_text_encode_n1.err; cat gpurun_out/bench_text_encode_n1.json | cut -c1-3000
timeout 900 python bench.py --workload text_encode --impl reference > gpurun_out/bench_text_encode_reference.json 2>/dev/null; cat gpurun_out/bench_text_encode_reference.json | cut -c1-600
