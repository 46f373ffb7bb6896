timeout 300 python -m pytest tests/test_step_gpu.py tests/test_loops_gpu.py tests/test_norm_delta_gpu.py -q 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl torch_gpu --steps 2 --warmup 2 > gpurun_out/bench_torch_gpu.json 2> gpurun_out/bench_torch_gpu.err; tail -3 gpurun_out/bench_torch_gpu.err; cat gpurun_out/bench_torch_gpu.json
