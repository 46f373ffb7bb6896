timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "ln or norm" 2>&1 | tail -3
timeout 200 python scratch/bench_elem.py 2>&1 | head -2
timeout 300 python -m pytest tests/test_step_gpu.py tests/test_adapters_gpu.py -q -x 2>&1 | tail -3
