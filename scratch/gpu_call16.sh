timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "rope or norm" 2>&1 | tail -3
timeout 200 python scratch/bench_elem.py 2>&1 | tee gpurun_out/elem3.txt
