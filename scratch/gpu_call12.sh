timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "weight_gradient" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_full_tta_gpu.py -q -x -s 2>&1 | grep -v "LoRA target" | tail -30
