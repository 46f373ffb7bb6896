PARITY_REPORT_ONLY=1 timeout 900 python -m pytest tests/test_block_parity_gpu.py -q -s -x 2>&1 | grep -v "LoRA target blocks" > gpurun_out/parity_report.txt; tail -5 gpurun_out/parity_report.txt
timeout 600 python -m pytest tests/test_loops_gpu.py -q -s 2>&1 | grep -v "LoRA target blocks" > gpurun_out/loops_report.txt; tail -30 gpurun_out/loops_report.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_block_parity_gpu.py --deselect tests/test_loops_gpu.py 2>&1 | tail -8
