PARITY_REPORT_ONLY=1 timeout 600 python -m pytest tests/test_edge_cases_gpu.py -m gpu -q -s 2>&1 | grep -v "LoRA target" | tail -30
